// Image loss that sits directly after the hot path in the reference's training step:
// ImageLoss(gradient=True, loss_weight=[w0, w1]) of loss/image_loss.py:10-58 --
//   loss[n] = w0 * mean_{C,H,W} (out - tgt)^2 + w1 * mean_{3,H,W} | m(out[:, :3]) - m(tgt[:, :3]) |,
//   m(x) = sqrt(((x[.,x+1] - x[.,x-1]) / 2)^2 + ((x[y-1,.] - x[y+1,.]) / 2)^2 + 1e-6)   (zero padding).
// NCHW fp32 (the layout of the model's tanh output).  Forward: one pass, per-sample fp32 block sums -> fp64 atomics;
// it also stores, per RGB pixel p, Gx(p) = s 0.25 (r - l) / m_out and Gy(p) = s 0.25 (t - b) / m_out with
// s = sign(m_out - m_tgt), so the backward is a 4-neighbour gather:
//   dout(q) = g[n] (w0 2 (out - tgt) / (C H W) + w1 / (3 H W) (Gx(y,x-1) - Gx(y,x+1) + Gy(y+1,x) - Gy(y-1,x))).
#include "common.cuh"

namespace {

__device__ __forceinline__ float grad_mag(const float* __restrict__ p, int y, int x, int H, int W, float& rl, float& tb) {
  const float r = (x + 1 < W) ? p[y * W + x + 1] : 0.f;
  const float l = (x > 0) ? p[y * W + x - 1] : 0.f;
  const float t = (y > 0) ? p[(y - 1) * W + x] : 0.f;
  const float b = (y + 1 < H) ? p[(y + 1) * W + x] : 0.f;
  rl = r - l;
  tb = t - b;
  const float a = rl * 0.5f, c = tb * 0.5f;
  return sqrtf(a * a + c * c + 1e-6f);
}

// grid: (ceil(C*H*W / 256), N)
__global__ void __launch_bounds__(256)
image_loss_fwd_kernel(const float* __restrict__ out, const float* __restrict__ tgt, float* __restrict__ G, int C, int H,
                      int W, double* __restrict__ acc) {
  __shared__ float red[2][8];
  const int n = blockIdx.y;
  const int HW = H * W;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float se = 0.f, gp = 0.f;
  if (i < C * HW) {
    const int c = i / HW, rem = i - c * HW;
    const int y = rem / W, x = rem - y * W;
    const float* o = out + ((long long)n * C + c) * HW;
    const float* t = tgt + ((long long)n * C + c) * HW;
    const float d = o[rem] - t[rem];
    se = d * d;
    if (c < 3) {
      float rlo, tbo, rlt, tbt;
      const float mo = grad_mag(o, y, x, H, W, rlo, tbo);
      const float mt = grad_mag(t, y, x, H, W, rlt, tbt);
      const float df = mo - mt;
      gp = fabsf(df);
      if (G) {
        const float s = (df > 0.f) ? 1.f : ((df < 0.f) ? -1.f : 0.f);
        const float k = s * 0.25f / mo;
        float2* g = reinterpret_cast<float2*>(G) + ((long long)n * 3 + c) * HW + rem;
        *g = make_float2(k * rlo, k * tbo);
      }
    }
  }
  se = warp_sum(se);
  gp = warp_sum(gp);
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = se;
    red[1][threadIdx.x >> 5] = gp;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[threadIdx.x][k];
    atomicAdd(acc + 2 * n + threadIdx.x, (double)t);
  }
}

__global__ void image_loss_finalize_kernel(const double* __restrict__ acc, float* __restrict__ loss, int N, double inv_mse,
                                           double inv_gp, float w0, float w1) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n < N) loss[n] = (float)((double)w0 * acc[2 * n] * inv_mse + (double)w1 * acc[2 * n + 1] * inv_gp);
}

__global__ void __launch_bounds__(256)
image_loss_bwd_kernel(const float* __restrict__ out, const float* __restrict__ tgt, const float* __restrict__ G,
                      const float* __restrict__ gloss, float* __restrict__ dout, int C, int H, int W, float k_mse,
                      float k_gp) {
  const int n = blockIdx.y;
  const int HW = H * W;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C * HW) return;
  const int c = i / HW, rem = i - c * HW;
  const int y = rem / W, x = rem - y * W;
  const long long base = ((long long)n * C + c) * HW;
  const float g = gloss[n];
  float d = k_mse * (out[base + rem] - tgt[base + rem]);
  if (c < 3) {
    const float2* gm = reinterpret_cast<const float2*>(G) + ((long long)n * 3 + c) * HW;
    float s = 0.f;
    if (x > 0) s += gm[rem - 1].x;
    if (x + 1 < W) s -= gm[rem + 1].x;
    if (y + 1 < H) s += gm[rem + W].y;
    if (y > 0) s -= gm[rem - W].y;
    d = fmaf(k_gp, s, d);
  }
  dout[base + rem] = g * d;
}

}  // namespace

extern "C" {

int tatt_image_loss_fwd(const float* out, const float* tgt, float* loss, float* G, int N, int C, int H, int W, float w0,
                        float w1, void* ws, void* stream) {
  TATT_REQUIRE(N >= 1 && C >= 3 && H >= 1 && W >= 1, "image_loss_fwd: bad shape [%d,%d,%d,%d] (needs >= 3 channels)", N, C,
               H, W);
  TATT_REQUIRE((long long)C * H * W < (1LL << 31) && N <= 65535, "image_loss_fwd: shape too large");
  cudaStream_t st = (cudaStream_t)stream;
  TATT_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 2 * N, st));
  dim3 grid((C * H * W + 255) / 256, N);
  image_loss_fwd_kernel<<<grid, 256, 0, st>>>(out, tgt, G, C, H, W, (double*)ws);
  TATT_LAUNCH_CHECK("image_loss_fwd_kernel");
  image_loss_finalize_kernel<<<(N + 127) / 128, 128, 0, st>>>((const double*)ws, loss, N, 1.0 / ((double)C * H * W),
                                                              1.0 / (3.0 * H * W), w0, w1);
  TATT_LAUNCH_CHECK("image_loss_finalize_kernel");
  return 0;
}

int tatt_image_loss_bwd(const float* out, const float* tgt, const float* G, const float* gloss, float* dout, int N, int C,
                        int H, int W, float w0, float w1, void* stream) {
  TATT_REQUIRE(N >= 1 && C >= 3 && H >= 1 && W >= 1 && N <= 65535, "image_loss_bwd: bad shape [%d,%d,%d,%d]", N, C, H, W);
  dim3 grid((C * H * W + 255) / 256, N);
  image_loss_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(out, tgt, G, gloss, dout, C, H, W,
                                                                2.f * w0 / ((float)C * H * W), w1 / (3.f * H * W));
  TATT_LAUNCH_CHECK("image_loss_bwd_kernel");
  return 0;
}

}  // extern "C"
