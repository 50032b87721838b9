// Rest of the loss block that follows the hot path in the reference's training step:
//   * SemanticLoss        loss/semantic_loss.py:20-37      mean |gt - pred| + KLDivLoss()(log(pred + 1e-20), gt + 1e-20)
//   * TRI_SSIM            utils/ssim_psnr.py:28-37,99-128,231-256   three-image SSIM, 11x11 Gaussian window, zero padding
//   * torch_rotate_img    interfaces/super_resolution.py:126-157    affine_grid + grid_sample (bilinear, zeros,
//                                                                   align_corners=False)
// All NCHW fp32 (the layout of the model's tanh output and of the HR batch).  Small HBM-bound stencils / reductions:
// one pass over the inputs per direction, block partial sums -> fp64 atomics.
#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------ SemanticLoss
// acc[0] += sum |g - p| ; acc[1] += sum t (log t - log(p + 1e-20)), t = g + 1e-20  (F.kl_div pointwise, target > 0)
__global__ void __launch_bounds__(256)
semantic_loss_fwd_kernel(const float* __restrict__ pred, const float* __restrict__ gt, long long n, double* __restrict__ acc) {
  __shared__ float red[2][8];
  float a = 0.f, k = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float p = pred[i], g = gt[i];
    a += fabsf(g - p);
    const float t = g + 1e-20f;
    if (t > 0.f) k += t * (logf(t) - logf(p + 1e-20f));
  }
  a = warp_sum(a);
  k = warp_sum(k);
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = a;
    red[1][threadIdx.x >> 5] = k;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += red[threadIdx.x][q];
    atomicAdd(acc + threadIdx.x, (double)t);
  }
}

__global__ void semantic_loss_finalize_kernel(const double* __restrict__ acc, float* __restrict__ loss, double inv_n) {
  if (threadIdx.x == 0 && blockIdx.x == 0) loss[0] = (float)((acc[0] + acc[1]) * inv_n);
}

// d/dpred = (-sign(g - p) - t / (p + 1e-20)) / n ; d/dgt = (sign(g - p) + log t + 1 - log(p + 1e-20)) / n
__global__ void __launch_bounds__(256)
semantic_loss_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ gt, const float* __restrict__ gloss,
                         float* __restrict__ dpred, float* __restrict__ dgt, long long n, float inv_n) {
  const float s = gloss[0] * inv_n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float p = pred[i], g = gt[i];
    const float d = g - p;
    const float sg = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
    const float t = g + 1e-20f, q = p + 1e-20f;
    if (dpred) dpred[i] = s * (-sg - (t > 0.f ? t / q : 0.f));
    if (dgt) dgt[i] = s * (sg + (t > 0.f ? logf(t) + 1.f - logf(q) : 0.f));
  }
}

// ------------------------------------------------------------------------------------------------ TRI_SSIM
constexpr int SS_R = 5;                 // 11-tap window
constexpr int SS_TH = 16, SS_TW = 32;   // output tile per block
constexpr int SS_IH = SS_TH + 2 * SS_R, SS_IW = SS_TW + 2 * SS_R;
constexpr float SS_C1 = 0.01f * 0.01f, SS_C2 = 0.03f * 0.03f;

struct Gauss11 {
  float w[11];
};

// blurred quantities of one pixel -> ssim value (+ the five partial derivatives the backward pass blurs)
__device__ __forceinline__ float ssim_point(const float (&b)[9], float* g5) {
  // b: mu1 mu2 mu3 e11 e22 e33 e12 e23 e31
  const float m1 = b[0], m2 = b[1], m3 = b[2];
  const float Mx = m1 * m2 + m2 * m3 + m3 * m1, Q = m1 * m1 + m2 * m2 + m3 * m3;
  const float A1 = Mx + SS_C1, B1 = Q + SS_C1;
  const float s12 = b[6] - m1 * m2, s23 = b[7] - m2 * m3, s31 = b[8] - m3 * m1;
  const float s1 = b[3] - m1 * m1, s2 = b[4] - m2 * m2, s3 = b[5] - m3 * m3;
  const float A2 = s12 + s23 + s31 + SS_C2, B2 = s1 + s2 + s3 + SS_C2;
  const float inv = 1.f / (B1 * B2);
  const float S = A1 * A2 * inv;
  if (g5) {
    const float ka = (A2 - A1) * inv, kb = 2.f * S * (B2 - B1) * inv;
    g5[0] = (m2 + m3) * ka - m1 * kb;     // dS / d mu1
    g5[1] = (m3 + m1) * ka - m2 * kb;     // dS / d mu2
    g5[2] = (m1 + m2) * ka - m3 * kb;     // dS / d mu3
    g5[3] = -S / B2;                      // dS / d e_kk
    g5[4] = A1 * inv;                     // dS / d e_kl (k != l)
  }
  return S;
}

// grid: (ceil(W / 32), ceil(H / 16), N * C); 256 threads.  acc[n] += sum of the ssim map of sample n.
// G (optional): [5][N*C*H*W] per-pixel derivatives for the backward pass.
__global__ void __launch_bounds__(256)
tri_ssim_fwd_kernel(const float* __restrict__ x1, const float* __restrict__ x2, const float* __restrict__ x3,
                    float* __restrict__ G, int C, int H, int W, long long plane_total, const Gauss11 gw,
                    double* __restrict__ acc) {
  __shared__ float in[3][SS_IH][SS_IW + 1];
  __shared__ float hb[9][SS_IH][SS_TW + 1];
  __shared__ float red[8];
  const int nc = blockIdx.z, y0 = blockIdx.y * SS_TH, x0 = blockIdx.x * SS_TW;
  const long long base = (long long)nc * H * W;
  const float* src[3] = {x1 + base, x2 + base, x3 + base};
  for (int i = threadIdx.x; i < SS_IH * SS_IW; i += 256) {
    const int r = i / SS_IW, c = i - r * SS_IW;
    const int y = y0 + r - SS_R, x = x0 + c - SS_R;
    const bool ok = y >= 0 && y < H && x >= 0 && x < W;
#pragma unroll
    for (int k = 0; k < 3; ++k) in[k][r][c] = ok ? __ldg(src[k] + (long long)y * W + x) : 0.f;
  }
  __syncthreads();
  // horizontal pass over the 26 x 32 (row, output column) grid
  for (int i = threadIdx.x; i < SS_IH * SS_TW; i += 256) {
    const int r = i / SS_TW, c = i - r * SS_TW;
    float s[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) s[k] = 0.f;
#pragma unroll
    for (int t = 0; t < 11; ++t) {
      const float a = in[0][r][c + t], b = in[1][r][c + t], d = in[2][r][c + t], w = gw.w[t];
      s[0] = fmaf(w, a, s[0]); s[1] = fmaf(w, b, s[1]); s[2] = fmaf(w, d, s[2]);
      s[3] = fmaf(w, a * a, s[3]); s[4] = fmaf(w, b * b, s[4]); s[5] = fmaf(w, d * d, s[5]);
      s[6] = fmaf(w, a * b, s[6]); s[7] = fmaf(w, b * d, s[7]); s[8] = fmaf(w, d * a, s[8]);
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) hb[k][r][c] = s[k];
  }
  __syncthreads();
  float part = 0.f;
  for (int i = threadIdx.x; i < SS_TH * SS_TW; i += 256) {
    const int r = i / SS_TW, c = i - r * SS_TW;
    const int y = y0 + r, x = x0 + c;
    if (y < H && x < W) {
      float b[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        float s = 0.f;
#pragma unroll
        for (int t = 0; t < 11; ++t) s = fmaf(gw.w[t], hb[k][r + t][c], s);
        b[k] = s;
      }
      float g5[5];
      part += ssim_point(b, G ? g5 : nullptr);
      if (G) {
        const long long o = base + (long long)y * W + x;
#pragma unroll
        for (int k = 0; k < 5; ++k) G[(long long)k * plane_total + o] = g5[k];
      }
    }
  }
  part = warp_sum(part);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += red[q];
    atomicAdd(acc + nc / C, (double)t);
  }
}

// per_sample == 0: out[0] = sum_n acc[n] / (N C H W) ; else out[n] = acc[n] / (C H W)
__global__ void tri_ssim_finalize_kernel(const double* __restrict__ acc, float* __restrict__ out, int N, double inv_chw,
                                         int per_sample) {
  if (per_sample) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < N) out[n] = (float)(acc[n] * inv_chw);
  } else if (blockIdx.x == 0 && threadIdx.x == 0) {
    double t = 0.0;
    for (int n = 0; n < N; ++n) t += acc[n];
    out[0] = (float)(t * inv_chw / N);
  }
}

// The Gaussian blur with zero padding is self-adjoint, so with g_* = upstream * dS/d(blurred quantity):
//   d x1 = blur(g_mu1) + 2 x1 blur(g_D) + (x2 + x3) blur(g_X)   (and cyclically for x2, x3)
// gout: [1] (size_average) or [N]; scale = 1 / (N C H W) or 1 / (C H W).
__global__ void __launch_bounds__(256)
tri_ssim_bwd_kernel(const float* __restrict__ x1, const float* __restrict__ x2, const float* __restrict__ x3,
                    const float* __restrict__ G, const float* __restrict__ gout, int per_sample, float scale,
                    float* __restrict__ d1, float* __restrict__ d2, float* __restrict__ d3, int C, int H, int W,
                    long long plane_total, const Gauss11 gw) {
  __shared__ float in[5][SS_IH][SS_IW + 1];
  __shared__ float hb[5][SS_IH][SS_TW + 1];
  const int nc = blockIdx.z, y0 = blockIdx.y * SS_TH, x0 = blockIdx.x * SS_TW;
  const long long base = (long long)nc * H * W;
  for (int i = threadIdx.x; i < SS_IH * SS_IW; i += 256) {
    const int r = i / SS_IW, c = i - r * SS_IW;
    const int y = y0 + r - SS_R, x = x0 + c - SS_R;
    const bool ok = y >= 0 && y < H && x >= 0 && x < W;
#pragma unroll
    for (int k = 0; k < 5; ++k) in[k][r][c] = ok ? __ldg(G + (long long)k * plane_total + base + (long long)y * W + x) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < SS_IH * SS_TW; i += 256) {
    const int r = i / SS_TW, c = i - r * SS_TW;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      float s = 0.f;
#pragma unroll
      for (int t = 0; t < 11; ++t) s = fmaf(gw.w[t], in[k][r][c + t], s);
      hb[k][r][c] = s;
    }
  }
  __syncthreads();
  const float up = gout[per_sample ? nc / C : 0] * scale;
  for (int i = threadIdx.x; i < SS_TH * SS_TW; i += 256) {
    const int r = i / SS_TW, c = i - r * SS_TW;
    const int y = y0 + r, x = x0 + c;
    if (y < H && x < W) {
      float b[5];
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        float s = 0.f;
#pragma unroll
        for (int t = 0; t < 11; ++t) s = fmaf(gw.w[t], hb[k][r + t][c], s);
        b[k] = s;
      }
      const long long o = base + (long long)y * W + x;
      const float a = x1[o], bb = x2[o], d = x3[o];
      if (d1) d1[o] = up * (b[0] + 2.f * a * b[3] + (bb + d) * b[4]);
      if (d2) d2[o] = up * (b[1] + 2.f * bb * b[3] + (d + a) * b[4]);
      if (d3) d3[o] = up * (b[2] + 2.f * d * b[3] + (a + bb) * b[4]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ rotate_img
// theta[n] = [[cos, sin * r, 0], [-sin / r, cos, 0]], r = H / W + offs * 2 * off_range - off_range
__device__ __forceinline__ void rot_src(const float* arcs, const float* offs, float off_range, int n, int H, int W, int y,
                                        int x, float& ix, float& iy) {
  const float r = (float)H / (float)W + (offs[n] * off_range * 2.f) - off_range;
  const float cs = cosf(arcs[n]), sn = sinf(arcs[n]);
  const float gx = (2.f * x + 1.f) / W - 1.f, gy = (2.f * y + 1.f) / H - 1.f;      // affine_grid, align_corners=False
  const float sx = cs * gx + (sn * r) * gy, sy = (-sn / r) * gx + cs * gy;
  ix = ((sx + 1.f) * W - 1.f) * 0.5f;                                              // grid_sample unnormalise
  iy = ((sy + 1.f) * H - 1.f) * 0.5f;
}

__global__ void __launch_bounds__(256)
rotate_img_fwd_kernel(const float* __restrict__ img, const float* __restrict__ arcs, const float* __restrict__ offs,
                      float off_range, float* __restrict__ out, int C, int H, int W) {
  const int n = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * W) return;
  const int y = i / W, x = i - y * W;
  float ix, iy;
  rot_src(arcs, offs, off_range, n, H, W, y, x, ix, iy);
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = (int)fx, y0 = (int)fy;
  const float wx1 = ix - fx, wy1 = iy - fy, wx0 = 1.f - wx1, wy0 = 1.f - wy1;
  const bool vx0 = x0 >= 0 && x0 < W, vx1 = x0 + 1 >= 0 && x0 + 1 < W, vy0 = y0 >= 0 && y0 < H, vy1 = y0 + 1 >= 0 && y0 + 1 < H;
  for (int c = 0; c < C; ++c) {
    const float* p = img + ((long long)n * C + c) * H * W;
    float v = 0.f;
    if (vy0 && vx0) v = fmaf(p[y0 * W + x0], wy0 * wx0, v);
    if (vy0 && vx1) v = fmaf(p[y0 * W + x0 + 1], wy0 * wx1, v);
    if (vy1 && vx0) v = fmaf(p[(y0 + 1) * W + x0], wy1 * wx0, v);
    if (vy1 && vx1) v = fmaf(p[(y0 + 1) * W + x0 + 1], wy1 * wx1, v);
    out[((long long)n * C + c) * H * W + i] = v;
  }
}

// dimg (zero-filled by the launcher) += scatter of dout with the same bilinear weights
__global__ void __launch_bounds__(256)
rotate_img_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ arcs, const float* __restrict__ offs,
                      float off_range, float* __restrict__ dimg, int C, int H, int W) {
  const int n = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * W) return;
  const int y = i / W, x = i - y * W;
  float ix, iy;
  rot_src(arcs, offs, off_range, n, H, W, y, x, ix, iy);
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = (int)fx, y0 = (int)fy;
  const float wx1 = ix - fx, wy1 = iy - fy, wx0 = 1.f - wx1, wy0 = 1.f - wy1;
  const bool vx0 = x0 >= 0 && x0 < W, vx1 = x0 + 1 >= 0 && x0 + 1 < W, vy0 = y0 >= 0 && y0 < H, vy1 = y0 + 1 >= 0 && y0 + 1 < H;
  for (int c = 0; c < C; ++c) {
    float* p = dimg + ((long long)n * C + c) * H * W;
    const float g = dout[((long long)n * C + c) * H * W + i];
    if (vy0 && vx0) atomicAdd(p + y0 * W + x0, g * wy0 * wx0);
    if (vy0 && vx1) atomicAdd(p + y0 * W + x0 + 1, g * wy0 * wx1);
    if (vy1 && vx0) atomicAdd(p + (y0 + 1) * W + x0, g * wy1 * wx0);
    if (vy1 && vx1) atomicAdd(p + (y0 + 1) * W + x0 + 1, g * wy1 * wx1);
  }
}

// the window of utils/ssim_psnr.py:28-31 evaluated like the reference does: fp32 tensor of exp(...) / sum
Gauss11 make_gauss11() {
  Gauss11 g;
  float e[11], s = 0.f;
  for (int x = 0; x < 11; ++x) {
    e[x] = (float)exp(-(double)((x - 5) * (x - 5)) / (2.0 * 1.5 * 1.5));
    s += e[x];
  }
  for (int x = 0; x < 11; ++x) g.w[x] = e[x] / s;
  return g;
}

}  // namespace

extern "C" {

int tatt_semantic_loss_fwd(const float* pred, const float* gt, long long n, float* loss, void* ws, void* stream) {
  TATT_REQUIRE(n >= 1, "semantic_loss_fwd: empty input");
  cudaStream_t st = (cudaStream_t)stream;
  TATT_CUDA(cudaMemsetAsync(ws, 0, 2 * sizeof(double), st));
  const int grid = (int)((n + 255) / 256 < 592 ? (n + 255) / 256 : 592);
  semantic_loss_fwd_kernel<<<grid, 256, 0, st>>>(pred, gt, n, (double*)ws);
  TATT_LAUNCH_CHECK("semantic_loss_fwd_kernel");
  semantic_loss_finalize_kernel<<<1, 32, 0, st>>>((const double*)ws, loss, 1.0 / (double)n);
  TATT_LAUNCH_CHECK("semantic_loss_finalize_kernel");
  return 0;
}

int tatt_semantic_loss_bwd(const float* pred, const float* gt, const float* gloss, float* dpred, float* dgt, long long n,
                           void* stream) {
  TATT_REQUIRE(n >= 1 && (dpred || dgt), "semantic_loss_bwd: nothing to compute");
  const int grid = (int)((n + 255) / 256 < 592 ? (n + 255) / 256 : 592);
  semantic_loss_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pred, gt, gloss, dpred, dgt, n, (float)(1.0 / (double)n));
  TATT_LAUNCH_CHECK("semantic_loss_bwd_kernel");
  return 0;
}

int tatt_tri_ssim_fwd(const float* x1, const float* x2, const float* x3, float* out, float* G, int N, int C, int H, int W,
                      int per_sample, void* ws, void* stream) {
  TATT_REQUIRE(N >= 1 && C >= 1 && H >= 1 && W >= 1 && (long long)N * C <= 65535, "tri_ssim_fwd: bad shape [%d,%d,%d,%d]", N,
               C, H, W);
  cudaStream_t st = (cudaStream_t)stream;
  TATT_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * N, st));
  dim3 grid((W + SS_TW - 1) / SS_TW, (H + SS_TH - 1) / SS_TH, N * C);
  tri_ssim_fwd_kernel<<<grid, 256, 0, st>>>(x1, x2, x3, G, C, H, W, (long long)N * C * H * W, make_gauss11(), (double*)ws);
  TATT_LAUNCH_CHECK("tri_ssim_fwd_kernel");
  tri_ssim_finalize_kernel<<<(N + 127) / 128, 128, 0, st>>>((const double*)ws, out, N, 1.0 / ((double)C * H * W), per_sample);
  TATT_LAUNCH_CHECK("tri_ssim_finalize_kernel");
  return 0;
}

int tatt_tri_ssim_bwd(const float* x1, const float* x2, const float* x3, const float* G, const float* gout, float* d1,
                      float* d2, float* d3, int N, int C, int H, int W, int per_sample, void* stream) {
  TATT_REQUIRE(N >= 1 && C >= 1 && H >= 1 && W >= 1 && (long long)N * C <= 65535, "tri_ssim_bwd: bad shape [%d,%d,%d,%d]", N,
               C, H, W);
  TATT_REQUIRE(G && gout && (d1 || d2 || d3), "tri_ssim_bwd: missing argument");
  dim3 grid((W + SS_TW - 1) / SS_TW, (H + SS_TH - 1) / SS_TH, N * C);
  const float scale = (float)(per_sample ? 1.0 / ((double)C * H * W) : 1.0 / ((double)N * C * H * W));
  tri_ssim_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x1, x2, x3, G, gout, per_sample, scale, d1, d2, d3, C, H, W,
                                                              (long long)N * C * H * W, make_gauss11());
  TATT_LAUNCH_CHECK("tri_ssim_bwd_kernel");
  return 0;
}

int tatt_rotate_img_fwd(const float* img, const float* arcs, const float* offs, float off_range, float* out, int N, int C,
                        int H, int W, void* stream) {
  TATT_REQUIRE(N >= 1 && N <= 65535 && C >= 1 && H >= 1 && W >= 1, "rotate_img_fwd: bad shape [%d,%d,%d,%d]", N, C, H, W);
  dim3 grid((H * W + 255) / 256, N);
  rotate_img_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(img, arcs, offs, off_range, out, C, H, W);
  TATT_LAUNCH_CHECK("rotate_img_fwd_kernel");
  return 0;
}

int tatt_rotate_img_bwd(const float* dout, const float* arcs, const float* offs, float off_range, float* dimg, int N, int C,
                        int H, int W, void* stream) {
  TATT_REQUIRE(N >= 1 && N <= 65535 && C >= 1 && H >= 1 && W >= 1, "rotate_img_bwd: bad shape [%d,%d,%d,%d]", N, C, H, W);
  cudaStream_t st = (cudaStream_t)stream;
  TATT_CUDA(cudaMemsetAsync(dimg, 0, sizeof(float) * (size_t)N * C * H * W, st));
  dim3 grid((H * W + 255) / 256, N);
  rotate_img_bwd_kernel<<<grid, 256, 0, st>>>(dout, arcs, offs, off_range, dimg, C, H, W);
  TATT_LAUNCH_CHECK("rotate_img_bwd_kernel");
  return 0;
}

}  // extern "C"
