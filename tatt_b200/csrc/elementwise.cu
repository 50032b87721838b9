// Element-wise / layout kernels (all HBM-bound; vectorised where the layout allows).
// Reference ops replaced: nn.PReLU (tsrn.py:598, 173), mish (tsrn.py:1061-1064), nn.PixelShuffle
// (tsrn.py:1046), torch.tanh (tsrn.py:675), nn.Dropout (transformer_v2.py:29,455,461-462,788,795-797),
// the NCHW<->NHWC permutes around GruBlock (tsrn.py:1076-1083) and the residual adds.
#include <cuda_bf16.h>
#include "common.cuh"

namespace {

static int ew_blocks(long long n, int per = 256) {
  long long b = (n + per - 1) / per;
  if (b > 148LL * 16) b = 148LL * 16;
  if (b < 1) b = 1;
  return (int)b;
}

#define GRID_STRIDE(i, n) \
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (n); i += (long long)gridDim.x * blockDim.x)

__global__ void axpby_kernel(const float* __restrict__ a, const float* __restrict__ b, float alpha, float beta,
                             float* __restrict__ out, long long n) {
  GRID_STRIDE(i, n) out[i] = alpha * a[i] + (b ? beta * b[i] : 0.f);
}
__global__ void axpby4_kernel(const float4* __restrict__ a, const float4* __restrict__ b, float alpha, float beta,
                              float4* __restrict__ out, long long n4) {
  GRID_STRIDE(i, n4) {
    float4 x = a[i], y = b[i];
    out[i] = make_float4(alpha * x.x + beta * y.x, alpha * x.y + beta * y.y, alpha * x.z + beta * y.z,
                         alpha * x.w + beta * y.w);
  }
}

// out[r][c] = a[r][c] + b[r % period][c]   (broadcast add over the leading axis; rows of `cols`)
__global__ void add_bcast_rows_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                      float* __restrict__ out, long long rows, long long period, int cols) {
  long long n = rows * cols;
  GRID_STRIDE(i, n) {
    long long r = i / cols;
    int c = (int)(i - r * cols);
    out[i] = (a ? a[i] : 0.f) + b[(r % period) * cols + c];
  }
}

__global__ void prelu_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y,
                                 long long n) {
  const float a = w[0];
  GRID_STRIDE(i, n) {
    float v = x[i];
    y[i] = v >= 0.f ? v : a * v;
  }
}
// dx = dy * (x>=0 ? 1 : a) ; dw += sum dy * x * [x<0]   (torch: x > 0 ? 1 : a; equal at 0 up to dw term = 0)
__global__ void prelu_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                 const float* __restrict__ dy, float* __restrict__ dx, float* __restrict__ dw,
                                 long long n) {
  __shared__ float red[8];
  const float a = w[0];
  float s = 0.f;
  GRID_STRIDE(i, n) {
    float v = x[i], g = dy[i];
    if (v > 0.f) {
      if (dx) dx[i] = g;
    } else {
      if (dx) dx[i] = a * g;
      s = fmaf(g, v, s);
    }
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    atomicAdd(dw, t);
  }
}

// in [N][H][W][4*C] -> out [N][2H][2W][C], out[n,2h+i,2w+j,c] = mish(in[n,h,w,4c+2i+j])
// One thread per input float4 = the four sub-pixels (i,j) of channel c: the read is a coalesced float4, the four scalar
// writes of C consecutive threads are four contiguous C-float runs (one per output pixel).
__global__ void pixshuf_mish_fwd_kernel(const float* __restrict__ in, float* __restrict__ out, long long npix_in,
                                        int H, int W, int C) {
  const long long n = npix_in * C;
  const long long orow = 2LL * W * C;          // one output row
  GRID_STRIDE(i, n) {
    const int c = (int)(i % C);
    const long long p = i / C;                 // input pixel (img, h, w)
    const int w = (int)(p % W);
    const long long q = p / W;                 // img * H + h
    const float4 v = __ldg(reinterpret_cast<const float4*>(in) + i);
    float* o = out + (2 * q) * orow + (2LL * w) * C + c;
    o[0] = mish_f(v.x);
    o[C] = mish_f(v.y);
    o[orow] = mish_f(v.z);
    o[orow + C] = mish_f(v.w);
  }
}
// the same map, written as bf16 hi / lo planes [N*2H*2W][C] (X operand of the convolution that follows the up-sampler)
__global__ void pixshuf_mish_planes_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ hi,
                                           __nv_bfloat16* __restrict__ lo, long long npix_in, int H, int W, int C) {
  const long long n = npix_in * C;
  const long long orow = 2LL * W * C;
  GRID_STRIDE(i, n) {
    const int c = (int)(i % C);
    const long long p = i / C;
    const int w = (int)(p % W);
    const long long q = p / W;
    const float4 v = __ldg(reinterpret_cast<const float4*>(in) + i);
    const long long o = (2 * q) * orow + (2LL * w) * C + c;
    const float y[4] = {mish_f(v.x), mish_f(v.y), mish_f(v.z), mish_f(v.w)};
    const long long off[4] = {0, C, orow, orow + C};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const __nv_bfloat16 h = __float2bfloat16_rn(y[k]);
      hi[o + off[k]] = h;
      lo[o + off[k]] = __float2bfloat16_rn(y[k] - __bfloat162float(h));
    }
  }
}
__global__ void pixshuf_mish_bwd_kernel(const float* __restrict__ in, const float* __restrict__ dout,
                                        float* __restrict__ din, long long npix_in, int H, int W, int C) {
  const long long n = npix_in * C;
  const long long orow = 2LL * W * C;
  GRID_STRIDE(i, n) {
    const int c = (int)(i % C);
    const long long p = i / C;
    const int w = (int)(p % W);
    const long long q = p / W;
    const float4 v = __ldg(reinterpret_cast<const float4*>(in) + i);
    const float* g = dout + (2 * q) * orow + (2LL * w) * C + c;
    float4 r;
    r.x = __ldg(g) * mish_grad(v.x);
    r.y = __ldg(g + C) * mish_grad(v.y);
    r.z = __ldg(g + orow) * mish_grad(v.z);
    r.w = __ldg(g + orow + C) * mish_grad(v.w);
    reinterpret_cast<float4*>(din)[i] = r;
  }
}

// NCHW (C real channels) -> NHWC with Cp >= C channels (zero padded)
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out, long long N, int C,
                                    int HW, int Cp) {
  long long n = N * HW * Cp;
  GRID_STRIDE(i, n) {
    int c = (int)(i % Cp);
    long long p = i / Cp;
    int hw = (int)(p % HW);
    long long img = p / HW;
    out[i] = c < C ? in[(img * C + c) * HW + hw] : 0.f;
  }
}
// NHWC (Cp channels) -> NCHW keeping the first C channels; optional tanh
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ in, float* __restrict__ out, long long N, int C,
                                    int HW, int Cp, int do_tanh) {
  long long n = N * C * HW;
  GRID_STRIDE(i, n) {
    int hw = (int)(i % HW);
    long long r = i / HW;
    int c = (int)(r % C);
    long long img = r / C;
    float v = in[(img * HW + hw) * Cp + c];
    out[i] = do_tanh ? tanhf(v) : v;
  }
}
// d(pre)[NHWC, Cp] = dout[NCHW, C] * (1 - out^2)   (padded channels get 0)
__global__ void tanh_bwd_nchw_to_nhwc_kernel(const float* __restrict__ dout, const float* __restrict__ out,
                                             float* __restrict__ dpre, long long N, int C, int HW, int Cp) {
  long long n = N * HW * Cp;
  GRID_STRIDE(i, n) {
    int c = (int)(i % Cp);
    long long p = i / Cp;
    int hw = (int)(p % HW);
    long long img = p / HW;
    float v = 0.f;
    if (c < C) {
      long long s = (img * C + c) * HW + hw;
      float o = out ? out[s] : 0.f;
      v = dout[s] * (1.f - o * o);
    }
    dpre[i] = v;
  }
}

// Mask of element e: 16-bit lane (e & 7) of Philox4x32-10(counter e >> 3) >= round(p * 65536)  (8 elements per Philox
// call; the keep probability is quantised to 1/65536 like the attention dropout of attn.cu).  The fused decoder-layer
// kernel (tc6_declayer.cu:row_dropout) draws the same bits.
__global__ void dropout_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, float p,
                               const unsigned long long* __restrict__ rng, unsigned long long site, int vec) {
  const float scale = 1.f / (1.f - p);
  const uint32_t thr = (uint32_t)(p * 65536.f + 0.5f);
  const unsigned long long seed = rng[0], offset = rng[1] * 65536ull + site;
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  const long long n8 = (n + 7) >> 3;
  GRID_STRIDE(i, n8) {
    const uint4 r = philox4x32_10(key, make_uint4((uint32_t)i, (uint32_t)((unsigned long long)i >> 32), (uint32_t)offset,
                                                  (uint32_t)(offset >> 32)));
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
    float m[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      m[2 * k] = (w[k] & 0xffffu) >= thr ? scale : 0.f;
      m[2 * k + 1] = (w[k] >> 16) >= thr ? scale : 0.f;
    }
    const long long b = i << 3;
    if (vec && b + 8 <= n) {
      const float4 a = *reinterpret_cast<const float4*>(x + b), c = *reinterpret_cast<const float4*>(x + b + 4);
      *reinterpret_cast<float4*>(y + b) = make_float4(a.x * m[0], a.y * m[1], a.z * m[2], a.w * m[3]);
      *reinterpret_cast<float4*>(y + b + 4) = make_float4(c.x * m[4], c.y * m[5], c.z * m[6], c.w * m[7]);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (b + k < n) y[b + k] = x[b + k] * m[k];
    }
  }
}

__global__ void relu_mask_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy,
                                     float* __restrict__ dx, long long n) {
  GRID_STRIDE(i, n) dx[i] = y[i] > 0.f ? dy[i] : 0.f;
}

// 2x2 / 1x2 max pooling, NHWC, stride == kernel
__global__ void maxpool_fwd_kernel(const float* __restrict__ in, float* __restrict__ out, long long N, int H,
                                   int W, int C, int kh, int kw) {
  int Ho = H / kh, Wo = W / kw;
  long long n = N * Ho * Wo * C;
  GRID_STRIDE(i, n) {
    int c = (int)(i % C);
    long long p = i / C;
    int ox = (int)(p % Wo);
    long long q = p / Wo;
    int oy = (int)(q % Ho);
    long long img = q / Ho;
    float m = -INFINITY;
    for (int a = 0; a < kh; ++a)
      for (int b = 0; b < kw; ++b) {
        float v = in[((img * H + oy * kh + a) * W + ox * kw + b) * C + c];
        m = v > m ? v : m;
      }
    out[i] = m;
  }
}
__global__ void maxpool_bwd_kernel(const float* __restrict__ in, const float* __restrict__ dout,
                                   float* __restrict__ din, long long N, int H, int W, int C, int kh, int kw) {
  int Ho = H / kh, Wo = W / kw;
  long long n = N * Ho * Wo * C;
  GRID_STRIDE(i, n) {
    int c = (int)(i % C);
    long long p = i / C;
    int ox = (int)(p % Wo);
    long long q = p / Wo;
    int oy = (int)(q % Ho);
    long long img = q / Ho;
    float m = -INFINITY;
    int am = 0;
    for (int a = 0; a < kh; ++a)
      for (int b = 0; b < kw; ++b) {
        float v = in[((img * H + oy * kh + a) * W + ox * kw + b) * C + c];
        if (v > m) {
          m = v;
          am = a * kw + b;
        }
      }
    float g = dout[i];
    for (int a = 0; a < kh; ++a)
      for (int b = 0; b < kw; ++b)
        din[((img * H + oy * kh + a) * W + ox * kw + b) * C + c] = (a * kw + b == am) ? g : 0.f;
  }
}

__global__ void sqnorm_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
  __shared__ float red[8];
  float s = 0.f;
  GRID_STRIDE(i, n) s = fmaf(x[i], x[i], s);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    atomicAdd(out, t);
  }
}

// Deterministic variant (fixed reduction order, no atomics): data-parallel replicas must derive the SAME clip coefficient
// from the same all-reduced gradient, bit for bit, or they drift apart by an ulp per step.
__global__ void sqnorm_partial_kernel(const float* __restrict__ x, long long n, float* __restrict__ part) {
  __shared__ float red[8];
  float s = 0.f;
  GRID_STRIDE(i, n) s = fmaf(x[i], x[i], s);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    part[blockIdx.x] = t;
  }
}
__global__ void sqnorm_final_kernel(const float* __restrict__ part, int nparts, float* __restrict__ out) {
  __shared__ float red[8];
  float s = 0.f;
  for (int i = threadIdx.x; i < nparts; i += 256) s += part[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    out[0] = t;
  }
}

// Gradient packing as ONE launch: chunk c copies tab[3c+2] floats from address tab[3c] to dst + tab[3c+1]
// (chunks are <= 16384 floats; 16-byte aligned sources take the float4 path) and, when sq != nullptr, adds the
// chunk's sum of squares to sq[0] (the global-norm numerator of clip_grad_norm_ for a single-rank step).
__global__ void multi_copy_kernel(const unsigned long long* __restrict__ tab, float* __restrict__ dst,
                                  float* __restrict__ sq) {
  __shared__ float red[8];
  const unsigned long long* e = tab + 3ull * blockIdx.x;
  const float* __restrict__ src = (const float*)e[0];
  float* __restrict__ out = dst + e[1];
  const int n = (int)e[2];
  float s = 0.f;
  if (((e[0] | (unsigned long long)(uintptr_t)out) & 15ull) == 0) {
    const int n4 = n >> 2;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
      const float4 v = ((const float4*)src)[i];
      ((float4*)out)[i] = v;
      s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    for (int i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) {
      const float v = src[i];
      out[i] = v;
      s = fmaf(v, v, s);
    }
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const float v = src[i];
      out[i] = v;
      s = fmaf(v, v, s);
    }
  }
  if (sq == nullptr) return;
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    atomicAdd(sq, t);
  }
}

// Fused global-norm clip + Adam on a flat fp32 buffer (super_resolution.py:1083-1085, base.py:557-558).
// sqnorm[0] holds sum(g^2) over the whole buffer (already all-reduced and averaged grads).
__global__ void adam_clip_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, long long n, const float* __restrict__ sqnorm,
                                 float max_norm, float lr, float b1, float b2, float eps,
                                 const unsigned long long* __restrict__ step_state, float grad_scale) {
  // step counter lives in device memory (CUDA-graph replays must see it advance)
  const float stepf = (float)step_state[1];
  const float bc1 = 1.f - powf(b1, stepf), bc2 = 1.f - powf(b2, stepf);
  float coef = 1.f;
  if (max_norm > 0.f) {
    float tn = sqrtf(sqnorm[0]) * grad_scale;
    coef = fminf(max_norm / (tn + 1e-6f), 1.f);
  }
  coef *= grad_scale;
  GRID_STRIDE(i, n) {
    float gi = g[i] * coef;
    float mi = b1 * m[i] + (1.f - b1) * gi;
    float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    float denom = sqrtf(vi) / sqrtf(bc2) + eps;
    p[i] -= (lr / bc1) * (mi / denom);
  }
}

}  // namespace

extern "C" {

// out = alpha*a + beta*b   (b may be NULL)
int tatt_axpby(const float* a, const float* b, float alpha, float beta, float* out, long long n, void* stream) {
  if (n <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  bool v4 = b && (n % 4 == 0) && ((((uintptr_t)a) | ((uintptr_t)b) | ((uintptr_t)out)) & 15) == 0;
  if (v4)
    axpby4_kernel<<<ew_blocks(n / 4), 256, 0, st>>>((const float4*)a, (const float4*)b, alpha, beta, (float4*)out,
                                                    n / 4);
  else
    axpby_kernel<<<ew_blocks(n), 256, 0, st>>>(a, b, alpha, beta, out, n);
  TATT_LAUNCH_CHECK("axpby_kernel");
  return 0;
}

int tatt_add_bcast_rows(const float* a, const float* b, float* out, long long rows, long long period, int cols,
                        void* stream) {
  if (rows <= 0) return 0;
  add_bcast_rows_kernel<<<ew_blocks(rows * cols), 256, 0, (cudaStream_t)stream>>>(a, b, out, rows, period, cols);
  TATT_LAUNCH_CHECK("add_bcast_rows_kernel");
  return 0;
}

int tatt_prelu_fwd(const float* x, const float* w, float* y, long long n, void* stream) {
  if (n <= 0) return 0;
  prelu_fwd_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(x, w, y, n);
  TATT_LAUNCH_CHECK("prelu_fwd_kernel");
  return 0;
}
// dw (1 element) is zeroed here and then accumulated; dx may be NULL.
int tatt_prelu_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, long long n,
                   void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  TATT_CUDA(cudaMemsetAsync(dw, 0, sizeof(float), st));
  if (n <= 0) return 0;
  prelu_bwd_kernel<<<ew_blocks(n, 1024), 256, 0, st>>>(x, w, dy, dx, dw, n);
  TATT_LAUNCH_CHECK("prelu_bwd_kernel");
  return 0;
}

int tatt_pixshuf2_mish_fwd(const float* in, float* out, long long nimg, int H, int W, int C, void* stream) {
  long long npix_in = nimg * (long long)H * W;
  if (npix_in <= 0) return 0;
  pixshuf_mish_fwd_kernel<<<ew_blocks(npix_in * C), 256, 0, (cudaStream_t)stream>>>(in, out, npix_in, H, W, C);
  TATT_LAUNCH_CHECK("pixshuf_mish_fwd_kernel");
  return 0;
}
int tatt_pixshuf2_mish_planes(const float* in, void* out_hi, void* out_lo, long long nimg, int H, int W, int C,
                              void* stream) {
  long long npix_in = nimg * (long long)H * W;
  if (npix_in <= 0) return 0;
  TATT_REQUIRE(out_hi && out_lo, "pixshuf2_mish_planes: null planes");
  pixshuf_mish_planes_kernel<<<ew_blocks(npix_in * C), 256, 0, (cudaStream_t)stream>>>(
      in, reinterpret_cast<__nv_bfloat16*>(out_hi), reinterpret_cast<__nv_bfloat16*>(out_lo), npix_in, H, W, C);
  TATT_LAUNCH_CHECK("pixshuf_mish_planes_kernel");
  return 0;
}
int tatt_pixshuf2_mish_bwd(const float* in, const float* dout, float* din, long long nimg, int H, int W, int C,
                           void* stream) {
  long long npix_in = nimg * (long long)H * W;
  if (npix_in <= 0) return 0;
  pixshuf_mish_bwd_kernel<<<ew_blocks(npix_in * C), 256, 0, (cudaStream_t)stream>>>(in, dout, din, npix_in, H, W, C);
  TATT_LAUNCH_CHECK("pixshuf_mish_bwd_kernel");
  return 0;
}

int tatt_nchw_to_nhwc(const float* in, float* out, long long N, int C, int H, int W, int Cp, void* stream) {
  long long n = N * H * W * Cp;
  if (n <= 0) return 0;
  nchw_to_nhwc_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(in, out, N, C, H * W, Cp);
  TATT_LAUNCH_CHECK("nchw_to_nhwc_kernel");
  return 0;
}
int tatt_nhwc_to_nchw(const float* in, float* out, long long N, int C, int H, int W, int Cp, int do_tanh,
                      void* stream) {
  long long n = N * C * H * W;
  if (n <= 0) return 0;
  nhwc_to_nchw_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(in, out, N, C, H * W, Cp, do_tanh);
  TATT_LAUNCH_CHECK("nhwc_to_nchw_kernel");
  return 0;
}
int tatt_tanh_bwd_nchw_to_nhwc(const float* dout, const float* out, float* dpre, long long N, int C, int H, int W,
                               int Cp, void* stream) {
  long long n = N * H * W * Cp;
  if (n <= 0) return 0;
  tanh_bwd_nchw_to_nhwc_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(dout, out, dpre, N, C, H * W, Cp);
  TATT_LAUNCH_CHECK("tanh_bwd_nchw_to_nhwc_kernel");
  return 0;
}

__global__ void rng_advance_kernel(unsigned long long* rng) { rng[1] += 1ull; }

// y = x * keep / (1-p); the same call on dy (same rng snapshot / site) is the backward.
// rng: device pointer to {seed, counter}; the mask of element i is Philox(seed; i, counter*65536+site).
int tatt_dropout(const float* x, float* y, long long n, float p, const unsigned long long* rng,
                 unsigned long long site, void* stream) {
  TATT_REQUIRE(p >= 0.f && p < 1.f, "dropout: p must be in [0,1)");
  if (n <= 0) return 0;
  const int vec = ((((uintptr_t)x) | ((uintptr_t)y)) & 15) == 0 ? 1 : 0;
  dropout_kernel<<<ew_blocks((n + 7) / 8), 256, 0, (cudaStream_t)stream>>>(x, y, n, p, rng, site, vec);
  TATT_LAUNCH_CHECK("dropout_kernel");
  return 0;
}
int tatt_rng_advance(unsigned long long* rng, void* stream) {
  rng_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(rng);
  TATT_LAUNCH_CHECK("rng_advance_kernel");
  return 0;
}

int tatt_relu_bwd(const float* y, const float* dy, float* dx, long long n, void* stream) {
  if (n <= 0) return 0;
  relu_mask_bwd_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(y, dy, dx, n);
  TATT_LAUNCH_CHECK("relu_mask_bwd_kernel");
  return 0;
}

int tatt_maxpool_fwd(const float* in, float* out, long long N, int H, int W, int C, int kh, int kw, void* stream) {
  long long n = N * (H / kh) * (W / kw) * C;
  if (n <= 0) return 0;
  maxpool_fwd_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(in, out, N, H, W, C, kh, kw);
  TATT_LAUNCH_CHECK("maxpool_fwd_kernel");
  return 0;
}
int tatt_maxpool_bwd(const float* in, const float* dout, float* din, long long N, int H, int W, int C, int kh,
                     int kw, void* stream) {
  TATT_REQUIRE(H % kh == 0 && W % kw == 0, "maxpool_bwd: H,W must be divisible by the window");
  long long n = N * (H / kh) * (W / kw) * C;
  if (n <= 0) return 0;
  maxpool_bwd_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(in, dout, din, N, H, W, C, kh, kw);
  TATT_LAUNCH_CHECK("maxpool_bwd_kernel");
  return 0;
}

// out[0] (+)= sum x^2
int tatt_sqnorm(const float* x, long long n, float* out, int zero_first, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (zero_first) TATT_CUDA(cudaMemsetAsync(out, 0, sizeof(float), st));
  if (n <= 0) return 0;
  sqnorm_kernel<<<ew_blocks(n, 2048), 256, 0, st>>>(x, n, out);
  TATT_LAUNCH_CHECK("sqnorm_kernel");
  return 0;
}

int tatt_sqnorm_det(const float* x, long long n, float* out, float* ws, int ws_floats, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  TATT_REQUIRE(ws != nullptr && ws_floats >= 1, "sqnorm_det: needs a scratch buffer");
  if (n <= 0) {
    TATT_CUDA(cudaMemsetAsync(out, 0, sizeof(float), st));
    return 0;
  }
  int nb = ew_blocks(n, 2048);
  if (nb > ws_floats) nb = ws_floats;
  sqnorm_partial_kernel<<<nb, 256, 0, st>>>(x, n, ws);
  TATT_LAUNCH_CHECK("sqnorm_partial_kernel");
  sqnorm_final_kernel<<<1, 256, 0, st>>>(ws, nb, out);
  TATT_LAUNCH_CHECK("sqnorm_final_kernel");
  return 0;
}

int tatt_multi_copy(const unsigned long long* table, int nchunks, float* dst, float* sq, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (sq != nullptr) TATT_CUDA(cudaMemsetAsync(sq, 0, sizeof(float), st));
  if (nchunks <= 0) return 0;
  TATT_REQUIRE(table != nullptr && dst != nullptr, "multi_copy: null table / destination");
  multi_copy_kernel<<<nchunks, 256, 0, st>>>(table, dst, sq);
  TATT_LAUNCH_CHECK("multi_copy_kernel");
  return 0;
}

int tatt_adam_clip_step(float* p, const float* g, float* m, float* v, long long n, const float* sqnorm,
                        float max_norm, float lr, float beta1, float beta2, float eps,
                        const unsigned long long* step_state, float grad_scale, void* stream) {
  if (n <= 0) return 0;
  TATT_REQUIRE(step_state != nullptr, "adam_clip_step: step_state (device {unused, step}) is required");
  adam_clip_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, sqnorm, max_norm, lr, beta1,
                                                                   beta2, eps, step_state, grad_scale);
  TATT_LAUNCH_CHECK("adam_clip_kernel");
  return 0;
}

}  // extern "C"
