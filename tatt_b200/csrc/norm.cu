// BatchNorm (channels-last, batch statistics over all rows) and LayerNorm(64) kernels, fwd + bwd.
// Reference ops replaced: nn.BatchNorm2d/1d (model/tsrn.py:878,886,612; model/stn_head.py:19,51)
// and nn.LayerNorm(64) (model/transformer_v2.py:460-461, 792-794, 166).
#include <cuda_bf16.h>
#include "common.cuh"

namespace {

// ---------------------------------------------------------------- BatchNorm statistics
// X[P][C]; per-channel sum and sum of squares -> double accumulators acc[2*C].
__global__ void bn_stats_kernel(const float* __restrict__ X, long long P, int C, int rows_per_cta,
                                double* __restrict__ acc) {
  __shared__ float red[2][4][64];
  const int lane = threadIdx.x & 63;
  const int ty = threadIdx.x >> 6;
  const int c = blockIdx.y * 64 + lane;
  long long r0 = (long long)blockIdx.x * rows_per_cta;
  long long r1 = r0 + rows_per_cta;
  if (r1 > P) r1 = P;
  float s = 0.f, q = 0.f;
  if (c < C) {
    for (long long r = r0 + ty; r < r1; r += 4) {
      float v = X[r * C + c];
      s += v;
      q = fmaf(v, v, q);
    }
  }
  red[0][ty][lane] = s;
  red[1][ty][lane] = q;
  __syncthreads();
  if (ty == 0 && c < C) {
    float ts = red[0][0][lane] + red[0][1][lane] + red[0][2][lane] + red[0][3][lane];
    float tq = red[1][0][lane] + red[1][1][lane] + red[1][2][lane] + red[1][3][lane];
    atomicAdd(acc + c, (double)ts);
    atomicAdd(acc + C + c, (double)tq);
  }
}

// C == 64 fast paths (every BatchNorm2d of the trunk, tsrn.py:878,886,612): float4 loads (16 threads per 256-byte row,
// 16 rows per pass, 4 passes in flight) instead of one scalar per thread -- the scalar kernels reach ~2 TB/s, these are
// bound by HBM.  Same accumulation scheme: fp32 partials per thread / CTA, fp64 atomics across CTAs.
__global__ void __launch_bounds__(256) bn_stats64_kernel(const float4* __restrict__ X, long long P, int rows_per_cta,
                                                        double* __restrict__ acc) {
  __shared__ float red[2][16][64];
  const int c4 = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const long long r0 = (long long)blockIdx.x * rows_per_cta;
  long long r1 = r0 + rows_per_cta;
  if (r1 > P) r1 = P;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
#pragma unroll 4
  for (long long r = r0 + ty; r < r1; r += 16) {
    const float4 v = __ldg(X + r * 16 + c4);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    q.x = fmaf(v.x, v.x, q.x); q.y = fmaf(v.y, v.y, q.y); q.z = fmaf(v.z, v.z, q.z); q.w = fmaf(v.w, v.w, q.w);
  }
  *reinterpret_cast<float4*>(&red[0][ty][c4 * 4]) = s;
  *reinterpret_cast<float4*>(&red[1][ty][c4 * 4]) = q;
  __syncthreads();
  if (threadIdx.x < 128) {
    const int which = threadIdx.x >> 6, c = threadIdx.x & 63;
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) t += red[which][k][c];
    atomicAdd(acc + which * 64 + c, (double)t);
  }
}
__global__ void __launch_bounds__(256)
bn_bwd_reduce64_kernel(const float4* __restrict__ X, const float4* __restrict__ dY, const float* __restrict__ mean,
                       const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                       int act, long long P, int rows_per_cta, double* __restrict__ acc) {
  __shared__ float red[2][16][64];
  const int c4 = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const long long r0 = (long long)blockIdx.x * rows_per_cta;
  long long r1 = r0 + rows_per_cta;
  if (r1 > P) r1 = P;
  const float4 m = __ldg(reinterpret_cast<const float4*>(mean) + c4), is = __ldg(reinterpret_cast<const float4*>(invstd) + c4);
  const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4), b = __ldg(reinterpret_cast<const float4*>(beta) + c4);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
#pragma unroll 4
  for (long long r = r0 + ty; r < r1; r += 16) {
    const float4 x = __ldg(X + r * 16 + c4);
    float4 dz = __ldg(dY + r * 16 + c4);
    const float4 xh = make_float4((x.x - m.x) * is.x, (x.y - m.y) * is.y, (x.z - m.z) * is.z, (x.w - m.w) * is.w);
    if (act != ACT_NONE) {
      dz.x *= act_grad(xh.x * g.x + b.x, act);
      dz.y *= act_grad(xh.y * g.y + b.y, act);
      dz.z *= act_grad(xh.z * g.z + b.z, act);
      dz.w *= act_grad(xh.w * g.w + b.w, act);
    }
    s.x += dz.x; s.y += dz.y; s.z += dz.z; s.w += dz.w;
    q.x = fmaf(dz.x, xh.x, q.x); q.y = fmaf(dz.y, xh.y, q.y); q.z = fmaf(dz.z, xh.z, q.z); q.w = fmaf(dz.w, xh.w, q.w);
  }
  *reinterpret_cast<float4*>(&red[0][ty][c4 * 4]) = s;
  *reinterpret_cast<float4*>(&red[1][ty][c4 * 4]) = q;
  __syncthreads();
  if (threadIdx.x < 128) {
    const int which = threadIdx.x >> 6, c = threadIdx.x & 63;
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) t += red[which][k][c];
    atomicAdd(acc + which * 64 + c, (double)t);
  }
}
static inline bool bn64_ok(const void* a, const void* b, int C, long long P) {
  return C == 64 && P >= 4096 && (((uintptr_t)a | (uintptr_t)b) & 15) == 0;
}
static inline int bn64_rows(long long P) {
  long long rows = (P + 148 * 8 - 1) / (148 * 8);
  if (rows < 64) rows = 64;
  return (int)rows;
}

__global__ void bn_finalize_kernel(const double* __restrict__ acc, long long P, int C, float eps, float momentum,
                                   float* __restrict__ mean, float* __restrict__ invstd,
                                   float* __restrict__ running_mean, float* __restrict__ running_var) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double m = acc[c] / (double)P;
  double var = acc[C + c] / (double)P - m * m;
  if (var < 0.0) var = 0.0;
  mean[c] = (float)m;
  invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) {
    double unb = P > 1 ? var * (double)P / (double)(P - 1) : var;
    running_mean[c] = (float)((1.0 - momentum) * (double)running_mean[c] + momentum * m);
    running_var[c] = (float)((1.0 - momentum) * (double)running_var[c] + momentum * unb);
  }
}

__global__ void bn_eval_stats_kernel(const float* __restrict__ rm, const float* __restrict__ rv, float eps, int C,
                                     float* __restrict__ mean, float* __restrict__ invstd) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  mean[c] = rm[c];
  invstd[c] = 1.f / sqrtf(rv[c] + eps);
}

// Y = act(X * scale + shift)   (C % 4 == 0)
__global__ void bn_apply_kernel(const float* __restrict__ X, float* __restrict__ Y, const float* __restrict__ mean,
                                const float* __restrict__ invstd, const float* __restrict__ gamma,
                                const float* __restrict__ beta, int act, long long total4, int C4,
                                __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C4) * 4;
    float4 x = reinterpret_cast<const float4*>(X)[i];
    float4 m = *reinterpret_cast<const float4*>(mean + c);
    float4 is = *reinterpret_cast<const float4*>(invstd + c);
    float4 g = *reinterpret_cast<const float4*>(gamma + c);
    float4 b = *reinterpret_cast<const float4*>(beta + c);
    float4 y;
    y.x = act_fwd((x.x - m.x) * is.x * g.x + b.x, act);
    y.y = act_fwd((x.y - m.y) * is.y * g.y + b.y, act);
    y.z = act_fwd((x.z - m.z) * is.z * g.z + b.z, act);
    y.w = act_fwd((x.w - m.w) * is.w * g.w + b.w, act);
    if (hi) {           // only consumer is a convolution: bf16 hi / lo operand planes instead of the fp32 tensor
      const __nv_bfloat162 h0 = __floats2bfloat162_rn(y.x, y.y), h1 = __floats2bfloat162_rn(y.z, y.w);
      const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
      const __nv_bfloat162 l0 = __floats2bfloat162_rn(y.x - f0.x, y.y - f0.y), l1 = __floats2bfloat162_rn(y.z - f1.x, y.w - f1.y);
      reinterpret_cast<uint2*>(hi)[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
      reinterpret_cast<uint2*>(lo)[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
    } else {
      reinterpret_cast<float4*>(Y)[i] = y;
    }
  }
}

// acc[c] += sum dz, acc[C+c] += sum dz * xhat   with dz = dY * act'(z)
__global__ void bn_bwd_reduce_kernel(const float* __restrict__ X, const float* __restrict__ dY,
                                     const float* __restrict__ mean, const float* __restrict__ invstd,
                                     const float* __restrict__ gamma, const float* __restrict__ beta, int act,
                                     long long P, int C, int rows_per_cta, double* __restrict__ acc) {
  __shared__ float red[2][4][64];
  const int lane = threadIdx.x & 63;
  const int ty = threadIdx.x >> 6;
  const int c = blockIdx.y * 64 + lane;
  long long r0 = (long long)blockIdx.x * rows_per_cta;
  long long r1 = r0 + rows_per_cta;
  if (r1 > P) r1 = P;
  float s = 0.f, q = 0.f;
  if (c < C) {
    float m = mean[c], is = invstd[c], g = gamma[c], b = beta[c];
    for (long long r = r0 + ty; r < r1; r += 4) {
      float xh = (X[r * C + c] - m) * is;
      float dz = dY[r * C + c];
      if (act != ACT_NONE) dz *= act_grad(xh * g + b, act);
      s += dz;
      q = fmaf(dz, xh, q);
    }
  }
  red[0][ty][lane] = s;
  red[1][ty][lane] = q;
  __syncthreads();
  if (ty == 0 && c < C) {
    float ts = red[0][0][lane] + red[0][1][lane] + red[0][2][lane] + red[0][3][lane];
    float tq = red[1][0][lane] + red[1][1][lane] + red[1][2][lane] + red[1][3][lane];
    atomicAdd(acc + c, (double)ts);
    atomicAdd(acc + C + c, (double)tq);
  }
}

__global__ void bn_bwd_finalize_kernel(const double* __restrict__ acc, int C, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  dbeta[c] = (float)acc[c];
  dgamma[c] = (float)acc[C + c];
}

__global__ void bn_bwd_apply_kernel(const float* __restrict__ X, const float* __restrict__ dY,
                                    const float* __restrict__ mean, const float* __restrict__ invstd,
                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                    const float* __restrict__ dgamma, const float* __restrict__ dbeta, int act,
                                    int training, float invP, long long total4, int C4, float* __restrict__ dX,
                                    __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    const float4 x = reinterpret_cast<const float4*>(X)[i];
    const float4 g4 = reinterpret_cast<const float4*>(dY)[i];
    const float4 m = *reinterpret_cast<const float4*>(mean + c);
    const float4 is = *reinterpret_cast<const float4*>(invstd + c);
    const float4 ga = *reinterpret_cast<const float4*>(gamma + c);
    const float4 be = *reinterpret_cast<const float4*>(beta + c);
    const float4 dg = *reinterpret_cast<const float4*>(dgamma + c);
    const float4 db = *reinterpret_cast<const float4*>(dbeta + c);
    const float xv[4] = {x.x, x.y, x.z, x.w}, gv[4] = {g4.x, g4.y, g4.z, g4.w};
    const float mv[4] = {m.x, m.y, m.z, m.w}, iv[4] = {is.x, is.y, is.z, is.w};
    const float gav[4] = {ga.x, ga.y, ga.z, ga.w}, bev[4] = {be.x, be.y, be.z, be.w};
    const float dgv[4] = {dg.x, dg.y, dg.z, dg.w}, dbv[4] = {db.x, db.y, db.z, db.w};
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float xh = (xv[j] - mv[j]) * iv[j];
      float dz = gv[j];
      if (act != ACT_NONE) dz *= act_grad(xh * gav[j] + bev[j], act);
      o[j] = training ? gav[j] * iv[j] * (dz - dbv[j] * invP - xh * dgv[j] * invP) : gav[j] * iv[j] * dz;
    }
    if (hi) {           // bf16 hi / lo operand planes of dX for the convolution in front (no fp32 copy, no split pass)
      const __nv_bfloat162 h0 = __floats2bfloat162_rn(o[0], o[1]), h1 = __floats2bfloat162_rn(o[2], o[3]);
      const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
      const __nv_bfloat162 l0 = __floats2bfloat162_rn(o[0] - f0.x, o[1] - f0.y), l1 = __floats2bfloat162_rn(o[2] - f1.x, o[3] - f1.y);
      reinterpret_cast<uint2*>(hi)[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
      reinterpret_cast<uint2*>(lo)[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
    } else {
      reinterpret_cast<float4*>(dX)[i] = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
}

// ---------------------------------------------------------------- LayerNorm over 64 channels
// one warp per row; lane owns channels lane and lane+32
__global__ void ln_fwd_kernel(const float* __restrict__ X, const float* __restrict__ R,
                              const float* __restrict__ gamma, const float* __restrict__ beta,
                              float* __restrict__ Y, float* __restrict__ S, float* __restrict__ mean,
                              float* __restrict__ rstd, long long P, float eps) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarp = ((long long)gridDim.x * blockDim.x) >> 5;
  const float g0 = gamma[lane], g1 = gamma[lane + 32], b0 = beta[lane], b1 = beta[lane + 32];
  for (long long r = warp; r < P; r += nwarp) {
    float v0 = X[r * 64 + lane], v1 = X[r * 64 + lane + 32];
    if (R) {
      v0 += R[r * 64 + lane];
      v1 += R[r * 64 + lane + 32];
    }
    if (S) {
      S[r * 64 + lane] = v0;
      S[r * 64 + lane + 32] = v1;
    }
    float m = warp_sum(v0 + v1) * (1.f / 64.f);
    float d0 = v0 - m, d1 = v1 - m;
    float var = warp_sum(d0 * d0 + d1 * d1) * (1.f / 64.f);
    float rs = 1.f / sqrtf(var + eps);
    Y[r * 64 + lane] = d0 * rs * g0 + b0;
    Y[r * 64 + lane + 32] = d1 * rs * g1 + b1;
    if (lane == 0 && mean) {
      mean[r] = m;
      rstd[r] = rs;
    }
  }
}

__global__ void ln_bwd_kernel(const float* __restrict__ dY, const float* __restrict__ S,
                              const float* __restrict__ mean, const float* __restrict__ rstd,
                              const float* __restrict__ gamma, float* __restrict__ dS,
                              float* __restrict__ dgamma, float* __restrict__ dbeta, long long P) {
  __shared__ float red[2][8][64];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarp = ((long long)gridDim.x * blockDim.x) >> 5;
  const float g0 = gamma[lane], g1 = gamma[lane + 32];
  float ag0 = 0.f, ag1 = 0.f, ab0 = 0.f, ab1 = 0.f;
  for (long long r = warp; r < P; r += nwarp) {
    float m = mean[r], rs = rstd[r];
    float xh0 = (S[r * 64 + lane] - m) * rs, xh1 = (S[r * 64 + lane + 32] - m) * rs;
    float dy0 = dY[r * 64 + lane], dy1 = dY[r * 64 + lane + 32];
    ag0 = fmaf(dy0, xh0, ag0);
    ag1 = fmaf(dy1, xh1, ag1);
    ab0 += dy0;
    ab1 += dy1;
    float w0 = dy0 * g0, w1 = dy1 * g1;
    float c1 = warp_sum(w0 + w1) * (1.f / 64.f);
    float c2 = warp_sum(w0 * xh0 + w1 * xh1) * (1.f / 64.f);
    dS[r * 64 + lane] = rs * (w0 - c1 - xh0 * c2);
    dS[r * 64 + lane + 32] = rs * (w1 - c1 - xh1 * c2);
  }
  red[0][wib][lane] = ag0;
  red[0][wib][lane + 32] = ag1;
  red[1][wib][lane] = ab0;
  red[1][wib][lane + 32] = ab1;
  __syncthreads();
  if (threadIdx.x < 128) {
    int which = threadIdx.x >> 6, c = threadIdx.x & 63;
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[which][w][c];
    atomicAdd((which == 0 ? dgamma : dbeta) + c, t);
  }
}

static int ew_blocks(long long n, int per = 256) {
  long long b = (n + per - 1) / per;
  if (b > 148LL * 16) b = 148LL * 16;
  if (b < 1) b = 1;
  return (int)b;
}

// mean / invstd (+ running-statistics update) from per-CTA partial {sum, sum of squares} rows produced by the convolution
// kernel's epilogue (tatt_conv3x3_stats): the rows are added in double in a fixed order
__global__ void bn_finalize_parts_kernel(const float* __restrict__ parts, int nparts, long long P, int C, float eps,
                                         float momentum, float* __restrict__ mean, float* __restrict__ invstd,
                                         float* __restrict__ running_mean, float* __restrict__ running_var) {
  // one warp per channel: lanes stride over the rows (a single thread walking the ~150 rows took 27 us on the critical
  // path of every convolution block), fixed-order shuffle reduction in double
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= C) return;
  double s = 0.0, q = 0.0;
  for (int i = lane; i < nparts; i += 32) {
    s += (double)parts[(long long)i * 2 * C + c];
    q += (double)parts[(long long)i * 2 * C + C + c];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if (lane != 0) return;
  const double m = s / (double)P;
  double var = q / (double)P - m * m;
  if (var < 0.0) var = 0.0;
  mean[c] = (float)m;
  invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) {
    const double unb = P > 1 ? var * (double)P / (double)(P - 1) : var;
    running_mean[c] = (float)((1.0 - momentum) * (double)running_mean[c] + momentum * m);
    running_var[c] = (float)((1.0 - momentum) * (double)running_var[c] + momentum * unb);
  }
}
}  // namespace

extern "C" {

// Batch statistics of X[P][C] (biased variance) -> mean/invstd; optional running-stat update with
// the unbiased variance (torch semantics).  ws: >= 2*C doubles of scratch.
int tatt_bn_stats(const float* X, long long P, int C, float eps, float momentum, float* mean, float* invstd,
                  float* running_mean, float* running_var, void* ws, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  TATT_REQUIRE(P >= 1 && C >= 1, "bn_stats: empty input");
  TATT_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 2 * C, st));
  if (bn64_ok(X, nullptr, C, P)) {
    const int rows = bn64_rows(P);
    bn_stats64_kernel<<<ceil_div(P, rows), 256, 0, st>>>(reinterpret_cast<const float4*>(X), P, rows, (double*)ws);
    TATT_LAUNCH_CHECK("bn_stats64_kernel");
  } else {
    int rows = 128;
    if (P > 128LL * 4096) rows = (int)((P + 4095) / 4096);
    dim3 grid(ceil_div(P, rows), ceil_div(C, 64));
    bn_stats_kernel<<<grid, 256, 0, st>>>(X, P, C, rows, (double*)ws);
    TATT_LAUNCH_CHECK("bn_stats_kernel");
  }
  bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, st>>>((const double*)ws, P, C, eps, momentum, mean, invstd,
                                                       running_mean, running_var);
  TATT_LAUNCH_CHECK("bn_finalize_kernel");
  return 0;
}

int tatt_bn_finalize(const float* parts, int nparts, long long P, int C, float eps, float momentum, float* mean,
                     float* invstd, float* running_mean, float* running_var, void* stream) {
  TATT_REQUIRE(P >= 1 && C >= 1 && nparts >= 1 && parts != nullptr, "bn_finalize: empty input");
  bn_finalize_parts_kernel<<<ceil_div(C, 8), 256, 0, (cudaStream_t)stream>>>(parts, nparts, P, C, eps, momentum, mean, invstd,
                                                                           running_mean, running_var);
  TATT_LAUNCH_CHECK("bn_finalize_parts_kernel");
  return 0;
}

int tatt_bn_eval_stats(const float* running_mean, const float* running_var, float eps, int C, float* mean,
                       float* invstd, void* stream) {
  bn_eval_stats_kernel<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(running_mean, running_var, eps, C,
                                                                            mean, invstd);
  TATT_LAUNCH_CHECK("bn_eval_stats_kernel");
  return 0;
}

int tatt_bn_apply_fwd(const float* X, float* Y, const float* mean, const float* invstd, const float* gamma,
                      const float* beta, int act, long long P, int C, void* stream) {
  TATT_REQUIRE(C % 4 == 0, "bn_apply: C must be a multiple of 4");
  long long total4 = P * C / 4;
  if (total4 == 0) return 0;
  bn_apply_kernel<<<ew_blocks(total4), 256, 0, (cudaStream_t)stream>>>(X, Y, mean, invstd, gamma, beta, act,
                                                                       total4, C / 4, nullptr, nullptr);
  TATT_LAUNCH_CHECK("bn_apply_kernel");
  return 0;
}
// the same, written only as bf16 hi / lo planes [P][C] (the X operand of the convolution that consumes it)
int tatt_bn_apply_planes(const float* X, void* y_hi, void* y_lo, const float* mean, const float* invstd, const float* gamma,
                         const float* beta, int act, long long P, int C, void* stream) {
  TATT_REQUIRE(C % 4 == 0 && y_hi && y_lo && ((((uintptr_t)y_hi) | ((uintptr_t)y_lo)) & 7) == 0,
               "bn_apply_planes: C %% 4 == 0 and 8-byte aligned planes required");
  long long total4 = P * C / 4;
  if (total4 == 0) return 0;
  bn_apply_kernel<<<ew_blocks(total4), 256, 0, (cudaStream_t)stream>>>(X, nullptr, mean, invstd, gamma, beta, act, total4,
                                                                       C / 4, reinterpret_cast<__nv_bfloat16*>(y_hi),
                                                                       reinterpret_cast<__nv_bfloat16*>(y_lo));
  TATT_LAUNCH_CHECK("bn_apply_kernel");
  return 0;
}

// dgamma/dbeta (with the activation's backward fused) then dX.  ws: >= 2*C doubles.
static int bn_bwd_impl(const float* X, const float* dY, const float* mean, const float* invstd, const float* gamma,
                       const float* beta, int act, int training, long long P, int C, float* dX, void* dx_hi, void* dx_lo,
                       float* dgamma, float* dbeta, void* ws, void* stream);
int tatt_bn_bwd(const float* X, const float* dY, const float* mean, const float* invstd, const float* gamma,
                const float* beta, int act, int training, long long P, int C, float* dX, float* dgamma,
                float* dbeta, void* ws, void* stream) {
  return bn_bwd_impl(X, dY, mean, invstd, gamma, beta, act, training, P, C, dX, nullptr, nullptr, dgamma, dbeta, ws, stream);
}
// The same backward, with dX written ONLY as bf16 hi / lo planes [P][C] (the operand format of the tcgen05 convolution
// kernels): the BatchNorm behind a convolution hands the conv's backward passes their dY operand directly.
int tatt_bn_bwd_planes(const float* X, const float* dY, const float* mean, const float* invstd, const float* gamma,
                       const float* beta, int act, int training, long long P, int C, void* dx_hi, void* dx_lo,
                       float* dgamma, float* dbeta, void* ws, void* stream) {
  TATT_REQUIRE(dx_hi && dx_lo && C % 4 == 0 && ((((uintptr_t)dx_hi) | ((uintptr_t)dx_lo)) & 7) == 0,
               "bn_bwd_planes: planes must be given, 8-byte aligned, C %% 4 == 0");
  return bn_bwd_impl(X, dY, mean, invstd, gamma, beta, act, training, P, C, nullptr, dx_hi, dx_lo, dgamma, dbeta, ws, stream);
}
static int bn_bwd_impl(const float* X, const float* dY, const float* mean, const float* invstd, const float* gamma,
                       const float* beta, int act, int training, long long P, int C, float* dX, void* dx_hi, void* dx_lo,
                       float* dgamma, float* dbeta, void* ws, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  TATT_REQUIRE(P >= 1 && C >= 1, "bn_bwd: empty input");
  TATT_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 2 * C, st));
  if (bn64_ok(X, dY, C, P) && bn64_ok(mean, invstd, C, P) && bn64_ok(gamma, beta, C, P)) {
    const int rows = bn64_rows(P);
    bn_bwd_reduce64_kernel<<<ceil_div(P, rows), 256, 0, st>>>(reinterpret_cast<const float4*>(X),
                                                              reinterpret_cast<const float4*>(dY), mean, invstd, gamma,
                                                              beta, act, P, rows, (double*)ws);
    TATT_LAUNCH_CHECK("bn_bwd_reduce64_kernel");
  } else {
    int rows = 128;
    if (P > 128LL * 4096) rows = (int)((P + 4095) / 4096);
    dim3 grid(ceil_div(P, rows), ceil_div(C, 64));
    bn_bwd_reduce_kernel<<<grid, 256, 0, st>>>(X, dY, mean, invstd, gamma, beta, act, P, C, rows, (double*)ws);
    TATT_LAUNCH_CHECK("bn_bwd_reduce_kernel");
  }
  bn_bwd_finalize_kernel<<<ceil_div(C, 128), 128, 0, st>>>((const double*)ws, C, dgamma, dbeta);
  TATT_LAUNCH_CHECK("bn_bwd_finalize_kernel");
  if (dX || dx_hi) {
    TATT_REQUIRE(C % 4 == 0, "bn_bwd: C must be a multiple of 4");
    long long total4 = P * C / 4;
    bn_bwd_apply_kernel<<<ew_blocks(total4), 256, 0, st>>>(X, dY, mean, invstd, gamma, beta, dgamma, dbeta, act,
                                                           training, 1.f / (float)P, total4, C / 4, dX,
                                                           reinterpret_cast<__nv_bfloat16*>(dx_hi),
                                                           reinterpret_cast<__nv_bfloat16*>(dx_lo));
    TATT_LAUNCH_CHECK("bn_bwd_apply_kernel");
  }
  return 0;
}

// Y = LN(X + R) over 64 channels; optionally stores S = X + R, mean, rstd for the backward.
int tatt_layernorm64_fwd(const float* X, const float* R, const float* gamma, const float* beta, float* Y, float* S,
                         float* mean, float* rstd, long long P, float eps, void* stream) {
  if (P <= 0) return 0;
  ln_fwd_kernel<<<ew_blocks(P, 8), 256, 0, (cudaStream_t)stream>>>(X, R, gamma, beta, Y, S, mean, rstd, P, eps);
  TATT_LAUNCH_CHECK("ln_fwd_kernel");
  return 0;
}

// dS (grad wrt the pre-norm sum), dgamma/dbeta (zeroed here, then accumulated).
int tatt_layernorm64_bwd(const float* dY, const float* S, const float* mean, const float* rstd, const float* gamma,
                         float* dS, float* dgamma, float* dbeta, long long P, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  TATT_CUDA(cudaMemsetAsync(dgamma, 0, sizeof(float) * 64, st));
  TATT_CUDA(cudaMemsetAsync(dbeta, 0, sizeof(float) * 64, st));
  if (P <= 0) return 0;
  int blocks = ew_blocks(P, 8 * 16);
  ln_bwd_kernel<<<blocks, 256, 0, st>>>(dY, S, mean, rstd, gamma, dS, dgamma, dbeta, P);
  TATT_LAUNCH_CHECK("ln_bwd_kernel");
  return 0;
}

}  // extern "C"
