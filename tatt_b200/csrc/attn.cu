// Multi-head attention core for the TP Interpreter: d_model 64, 4 heads x head_dim 16, <= 32 keys.
// Used for the decoder cross-attention (L_q = H*W image tokens, L_k = 26 text-prior tokens;
// TransformerDecoderLayer_TP.forward_post, model/transformer_v2.py:806-833) and for the encoder
// self-attention (L_q = L_k = 26; transformer_v2.py:470-484).  Inputs are the already projected
// Q/K/V ([N][L][64], token-major); the kernel does scale, QK^T, softmax, (dropout), PV and the
// head-averaged weights nn.MultiheadAttention returns (need_weights=True).
// One CTA = 32 queries x 4 heads (warp == head, lane == query); K/V of the sample live in SMEM.
#include <stdlib.h>
#include "common.cuh"

namespace {

constexpr int HD = 16;    // head dim
constexpr int NH = 4;     // heads
constexpr int DM = 64;    // model dim
constexpr int LKMAX = 32;

__device__ __forceinline__ void load_kv(const float* __restrict__ K, const float* __restrict__ V, int Lk,
                                        float (*Ks)[DM], float (*Vs)[DM], int tid, int nthreads) {
  for (int i = tid; i < Lk * (DM / 4); i += nthreads) {
    int r = i / (DM / 4), c4 = i % (DM / 4);
    *reinterpret_cast<float4*>(&Ks[r][c4 * 4]) = __ldg(reinterpret_cast<const float4*>(K + r * DM) + c4);
    *reinterpret_cast<float4*>(&Vs[r][c4 * 4]) = __ldg(reinterpret_cast<const float4*>(V + r * DM) + c4);
  }
}

// softmax probabilities p[j] (pre-dropout) for one (query, head); qv is the scaled query
__device__ __forceinline__ void probs(const float qv[HD], const float (*Ks)[DM], int h, int Lk, float p[LKMAX]) {
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < LKMAX; ++j) {
    float s = -INFINITY;
    if (j < Lk) {
      s = 0.f;
#pragma unroll
      for (int d4 = 0; d4 < HD / 4; ++d4) {
        float4 k = *reinterpret_cast<const float4*>(&Ks[j][h * HD + d4 * 4]);
        s = fmaf(qv[d4 * 4 + 0], k.x, s);
        s = fmaf(qv[d4 * 4 + 1], k.y, s);
        s = fmaf(qv[d4 * 4 + 2], k.z, s);
        s = fmaf(qv[d4 * 4 + 3], k.w, s);
      }
    }
    p[j] = s;
    mx = fmaxf(mx, s);
  }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < LKMAX; ++j) {
    float e = (j < Lk) ? __expf(p[j] - mx) : 0.f;   // ex2.approx: ~2 ulp, far inside the 1e-3 tolerance
    p[j] = e;
    sum += e;
  }
  float inv = 1.f / sum;
#pragma unroll
  for (int j = 0; j < LKMAX; ++j) p[j] *= inv;
}

// keep-mask scale factors (0 or 1/(1-p)) for the 32 key slots of element (n,h,q).  16 random bits per slot (keep iff
// u16 >= round(p * 65536): the keep probability is exact to 2^-16), i.e. four Philox4x32 calls per (query, head)
// instead of eight -- the kernels are bound by this integer work, not by the 26 x 16 dot products.
__device__ __forceinline__ void drop_scales(float ds[LKMAX], float pdrop, unsigned long long seed,
                                            unsigned long long offset, long long elem) {
  const float sc = 1.f / (1.f - pdrop);
  const uint32_t thr = (uint32_t)(pdrop * 65536.f + 0.5f);
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
#pragma unroll
  for (int j8 = 0; j8 < LKMAX / 8; ++j8) {
    const unsigned long long idx = (unsigned long long)(elem * (LKMAX / 8) + j8);
    const uint4 r = philox4x32_10(key, make_uint4((uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)offset,
                                                  (uint32_t)(offset >> 32)));
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      ds[j8 * 8 + 2 * k + 0] = (w[k] & 0xffffu) >= thr ? sc : 0.f;
      ds[j8 * 8 + 2 * k + 1] = (w[k] >> 16) >= thr ? sc : 0.f;
    }
  }
}

__global__ void __launch_bounds__(128)
mha_fwd_kernel(const float* __restrict__ Q, const float* __restrict__ K, const float* __restrict__ V,
               float* __restrict__ O, float* __restrict__ AW, int Lq, int Lk, float pdrop,
               const unsigned long long* __restrict__ rng, unsigned long long site) {
  __shared__ __align__(16) float Ks[LKMAX][DM];
  __shared__ __align__(16) float Vs[LKMAX][DM];
  __shared__ float Ps[NH][32][LKMAX + 1];
  const int n = blockIdx.y;
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * 32 + lane;
  const bool valid = q < Lq;
  load_kv(K + (long long)n * Lk * DM, V + (long long)n * Lk * DM, Lk, Ks, Vs, threadIdx.x, 128);
  __syncthreads();

  float qv[HD];
  if (valid) {
    const float4* qp = reinterpret_cast<const float4*>(Q + ((long long)n * Lq + q) * DM + h * HD);
#pragma unroll
    for (int d4 = 0; d4 < HD / 4; ++d4) {
      float4 t = __ldg(qp + d4);
      qv[d4 * 4 + 0] = t.x * 0.25f;
      qv[d4 * 4 + 1] = t.y * 0.25f;
      qv[d4 * 4 + 2] = t.z * 0.25f;
      qv[d4 * 4 + 3] = t.w * 0.25f;
    }
  } else {
#pragma unroll
    for (int d = 0; d < HD; ++d) qv[d] = 0.f;
  }
  float p[LKMAX];
  probs(qv, Ks, h, Lk, p);
  if (pdrop > 0.f) {
    float ds[LKMAX];
    drop_scales(ds, pdrop, rng[0], rng[1] * 65536ull + site, ((long long)n * NH + h) * Lq + q);
#pragma unroll
    for (int j = 0; j < LKMAX; ++j) p[j] *= ds[j];
  }
  float o[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) o[d] = 0.f;
#pragma unroll
  for (int j = 0; j < LKMAX; ++j) {
    if (j < Lk) {
#pragma unroll
      for (int d4 = 0; d4 < HD / 4; ++d4) {
        float4 v = *reinterpret_cast<const float4*>(&Vs[j][h * HD + d4 * 4]);
        o[d4 * 4 + 0] = fmaf(p[j], v.x, o[d4 * 4 + 0]);
        o[d4 * 4 + 1] = fmaf(p[j], v.y, o[d4 * 4 + 1]);
        o[d4 * 4 + 2] = fmaf(p[j], v.z, o[d4 * 4 + 2]);
        o[d4 * 4 + 3] = fmaf(p[j], v.w, o[d4 * 4 + 3]);
      }
    }
  }
  if (valid) {
    float4* op = reinterpret_cast<float4*>(O + ((long long)n * Lq + q) * DM + h * HD);
#pragma unroll
    for (int d4 = 0; d4 < HD / 4; ++d4)
      op[d4] = make_float4(o[d4 * 4 + 0], o[d4 * 4 + 1], o[d4 * 4 + 2], o[d4 * 4 + 3]);
  }
  if (AW) {
#pragma unroll
    for (int j = 0; j < LKMAX; ++j) Ps[h][lane][j] = p[j];
    __syncthreads();
    const int q0 = blockIdx.x * 32;
    int nq = Lq - q0;
    if (nq > 32) nq = 32;
    for (int i = threadIdx.x; i < nq * Lk; i += 128) {
      int qq = i / Lk, j = i - qq * Lk;
      float a = 0.25f * (Ps[0][qq][j] + Ps[1][qq][j] + Ps[2][qq][j] + Ps[3][qq][j]);
      AW[((long long)n * Lq + q0) * Lk + i] = a;
    }
  }
}

// (3 CTAs per SM through __launch_bounds__(128, 3) -- 168 registers, small spills -- was measured: 413 us vs 410 us,
// the kernel is not occupancy-bound)
struct BwdSmem {
  float Ks[LKMAX][DM];
  float Vs[LKMAX][DM];
  float dS[NH][32][LKMAX + 1];
  float Pd[NH][32][LKMAX + 1];
  float Qs[NH][32][HD + 4];
  float dOs[NH][32][HD + 4];
};

__global__ void __launch_bounds__(128)
mha_bwd_kernel(const float* __restrict__ Q, const float* __restrict__ K, const float* __restrict__ V,
               const float* __restrict__ dO, float* __restrict__ dQ, float* __restrict__ dK,
               float* __restrict__ dV, int Lq, int Lk, int tiles_per_cta, float pdrop,
               const unsigned long long* __restrict__ rng, unsigned long long site) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BwdSmem& sm = *reinterpret_cast<BwdSmem*>(smem_raw);
  const int n = blockIdx.y;
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  load_kv(K + (long long)n * Lk * DM, V + (long long)n * Lk * DM, Lk, sm.Ks, sm.Vs, threadIdx.x, 128);
  __syncthreads();

  float accK[4][4], accV[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) accK[i][e] = accV[i][e] = 0.f;

  const int ntiles = (Lq + 31) / 32;
  for (int tt = 0; tt < tiles_per_cta; ++tt) {
    const int tile = blockIdx.x * tiles_per_cta + tt;
    if (tile >= ntiles) break;
    const int q = tile * 32 + lane;
    const bool valid = q < Lq;
    float qv[HD], g[HD];
    if (valid) {
      const float4* qp = reinterpret_cast<const float4*>(Q + ((long long)n * Lq + q) * DM + h * HD);
      const float4* gp = reinterpret_cast<const float4*>(dO + ((long long)n * Lq + q) * DM + h * HD);
#pragma unroll
      for (int d4 = 0; d4 < HD / 4; ++d4) {
        float4 t = __ldg(qp + d4);
        qv[d4 * 4 + 0] = t.x * 0.25f;
        qv[d4 * 4 + 1] = t.y * 0.25f;
        qv[d4 * 4 + 2] = t.z * 0.25f;
        qv[d4 * 4 + 3] = t.w * 0.25f;
        float4 u = __ldg(gp + d4);
        g[d4 * 4 + 0] = u.x;
        g[d4 * 4 + 1] = u.y;
        g[d4 * 4 + 2] = u.z;
        g[d4 * 4 + 3] = u.w;
      }
    } else {
#pragma unroll
      for (int d = 0; d < HD; ++d) qv[d] = g[d] = 0.f;
    }
    float p[LKMAX];
    probs(qv, sm.Ks, h, Lk, p);
    float ds[LKMAX];
    if (pdrop > 0.f) {
      drop_scales(ds, pdrop, rng[0], rng[1] * 65536ull + site, ((long long)n * NH + h) * Lq + q);
    } else {
#pragma unroll
      for (int j = 0; j < LKMAX; ++j) ds[j] = 1.f;
    }
    // dp_j = (dO . V_j) * ds_j ; dS_j = p_j (dp_j - sum_k p_k dp_k)
    float dp[LKMAX];
    float dot = 0.f;
#pragma unroll
    for (int j = 0; j < LKMAX; ++j) {
      float a = 0.f;
      if (j < Lk) {
#pragma unroll
        for (int d4 = 0; d4 < HD / 4; ++d4) {
          float4 v = *reinterpret_cast<const float4*>(&sm.Vs[j][h * HD + d4 * 4]);
          a = fmaf(g[d4 * 4 + 0], v.x, a);
          a = fmaf(g[d4 * 4 + 1], v.y, a);
          a = fmaf(g[d4 * 4 + 2], v.z, a);
          a = fmaf(g[d4 * 4 + 3], v.w, a);
        }
        a *= ds[j];
      }
      dp[j] = a;
      dot = fmaf(p[j], a, dot);
    }
    float dq[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) dq[d] = 0.f;
#pragma unroll
    for (int j = 0; j < LKMAX; ++j) {
      float dsj = (valid && j < Lk) ? p[j] * (dp[j] - dot) : 0.f;
      sm.dS[h][lane][j] = dsj;
      sm.Pd[h][lane][j] = valid ? p[j] * ds[j] : 0.f;
      if (j < Lk) {
#pragma unroll
        for (int d4 = 0; d4 < HD / 4; ++d4) {
          float4 k = *reinterpret_cast<const float4*>(&sm.Ks[j][h * HD + d4 * 4]);
          dq[d4 * 4 + 0] = fmaf(dsj, k.x, dq[d4 * 4 + 0]);
          dq[d4 * 4 + 1] = fmaf(dsj, k.y, dq[d4 * 4 + 1]);
          dq[d4 * 4 + 2] = fmaf(dsj, k.z, dq[d4 * 4 + 2]);
          dq[d4 * 4 + 3] = fmaf(dsj, k.w, dq[d4 * 4 + 3]);
        }
      }
    }
#pragma unroll
    for (int d = 0; d < HD; ++d) {
      sm.Qs[h][lane][d] = qv[d];
      sm.dOs[h][lane][d] = g[d];
    }
    if (valid) {
      float4* qo = reinterpret_cast<float4*>(dQ + ((long long)n * Lq + q) * DM + h * HD);
#pragma unroll
      for (int d4 = 0; d4 < HD / 4; ++d4)
        qo[d4] = make_float4(0.25f * dq[d4 * 4 + 0], 0.25f * dq[d4 * 4 + 1], 0.25f * dq[d4 * 4 + 2],
                             0.25f * dq[d4 * 4 + 3]);
    }
    __syncwarp();
    // dK[j][d] += sum_q dS[q][j] * qv[q][d] ; dV[j][d] += sum_q Pd[q][j] * dO[q][d]
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = lane + 32 * i;
      const int j = idx >> 2, dg = idx & 3;
      if (j < Lk) {
#pragma unroll 8
        for (int qq = 0; qq < 32; ++qq) {
          float a = sm.dS[h][qq][j], b = sm.Pd[h][qq][j];
          float4 x = *reinterpret_cast<const float4*>(&sm.Qs[h][qq][dg * 4]);
          float4 y = *reinterpret_cast<const float4*>(&sm.dOs[h][qq][dg * 4]);
          accK[i][0] = fmaf(a, x.x, accK[i][0]);
          accK[i][1] = fmaf(a, x.y, accK[i][1]);
          accK[i][2] = fmaf(a, x.z, accK[i][2]);
          accK[i][3] = fmaf(a, x.w, accK[i][3]);
          accV[i][0] = fmaf(b, y.x, accV[i][0]);
          accV[i][1] = fmaf(b, y.y, accV[i][1]);
          accV[i][2] = fmaf(b, y.z, accV[i][2]);
          accV[i][3] = fmaf(b, y.w, accV[i][3]);
        }
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = lane + 32 * i;
    const int j = idx >> 2, dg = idx & 3;
    if (j < Lk) {
      float* kd = dK + ((long long)n * Lk + j) * DM + h * HD + dg * 4;
      float* vd = dV + ((long long)n * Lk + j) * DM + h * HD + dg * 4;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        atomicAdd(kd + e, accK[i][e]);
        atomicAdd(vd + e, accV[i][e]);
      }
    }
  }
}

}  // namespace

// warp-level tensor-core version of the backward (attn_mma.cu)
int tatt_mha64_bwd_mma_launch(const float* Q, const float* K, const float* V, const float* dO, float* dQ, float* dK,
                              float* dV, int N, int Lq, int Lk, float pdrop, const unsigned long long* rng,
                              unsigned long long site, cudaStream_t st);

extern "C" {

// Q [N][Lq][64], K/V [N][Lk][64] (projected, unscaled) -> O [N][Lq][64]; AW [N][Lq][Lk] head-averaged
// (post-dropout, like torch) or NULL.
int tatt_mha64_fwd(const float* Q, const float* K, const float* V, float* O, float* AW, int N, int Lq, int Lk,
                   float pdrop, const unsigned long long* rng, unsigned long long site, void* stream) {
  TATT_REQUIRE(pdrop == 0.f || rng != nullptr, "mha64_fwd: dropout needs an rng state pointer");
  TATT_REQUIRE(Lk >= 1 && Lk <= LKMAX, "mha64_fwd: Lk=%d out of range [1,%d]", Lk, LKMAX);
  TATT_REQUIRE(pdrop >= 0.f && pdrop < 1.f, "mha64_fwd: bad dropout p");
  if (N <= 0 || Lq <= 0) return 0;
  dim3 grid((Lq + 31) / 32, N);
  mha_fwd_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(Q, K, V, O, AW, Lq, Lk, pdrop, rng, site);
  TATT_LAUNCH_CHECK("mha_fwd_kernel");
  return 0;
}

// dQ [N][Lq][64] written; dK, dV [N][Lk][64] zeroed here then accumulated with atomics.
int tatt_mha64_bwd(const float* Q, const float* K, const float* V, const float* dO, float* dQ, float* dK, float* dV,
                   int N, int Lq, int Lk, float pdrop, const unsigned long long* rng, unsigned long long site,
                   void* stream) {
  TATT_REQUIRE(pdrop == 0.f || rng != nullptr, "mha64_bwd: dropout needs an rng state pointer");
  TATT_REQUIRE(Lk >= 1 && Lk <= LKMAX, "mha64_bwd: Lk=%d out of range [1,%d]", Lk, LKMAX);
  cudaStream_t st = (cudaStream_t)stream;
  if (N <= 0) return 0;
  TATT_CUDA(cudaMemsetAsync(dK, 0, sizeof(float) * (size_t)N * Lk * DM, st));
  TATT_CUDA(cudaMemsetAsync(dV, 0, sizeof(float) * (size_t)N * Lk * DM, st));
  if (Lq <= 0) return 0;
  static const bool use_mma = []() {               // TATT_MHA_MMA=0: the CUDA-core kernel below
    const char* e = getenv("TATT_MHA_MMA");
    return !(e && e[0] == '0');
  }();
  if (use_mma) return tatt_mha64_bwd_mma_launch(Q, K, V, dO, dQ, dK, dV, N, Lq, Lk, pdrop, rng, site, st);
  TATT_CUDA(cudaFuncSetAttribute(mha_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)sizeof(BwdSmem)));
  int ntiles = (Lq + 31) / 32;
  int tpc = 8;
  if (ntiles < 8) tpc = ntiles;
  dim3 grid((ntiles + tpc - 1) / tpc, N);
  mha_bwd_kernel<<<grid, 128, sizeof(BwdSmem), st>>>(Q, K, V, dO, dQ, dK, dV, Lq, Lk, tpc, pdrop, rng, site);
  TATT_LAUNCH_CHECK("mha_bwd_kernel");
  return 0;
}

}  // extern "C"
