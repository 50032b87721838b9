// Attention backward of the TP Interpreter on warp-level tensor cores (mma.sync.m16n8k16, bf16 hi/lo operand split,
// fp32 accumulation: the fp32-parity scheme of the rest of the engine).  Same contract as mha_bwd_kernel (attn.cu):
// Q [N][Lq][64], K / V [N][Lk <= 32][64] projected and unscaled, dO [N][Lq][64]  ->  dQ, dK, dV; dropout masks are the
// bits of attn.cu:drop_scales (16-bit lanes of Philox4x32-10, element (n, h, q), call j8 covers keys 8 j8 .. 8 j8 + 7).
// nn.MultiheadAttention semantics as SURVEY 8a (transformer_v2.py:806-833, 470-484).
//
// The CUDA-core kernel spends ~4800 instructions per (query, head) -- 26 x 16 dot products four times over, every K / V
// element re-read from shared memory per query -- and runs at 36 % issue utilisation (ncu, profiles/r2c_*).  Here a
// warp owns one head and walks over tiles of 16 queries; per tile
//   S  = (Q/4) K^T        [16 q x 32 keys]   4 n-tiles x 3 products
//   P  = softmax(S), Pd = P * drop, dP = (dO V^T) * drop, dS = P * (dP - rowsum(P * dP))     (fragment math, quad shuffles)
//   dQ = dS K / 4          [16 q x 16 d]      accumulator fragments of dS ARE the A fragments (as in gru_mma.cu)
//   dK^T += Q^T dS, dV^T += dO^T Pd   [16 d x 32 keys], operands through a small per-warp transposed staging tile
// i.e. 60 MMAs + ~600 other instructions per 16 (query, head) pairs.  dK^T / dV^T accumulate in registers over the
// warp's tiles and are added to global memory once per CTA.
#include <cuda_bf16.h>
#include "common.cuh"

namespace {

constexpr int AM_LK = 32;
constexpr int AM_KLD = 72;       // bf16 row stride of the K / V tiles [32 keys][64 ch] (conflict-free B-fragment loads)
constexpr int AM_TLD = 40;       // K^T tile [64 ch][32 keys]
constexpr int AM_SLD = 24;       // per-warp staging tiles [rows][16 q]

__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
  const float2 hf = __bfloat1622float2(h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(x0 - hf.x, x1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void split1(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
__device__ __forceinline__ void mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// D += A B with A = (ah + al), B = (bh + bl), dropping al * bl
__device__ __forceinline__ void mma3(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], uint32_t bh0, uint32_t bh1,
                                     uint32_t bl0, uint32_t bl1) {
  mma(c, al, bh0, bh1);
  mma(c, ah, bl0, bl1);
  mma(c, ah, bh0, bh1);
}
__device__ __forceinline__ uint32_t lds32(const __nv_bfloat16* p) { return *reinterpret_cast<const uint32_t*>(p); }

struct AmSmem {
  __nv_bfloat16 Kh[AM_LK][AM_KLD], Kl[AM_LK][AM_KLD];
  __nv_bfloat16 Vh[AM_LK][AM_KLD], Vl[AM_LK][AM_KLD];
  __nv_bfloat16 Kth[64][AM_TLD], Ktl[64][AM_TLD];
  // per warp (= head): Q^T, dO^T [16 d][16 q]; dS^T, Pd^T [32 keys][16 q]; hi and lo
  __nv_bfloat16 Qt[4][2][16][AM_SLD], Ot[4][2][16][AM_SLD], St[4][2][32][AM_SLD], Pt[4][2][32][AM_SLD];
};

__global__ void __launch_bounds__(128)
mha_bwd_mma_kernel(const float* __restrict__ Q, const float* __restrict__ K, const float* __restrict__ V,
                   const float* __restrict__ dO, float* __restrict__ dQ, float* __restrict__ dK, float* __restrict__ dV,
                   int Lq, int Lk, int tiles_per_cta, float pdrop, const unsigned long long* __restrict__ rng,
                   unsigned long long site) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  AmSmem& sm = *reinterpret_cast<AmSmem*>(smem_raw);
  const int n = blockIdx.y;
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;

  // ---- K, V (and K^T) of the sample as bf16 hi / lo tiles; keys >= Lk are zero
  for (int i = threadIdx.x; i < AM_LK * 64; i += 128) {
    const int j = i >> 6, c = i & 63;
    float kv = 0.f, vv = 0.f;
    if (j < Lk) {
      kv = __ldg(K + ((long long)n * Lk + j) * 64 + c);
      vv = __ldg(V + ((long long)n * Lk + j) * 64 + c);
    }
    __nv_bfloat16 a, b;
    split1(kv, a, b);
    sm.Kh[j][c] = a; sm.Kl[j][c] = b;
    sm.Kth[c][j] = a; sm.Ktl[c][j] = b;
    split1(vv, a, b);
    sm.Vh[j][c] = a; sm.Vl[j][c] = b;
  }
  __syncthreads();

  float dKt[4][4], dVt[4][4];                     // [n-tile over keys][d = g, g + 8 ; key = 8 nt + 2t, 2t + 1]
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) dKt[i][e] = dVt[i][e] = 0.f;

  const float drop_sc = pdrop > 0.f ? 1.f / (1.f - pdrop) : 1.f;
  const uint32_t thr = (uint32_t)(pdrop * 65536.f + 0.5f);
  unsigned long long seed = 0, offset = 0;
  if (pdrop > 0.f) {
    seed = rng[0];
    offset = rng[1] * 65536ull + site;
  }
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  const int ntiles = (Lq + 15) / 16;

  for (int tt = 0; tt < tiles_per_cta; ++tt) {
    const int tile = blockIdx.x * tiles_per_cta + tt;
    if (tile >= ntiles) break;
    const int q0 = tile * 16;
    const int qr[2] = {q0 + g, q0 + g + 8};
    const bool val[2] = {qr[0] < Lq, qr[1] < Lq};

    // ---- A fragments of Q / 4 and dO: rows g, g + 8; columns 2t, 2t+1, 2t+8, 2t+9 of the head's 16 channels
    uint32_t qh[4], ql[4], oh[4], ol[4];
    float qv[2][4], ov[2][4];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const long long base = ((long long)n * Lq + (val[r] ? qr[r] : 0)) * 64 + h * 16 + 2 * t;
      float2 a = make_float2(0.f, 0.f), b = a, c = a, d = a;
      if (val[r]) {
        a = __ldg(reinterpret_cast<const float2*>(Q + base));
        b = __ldg(reinterpret_cast<const float2*>(Q + base + 8));
        c = __ldg(reinterpret_cast<const float2*>(dO + base));
        d = __ldg(reinterpret_cast<const float2*>(dO + base + 8));
      }
      qv[r][0] = 0.25f * a.x; qv[r][1] = 0.25f * a.y; qv[r][2] = 0.25f * b.x; qv[r][3] = 0.25f * b.y;
      ov[r][0] = c.x; ov[r][1] = c.y; ov[r][2] = d.x; ov[r][3] = d.y;
    }
    split2(qv[0][0], qv[0][1], qh[0], ql[0]);
    split2(qv[1][0], qv[1][1], qh[1], ql[1]);
    split2(qv[0][2], qv[0][3], qh[2], ql[2]);
    split2(qv[1][2], qv[1][3], qh[3], ql[3]);
    split2(ov[0][0], ov[0][1], oh[0], ol[0]);
    split2(ov[1][0], ov[1][1], oh[1], ol[1]);
    split2(ov[0][2], ov[0][3], oh[2], ol[2]);
    split2(ov[1][2], ov[1][3], oh[3], ol[3]);
    // transposed copies for the dK^T / dV^T products: Qt[d][q], Ot[d][q]
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int d = 2 * t + (e & 1) + 8 * (e >> 1), qq = g + 8 * r;
        __nv_bfloat16 a, b;
        split1(qv[r][e], a, b);
        sm.Qt[h][0][d][qq] = a; sm.Qt[h][1][d][qq] = b;
        split1(ov[r][e], a, b);
        sm.Ot[h][0][d][qq] = a; sm.Ot[h][1][d][qq] = b;
      }

    // ---- S = (Q/4) K^T and dP0 = dO V^T: B[k = d][n = key] = K[key][16 h + d]
    float S[4][4], dP[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) S[nt][e] = dP[nt][e] = 0.f;
      const int kj = 8 * nt + g, c0 = h * 16 + 2 * t;
      mma3(S[nt], qh, ql, lds32(&sm.Kh[kj][c0]), lds32(&sm.Kh[kj][c0 + 8]), lds32(&sm.Kl[kj][c0]), lds32(&sm.Kl[kj][c0 + 8]));
      mma3(dP[nt], oh, ol, lds32(&sm.Vh[kj][c0]), lds32(&sm.Vh[kj][c0 + 8]), lds32(&sm.Vl[kj][c0]), lds32(&sm.Vl[kj][c0 + 8]));
    }

    // ---- softmax over the keys (row g: elements [nt][0..1], row g + 8: [nt][2..3]); quad reductions
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float mx = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = 8 * nt + 2 * t + e;
          if (j >= Lk) S[nt][2 * r + e] = -INFINITY;
          mx = fmaxf(mx, S[nt][2 * r + e]);
        }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      float sum = 0.f;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = 8 * nt + 2 * t + e;
          const float ex = (j < Lk) ? __expf(S[nt][2 * r + e] - mx) : 0.f;
          S[nt][2 * r + e] = ex;
          sum += ex;
        }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      const float inv = 1.f / sum;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        S[nt][2 * r] *= inv;
        S[nt][2 * r + 1] *= inv;
      }
    }

    // ---- dropout scales: thread t of a quad draws Philox call j8 = t of both rows; word t of call nt comes back through
    // four rotating quad shuffles (destination t reads call (t + rot) & 3 from lane (t + rot) & 3, which sends word t)
    float ds[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) ds[nt][e] = 1.f;
    if (pdrop > 0.f) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const long long elem = ((long long)n * 4 + h) * Lq + qr[r];
        const unsigned long long idx = (unsigned long long)(elem * 4 + t);
        const uint4 rr = philox4x32_10(key, make_uint4((uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)offset,
                                                       (uint32_t)(offset >> 32)));
        const uint32_t w[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
        for (int rot = 0; rot < 4; ++rot) {
          const int want = (t - rot) & 3;              // the destination that reads from this lane in this round
          const uint32_t send = want == 0 ? w[0] : (want == 1 ? w[1] : (want == 2 ? w[2] : w[3]));
          const uint32_t got = __shfl_sync(0xffffffffu, send, (lane & ~3) | ((t + rot) & 3));
          const int nt = (t + rot) & 3;                // call index = key block the received word belongs to
          const float s0 = (got & 0xffffu) >= thr ? drop_sc : 0.f, s1 = (got >> 16) >= thr ? drop_sc : 0.f;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k == nt) {
              ds[k][2 * r] = s0;
              ds[k][2 * r + 1] = s1;
            }
        }
      }
    }

    // ---- dS = P (dP ds - sum_j P dP ds), Pd = P ds; invalid rows / keys contribute zero
    float Pd[4][4];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float dot = 0.f;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int i = 2 * r + e;
          dP[nt][i] *= ds[nt][i];
          dot = fmaf(S[nt][i], dP[nt][i], dot);
        }
      dot += __shfl_xor_sync(0xffffffffu, dot, 1);
      dot += __shfl_xor_sync(0xffffffffu, dot, 2);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int i = 2 * r + e;
          const float p = S[nt][i];
          Pd[nt][i] = val[r] ? p * ds[nt][i] : 0.f;
          dP[nt][i] = val[r] ? p * (dP[nt][i] - dot) : 0.f;          // dP now holds dS
        }
    }

    // ---- dQ = dS K / 4: A fragments straight from the dS accumulators; B[k = key][n = d] = K^T[16 h + d][key]
    float dq[2][4];
#pragma unroll
    for (int nd = 0; nd < 2; ++nd)
#pragma unroll
      for (int e = 0; e < 4; ++e) dq[nd][e] = 0.f;
#pragma unroll
    for (int kt = 0; kt < 2; ++kt) {
      uint32_t ah[4], al[4];
      split2(dP[2 * kt][0], dP[2 * kt][1], ah[0], al[0]);
      split2(dP[2 * kt][2], dP[2 * kt][3], ah[1], al[1]);
      split2(dP[2 * kt + 1][0], dP[2 * kt + 1][1], ah[2], al[2]);
      split2(dP[2 * kt + 1][2], dP[2 * kt + 1][3], ah[3], al[3]);
#pragma unroll
      for (int nd = 0; nd < 2; ++nd) {
        const int dr = h * 16 + 8 * nd + g, c0 = 16 * kt + 2 * t;
        mma3(dq[nd], ah, al, lds32(&sm.Kth[dr][c0]), lds32(&sm.Kth[dr][c0 + 8]), lds32(&sm.Ktl[dr][c0]),
             lds32(&sm.Ktl[dr][c0 + 8]));
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r)
      if (val[r]) {
        float* o = dQ + ((long long)n * Lq + qr[r]) * 64 + h * 16 + 2 * t;
        *reinterpret_cast<float2*>(o) = make_float2(0.25f * dq[0][2 * r], 0.25f * dq[0][2 * r + 1]);
        *reinterpret_cast<float2*>(o + 8) = make_float2(0.25f * dq[1][2 * r], 0.25f * dq[1][2 * r + 1]);
      }

    // ---- transposed staging of dS and Pd: St[key][q], Pt[key][q]
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int kj = 8 * nt + 2 * t + (e & 1), qq = g + 8 * (e >> 1);
        __nv_bfloat16 a, b;
        split1(dP[nt][e], a, b);
        sm.St[h][0][kj][qq] = a; sm.St[h][1][kj][qq] = b;
        split1(Pd[nt][e], a, b);
        sm.Pt[h][0][kj][qq] = a; sm.Pt[h][1][kj][qq] = b;
      }
    __syncwarp();

    // ---- dK^T += Q^T dS, dV^T += dO^T Pd: A[row = d][col = q] from Qt / Ot, B[k = q][n = key] from St / Pt
    {
      uint32_t ah[4], al[4], bh[4], bl[4];
      ah[0] = lds32(&sm.Qt[h][0][g][2 * t]);     ah[1] = lds32(&sm.Qt[h][0][g + 8][2 * t]);
      ah[2] = lds32(&sm.Qt[h][0][g][2 * t + 8]); ah[3] = lds32(&sm.Qt[h][0][g + 8][2 * t + 8]);
      al[0] = lds32(&sm.Qt[h][1][g][2 * t]);     al[1] = lds32(&sm.Qt[h][1][g + 8][2 * t]);
      al[2] = lds32(&sm.Qt[h][1][g][2 * t + 8]); al[3] = lds32(&sm.Qt[h][1][g + 8][2 * t + 8]);
      bh[0] = lds32(&sm.Ot[h][0][g][2 * t]);     bh[1] = lds32(&sm.Ot[h][0][g + 8][2 * t]);
      bh[2] = lds32(&sm.Ot[h][0][g][2 * t + 8]); bh[3] = lds32(&sm.Ot[h][0][g + 8][2 * t + 8]);
      bl[0] = lds32(&sm.Ot[h][1][g][2 * t]);     bl[1] = lds32(&sm.Ot[h][1][g + 8][2 * t]);
      bl[2] = lds32(&sm.Ot[h][1][g][2 * t + 8]); bl[3] = lds32(&sm.Ot[h][1][g + 8][2 * t + 8]);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int kj = 8 * nt + g;
        mma3(dKt[nt], ah, al, lds32(&sm.St[h][0][kj][2 * t]), lds32(&sm.St[h][0][kj][2 * t + 8]),
             lds32(&sm.St[h][1][kj][2 * t]), lds32(&sm.St[h][1][kj][2 * t + 8]));
        mma3(dVt[nt], bh, bl, lds32(&sm.Pt[h][0][kj][2 * t]), lds32(&sm.Pt[h][0][kj][2 * t + 8]),
             lds32(&sm.Pt[h][1][kj][2 * t]), lds32(&sm.Pt[h][1][kj][2 * t + 8]));
      }
    }
    __syncwarp();                                  // the staging tiles are rewritten by the next tile
  }

  // ---- dK[key][16 h + d] += dKt[d][key]
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int kj = 8 * nt + 2 * t + (e & 1), d = g + 8 * (e >> 1);
      if (kj < Lk) {
        atomicAdd(dK + ((long long)n * Lk + kj) * 64 + h * 16 + d, dKt[nt][e]);
        atomicAdd(dV + ((long long)n * Lk + kj) * 64 + h * 16 + d, dVt[nt][e]);
      }
    }
}

}  // namespace

// Returns 0 ok, 1 error.  dK / dV must be zeroed by the caller.
int tatt_mha64_bwd_mma_launch(const float* Q, const float* K, const float* V, const float* dO, float* dQ, float* dK,
                              float* dV, int N, int Lq, int Lk, float pdrop, const unsigned long long* rng,
                              unsigned long long site, cudaStream_t st) {
  TATT_CUDA(cudaFuncSetAttribute(mha_bwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AmSmem)));
  const int ntiles = (Lq + 15) / 16;
  int tpc = 16;
  if (ntiles < tpc) tpc = ntiles;
  dim3 grid((ntiles + tpc - 1) / tpc, N);
  mha_bwd_mma_kernel<<<grid, 128, sizeof(AmSmem), st>>>(Q, K, V, dO, dQ, dK, dV, Lq, Lk, tpc, pdrop, rng, site);
  TATT_LAUNCH_CHECK("mha_bwd_mma_kernel");
  return 0;
}
