// SRB BiGRU recurrences (GruBlock, model/tsrn.py:1067-1084) on the warp-level tensor-core path.
//
// The recurrent product of one time step is tiny (hidden 32 -> 96 gate pre-activations) but it sits on a serial
// chain of T = 32 / 128 steps, so the scan is bound by per-step latency and instruction issue, not by FLOPs.  One
// warp therefore advances SIXTEEN sequences of one direction in lock-step and computes G[16 x 96] = h[16 x 32] W_hh^T
// with mma.sync.m16n8k16 (bf16 hi/lo operand split, 3 MMAs per product, fp32 accumulators: the same fp32-parity
// scheme as the tcgen05 GEMMs).  tcgen05 does not fit here: its minimum tile is M = 64 rows per CTA and every step
// would pay a TMEM round trip, while 16 rows per warp keep 256..1024 independent warps in flight.
//
// Register-resident recurrence: the m16n8 accumulator fragment of h (4 n-tiles of 8 hidden units) IS the A fragment
// of the next step's m16k16 tiles (two adjacent n-tiles = one k-tile), so h never leaves registers.  MMA column c
// of a gate maps to hidden unit perm(c) = 8*((c%8)/2) + 2*(c/8) + c%2, which gives thread (g,t) the 8 CONSECUTIVE
// units 8t..8t+7 of rows g and g+8: every global / shared access of a thread is two 16-byte vectors.
//
// Inputs of the next steps are staged by a per-warp cp.async ring (no registers, no block-level barriers).
// Layouts (unchanged from the scalar kernels in gru.cu):
//   GI [rows][192] = [dir][gate r,z,n][32];  OUT [rows][64] = [dir][32];
//   GATES [rows][320] = [dir][r,z,n,ghn,hprev][32];  dGI / dGH [rows][192].
#include <cuda_bf16.h>
#include "common.cuh"

namespace {

constexpr int GM_ROWS = 16;                      // sequences per warp (MMA M)
constexpr int GM_F_ROWF = 96 + 4;                // staged floats per row, forward (pad 4: conflict-free LDS.128)
constexpr int GM_F_NST = 4;                      // forward ring depth
constexpr int GM_F_STAGE = GM_ROWS * GM_F_ROWF;  // floats
constexpr int GM_B_ROWF = 192 + 4;               // backward: dOUT(32) | r z n ghn hprev (160)
constexpr int GM_B_NST = 3;
constexpr int GM_B_STAGE = GM_ROWS * GM_B_ROWF;

__device__ __forceinline__ int gperm(int c) { return 8 * ((c & 7) >> 1) + 2 * (c >> 3) + (c & 1); }

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// (x0, x1) -> packed bf16x2 hi and lo planes (element 0 in the low half)
__device__ __forceinline__ void split_pair(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
  const float2 hf = __bfloat1622float2(h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(x0 - hf.x, x1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float fsigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float ftanh(float x) { return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x)); }

__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void st8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}

// ------------------------------------------------------------------------------------------------ forward
// HALF: only rows 0..7 of the MMA tile carry sequences (twice the warps for scans with few sequences: the
// elementwise / MUFU work of a step halves, the wasted MMA rows are free on this latency-bound chain)
template <bool HALF>
__global__ void __launch_bounds__(32)
gru32_scan_fwd_mma_kernel(const float* __restrict__ GI, const float* __restrict__ Whh, const float* __restrict__ bhh,
                          float* __restrict__ OUT, float* __restrict__ GATES, int nseq, int T, int s_inner,
                          long long outer_stride, long long inner_stride, long long t_stride) {
  extern __shared__ __align__(16) float smem[];
  float* ring = smem;                                                     // [NST][16][100]
  uint32_t* wlo = reinterpret_cast<uint32_t*>(ring + GM_F_NST * GM_F_STAGE);   // [48][32] lo plane of the W fragments
  float* bsm = reinterpret_cast<float*>(wlo + 48 * 32);                   // [96]
  long long* rowbase = reinterpret_cast<long long*>(bsm + 96);            // [16]
  const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
  const int dir = blockIdx.x & 1;
  constexpr int NROWS = HALF ? 8 : 16;
  const long long seq0 = (long long)(blockIdx.x >> 1) * NROWS;
  if (seq0 >= nseq) return;
  if (lane < NROWS) {
    long long s = seq0 + lane;
    if (s > nseq - 1) s = nseq - 1;                                       // masked rows re-read the last sequence
    rowbase[lane] = (s / s_inner) * outer_stride + (s % s_inner) * inner_stride;
  }
  for (int i = lane; i < 96; i += 32) bsm[i] = bhh[dir * 96 + i];
  // B fragments of W_hh^T: n-tile nt = (gate, j), k-tile kt; B[k][n] = W_hh[gate*32 + perm(8j+g)][perm(k)]
  const float* W = Whh + dir * 96 * 32;
  uint32_t whi[12][2][2];
#pragma unroll
  for (int nt = 0; nt < 12; ++nt)
#pragma unroll
    for (int kt = 0; kt < 2; ++kt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int n = (nt >> 2) * 32 + gperm(8 * (nt & 3) + g);
        const int k0 = gperm(16 * kt + 8 * h + 2 * t);                    // even; perm(c+1) = perm(c)+1
        const float2 w = *reinterpret_cast<const float2*>(W + n * 32 + k0);
        uint32_t lo;
        split_pair(w.x, w.y, whi[nt][kt][h], lo);
        wlo[((nt * 2 + kt) * 2 + h) * 32 + lane] = lo;
      }
  __syncwarp();
  const long long rb[2] = {rowbase[g], HALF ? 0 : rowbase[g + 8]};
  const bool valid[2] = {seq0 + g < nseq, !HALF && seq0 + g + 8 < nseq};

  auto issue = [&](int s) {
    float* dst = ring + (s % GM_F_NST) * GM_F_STAGE;
    const long long toff = (long long)(dir == 0 ? s : T - 1 - s) * t_stride;
#pragma unroll
    for (int c = 0; c < (HALF ? 6 : 12); ++c) {
      const int id = lane + 32 * c;
      const int row = id / 24, ch = id - row * 24;
      cp_async16(smem_addr(dst + row * GM_F_ROWF + ch * 4), GI + (rowbase[row] + toff) * 192 + dir * 96 + ch * 4);
    }
  };
#pragma unroll
  for (int s = 0; s < GM_F_NST - 1; ++s) {
    if (s < T) issue(s);
    cp_async_commit();
  }

  float hq[4][4];                                   // h: [n-tile j][row g: e0,e1 | row g+8: e0,e1]
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) hq[j][i] = 0.f;

  for (int s = 0; s < T; ++s) {
    cp_async_wait<GM_F_NST - 2>();
    __syncwarp();
    if (s + GM_F_NST - 1 < T) issue(s + GM_F_NST - 1);
    cp_async_commit();

    uint32_t ahi[2][4], alo[2][4];
#pragma unroll
    for (int kt = 0; kt < 2; ++kt) {
      split_pair(hq[2 * kt][0], hq[2 * kt][1], ahi[kt][0], alo[kt][0]);
      split_pair(hq[2 * kt][2], hq[2 * kt][3], ahi[kt][1], alo[kt][1]);
      split_pair(hq[2 * kt + 1][0], hq[2 * kt + 1][1], ahi[kt][2], alo[kt][2]);
      split_pair(hq[2 * kt + 1][2], hq[2 * kt + 1][3], ahi[kt][3], alo[kt][3]);
    }
    float acc[12][4];
#pragma unroll
    for (int nt = 0; nt < 12; ++nt) {
      const float2 b = *reinterpret_cast<const float2*>(bsm + (nt >> 2) * 32 + 8 * t + 2 * (nt & 3));
      acc[nt][0] = b.x; acc[nt][1] = b.y; acc[nt][2] = b.x; acc[nt][3] = b.y;
    }
    // small terms first; consecutive MMAs hit different accumulators (12 independent chains of 6)
#pragma unroll
    for (int kt = 0; kt < 2; ++kt)
#pragma unroll
      for (int nt = 0; nt < 12; ++nt) mma16816(acc[nt], alo[kt], whi[nt][kt][0], whi[nt][kt][1]);
#pragma unroll
    for (int kt = 0; kt < 2; ++kt)
#pragma unroll
      for (int nt = 0; nt < 12; ++nt)
        mma16816(acc[nt], ahi[kt], wlo[((nt * 2 + kt) * 2 + 0) * 32 + lane], wlo[((nt * 2 + kt) * 2 + 1) * 32 + lane]);
#pragma unroll
    for (int kt = 0; kt < 2; ++kt)
#pragma unroll
      for (int nt = 0; nt < 12; ++nt) mma16816(acc[nt], ahi[kt], whi[nt][kt][0], whi[nt][kt][1]);
    const float* st = ring + (s % GM_F_NST) * GM_F_STAGE;
    const long long toff = (long long)(dir == 0 ? s : T - 1 - s) * t_stride;
#pragma unroll
    for (int rh = 0; rh < (HALF ? 1 : 2); ++rh) {
      float gr[8], gz[8], gn[8], o_r[8], o_z[8], o_n[8], o_a[8], o_p[8], o_h[8];
      const float* sr = st + (g + 8 * rh) * GM_F_ROWF + 8 * t;
      ld8(sr, gr);
      ld8(sr + 32, gz);
      ld8(sr + 64, gn);
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const int j = m >> 1, i = rh * 2 + (m & 1);
        const float r = fsigmoid(gr[m] + acc[j][i]);
        const float z = fsigmoid(gz[m] + acc[4 + j][i]);
        const float an = acc[8 + j][i];
        const float n = ftanh(gn[m] + r * an);
        const float hp = hq[j][i];
        const float hn = n + z * (hp - n);
        o_r[m] = r; o_z[m] = z; o_n[m] = n; o_a[m] = an; o_p[m] = hp; o_h[m] = hn;
        hq[j][i] = hn;
      }
      if (valid[rh]) {
        const long long row = rb[rh] + toff;
        st8(OUT + row * 64 + dir * 32 + 8 * t, o_h);
        if (GATES) {
          float* gp = GATES + row * 320 + dir * 160 + 8 * t;
          st8(gp, o_r);
          st8(gp + 32, o_z);
          st8(gp + 64, o_n);
          st8(gp + 96, o_a);
          st8(gp + 128, o_p);
        }
      }
    }
  }
  cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------ backward
// dh_prev = dh * z + dGH W_hh  (dGH = [dpr, dpz, dpn * r]): A = dGH [16 x 96], B[n][k] = W_hh[n][k]
template <bool HALF>
__global__ void __launch_bounds__(32)
gru32_scan_bwd_mma_kernel(const float* __restrict__ dOUT, const float* __restrict__ GATES,
                          const float* __restrict__ Whh, float* __restrict__ dGI, float* __restrict__ dGH, int nseq,
                          int T, int s_inner, long long outer_stride, long long inner_stride, long long t_stride) {
  extern __shared__ __align__(16) float smem[];
  float* ring = smem;                                                     // [NST][16][196]
  uint32_t* wlo = reinterpret_cast<uint32_t*>(ring + GM_B_NST * GM_B_STAGE);   // [48][32]
  long long* rowbase = reinterpret_cast<long long*>(wlo + 48 * 32);       // [16]
  const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
  const int dir = blockIdx.x & 1;
  constexpr int NROWS = HALF ? 8 : 16;
  const long long seq0 = (long long)(blockIdx.x >> 1) * NROWS;
  if (seq0 >= nseq) return;
  if (lane < NROWS) {
    long long s = seq0 + lane;
    if (s > nseq - 1) s = nseq - 1;
    rowbase[lane] = (s / s_inner) * outer_stride + (s % s_inner) * inner_stride;
  }
  // B fragments: k-tile kt = (gate, half) over the 96 gate rows, n-tile nt over the 32 hidden units
  const float* W = Whh + dir * 96 * 32;
  uint32_t whi[6][4][2];
#pragma unroll
  for (int kt = 0; kt < 6; ++kt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int n = (kt >> 1) * 32 + gperm(16 * (kt & 1) + 8 * h + 2 * t);   // rows n, n+1
        const int k = gperm(8 * nt + g);
        uint32_t lo;
        split_pair(W[n * 32 + k], W[(n + 1) * 32 + k], whi[kt][nt][h], lo);
        wlo[((kt * 4 + nt) * 2 + h) * 32 + lane] = lo;
      }
  __syncwarp();
  const long long rb[2] = {rowbase[g], HALF ? 0 : rowbase[g + 8]};
  const bool valid[2] = {seq0 + g < nseq, !HALF && seq0 + g + 8 < nseq};

  // the recurrence is walked backwards: step s visits time T-1-s (dir 0) or s (dir 1)
  auto issue = [&](int s) {
    float* dst = ring + (s % GM_B_NST) * GM_B_STAGE;
    const long long toff = (long long)(dir == 0 ? T - 1 - s : s) * t_stride;
#pragma unroll
    for (int c = 0; c < (HALF ? 12 : 24); ++c) {
      const int id = lane + 32 * c;
      const int row = id / 48, ch = id - row * 48;
      const long long r = rowbase[row] + toff;
      const float* src = ch < 8 ? dOUT + r * 64 + dir * 32 + ch * 4 : GATES + r * 320 + dir * 160 + (ch - 8) * 4;
      cp_async16(smem_addr(dst + row * GM_B_ROWF + ch * 4), src);
    }
  };
#pragma unroll
  for (int s = 0; s < GM_B_NST - 1; ++s) {
    if (s < T) issue(s);
    cp_async_commit();
  }

  float dh[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) dh[j][i] = 0.f;

  for (int s = 0; s < T; ++s) {
    cp_async_wait<GM_B_NST - 2>();
    __syncwarp();
    if (s + GM_B_NST - 1 < T) issue(s + GM_B_NST - 1);
    cp_async_commit();

    const float* st = ring + (s % GM_B_NST) * GM_B_STAGE;
    const long long toff = (long long)(dir == 0 ? T - 1 - s : s) * t_stride;
    float dgh[3][4][4];                               // [gate][n-tile j][fragment slot]
    float nxt[3][4][4];                               // per-gate accumulators (12 independent MMA chains)
#pragma unroll
    for (int q = 0; q < 3; ++q)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) nxt[q][j][i] = dgh[q][j][i] = 0.f;
#pragma unroll
    for (int rh = 0; rh < (HALF ? 1 : 2); ++rh) {
      float go[8], vr[8], vz[8], vn[8], va[8], vp[8], o_r[8], o_z[8], o_n[8], o_h[8];
      const float* sr = st + (g + 8 * rh) * GM_B_ROWF + 8 * t;
      ld8(sr, go);
      ld8(sr + 32, vr);
      ld8(sr + 64, vz);
      ld8(sr + 96, vn);
      ld8(sr + 128, va);
      ld8(sr + 160, vp);
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const int j = m >> 1, i = rh * 2 + (m & 1);
        const float r = vr[m], z = vz[m], n = vn[m];
        const float d = go[m] + dh[j][i];
        const float dn = d * (1.f - z);
        const float dz = d * (vp[m] - n);
        const float dpn = dn * (1.f - n * n);
        const float dpr = dpn * va[m] * r * (1.f - r);
        const float dpz = dz * z * (1.f - z);
        const float dhn = dpn * r;
        o_r[m] = dpr; o_z[m] = dpz; o_n[m] = dpn; o_h[m] = dhn;
        dgh[0][j][i] = dpr; dgh[1][j][i] = dpz; dgh[2][j][i] = dhn;
        nxt[0][j][i] = d * z;
      }
      if (valid[rh]) {
        const long long row = rb[rh] + toff;
        float* o = dGI + row * 192 + dir * 96 + 8 * t;
        st8(o, o_r);
        st8(o + 32, o_z);
        st8(o + 64, o_n);
        float* o2 = dGH + row * 192 + dir * 96 + 8 * t;
        st8(o2, o_r);
        st8(o2 + 32, o_z);
        st8(o2 + 64, o_h);
      }
    }
    uint32_t ahi[6][4], alo[6][4];
#pragma unroll
    for (int kt = 0; kt < 6; ++kt) {
      const int q = kt >> 1, j0 = 2 * (kt & 1);
      split_pair(dgh[q][j0][0], dgh[q][j0][1], ahi[kt][0], alo[kt][0]);
      split_pair(dgh[q][j0][2], dgh[q][j0][3], ahi[kt][1], alo[kt][1]);
      split_pair(dgh[q][j0 + 1][0], dgh[q][j0 + 1][1], ahi[kt][2], alo[kt][2]);
      split_pair(dgh[q][j0 + 1][2], dgh[q][j0 + 1][3], ahi[kt][3], alo[kt][3]);
    }
    // k-tile order 0,2,4,1,3,5: the same (gate, n-tile) accumulator recurs every 12 MMAs
#pragma unroll
    for (int kk = 0; kk < 6; ++kk) {
      const int kt = (kk % 3) * 2 + kk / 3;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) mma16816(nxt[kt >> 1][nt], alo[kt], whi[kt][nt][0], whi[kt][nt][1]);
    }
#pragma unroll
    for (int kk = 0; kk < 6; ++kk) {
      const int kt = (kk % 3) * 2 + kk / 3;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
        mma16816(nxt[kt >> 1][nt], ahi[kt], wlo[((kt * 4 + nt) * 2 + 0) * 32 + lane],
                 wlo[((kt * 4 + nt) * 2 + 1) * 32 + lane]);
    }
#pragma unroll
    for (int kk = 0; kk < 6; ++kk) {
      const int kt = (kk % 3) * 2 + kk / 3;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) mma16816(nxt[kt >> 1][nt], ahi[kt], whi[kt][nt][0], whi[kt][nt][1]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) dh[j][i] = nxt[0][j][i] + nxt[1][j][i] + nxt[2][j][i];
  }
  cp_async_wait<0>();
}

constexpr size_t GM_F_SMEM = sizeof(float) * (GM_F_NST * GM_F_STAGE + 96) + 4 * 48 * 32 + 8 * GM_ROWS;
constexpr size_t GM_B_SMEM = sizeof(float) * (GM_B_NST * GM_B_STAGE) + 4 * 48 * 32 + 8 * GM_ROWS;

}  // namespace

int tatt_gru32_scan_fwd_mma_launch(const float* GI, const float* Whh, const float* bhh, float* OUT, float* GATES,
                                   int nseq, int T, int s_inner, long long outer_stride, long long inner_stride,
                                   long long t_stride, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    TATT_CUDA(cudaFuncSetAttribute(gru32_scan_fwd_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)GM_F_SMEM));
    TATT_CUDA(cudaFuncSetAttribute(gru32_scan_fwd_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)GM_F_SMEM));
    attr = true;
  }
  // fewer than one 16-row warp per SM sub-partition (4 x 148): put 8 sequences in each warp instead
  const int blocks16 = 2 * ((nseq + 15) / 16);
  if (blocks16 < 4 * 148) {
    gru32_scan_fwd_mma_kernel<true><<<2 * ((nseq + 7) / 8), 32, GM_F_SMEM, st>>>(
        GI, Whh, bhh, OUT, GATES, nseq, T, s_inner, outer_stride, inner_stride, t_stride);
  } else {
    gru32_scan_fwd_mma_kernel<false><<<blocks16, 32, GM_F_SMEM, st>>>(GI, Whh, bhh, OUT, GATES, nseq, T, s_inner,
                                                                      outer_stride, inner_stride, t_stride);
  }
  TATT_LAUNCH_CHECK("gru32_scan_fwd_mma_kernel");
  return 0;
}

int tatt_gru32_scan_bwd_mma_launch(const float* dOUT, const float* GATES, const float* Whh, float* dGI, float* dGH,
                                   int nseq, int T, int s_inner, long long outer_stride, long long inner_stride,
                                   long long t_stride, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    TATT_CUDA(cudaFuncSetAttribute(gru32_scan_bwd_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)GM_B_SMEM));
    TATT_CUDA(cudaFuncSetAttribute(gru32_scan_bwd_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)GM_B_SMEM));
    attr = true;
  }
  const int blocks16 = 2 * ((nseq + 15) / 16);
  if (blocks16 < 4 * 148) {
    gru32_scan_bwd_mma_kernel<true><<<2 * ((nseq + 7) / 8), 32, GM_B_SMEM, st>>>(
        dOUT, GATES, Whh, dGI, dGH, nseq, T, s_inner, outer_stride, inner_stride, t_stride);
  } else {
    gru32_scan_bwd_mma_kernel<false><<<blocks16, 32, GM_B_SMEM, st>>>(dOUT, GATES, Whh, dGI, dGH, nseq, T, s_inner,
                                                                      outer_stride, inner_stride, t_stride);
  }
  TATT_LAUNCH_CHECK("gru32_scan_bwd_mma_kernel");
  return 0;
}
