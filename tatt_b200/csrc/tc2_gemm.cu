// tcgen05 GEMM / implicit-GEMM engine, version 2: operands pre-split into bf16 hi/lo planes.
//
// v1 (tc_gemm.cu) splits fp32 -> bf16 hi/lo inside the main loop; ncu shows it instruction-bound (tensor pipe
// ~10 %, issue slots ~46 %: profiles/r1_ncu_full_tc_conv3x3_v1.csv) because a 3x3 convolution re-splits every
// input element once per tap.  Here each fp32 operand is split ONCE by a streaming kernel into two bf16 planes in
// a caller-provided workspace; the GEMM main loop is then pure 16-byte cp.async copies (zero-filled at image
// borders / tile edges) into the canonical UMMA shared-memory layouts, and tcgen05.mma does the rest:
//   * K-major SWIZZLE_128B tiles for operands whose reduction axis is contiguous (activations x weights,
//     im2col of NHWC activations);
//   * MN-major SWIZZLE_128B tiles for the weight-gradient GEMMs (dW = dY^T X, conv wgrad), whose reduction axis is
//     the pixel/row axis: a shared-memory row is simply one pixel's 64 channels (128 B), no transposition.
// Three MMAs per k-step (lo*hi + hi*lo + hi*hi) accumulate in fp32 in TMEM, as in v1.
#include <cuda_bf16.h>
#include <stdlib.h>
#include "common.cuh"
#include "gemm_params.cuh"

namespace {

constexpr int BM = 128;
constexpr int BN = 64;
constexpr int BK = 64;
constexpr int NT = 256;
constexpr int A_PLANE = BM * 128;   // bytes
constexpr int B_PLANE = BN * 128;
constexpr int STAGE_BYTES = 2 * A_PLANE + 2 * B_PLANE;

enum { K2_DENSE_K = 0, K2_IM2COL_K = 1, K2_DENSE_MN = 2, K2_IM2COL_MN = 3 };

struct Tc2P {
  const __nv_bfloat16 *Ahi, *Alo, *Bhi, *Blo;
  float* C;
  float* partial;                 // split-K: per-split partial tiles [splitk][batch][M][N] (NULL -> atomics on C)
  const float* bias;
  int M, N, K;
  long long lda, ldb, ldc;        // plane leading dimensions (elements) / C leading dimension
  long long sA, sB, sC, sBias;    // batch strides (elements)
  int batch, splitk, kper, flags;
  int cH, cW, cC, KH, KW, padH, padW;
  FastDiv fdHW, fdW, fdC, fdKW, fdCB;   // fdCB: C / 64 (MN kinds)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_c),
      "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// warp-uniform issue: warp 0 runs the (uniform) issue code, one elected lane executes the tcgen05 instruction -- keeps
// descriptors in uniform registers instead of a per-lane waterfall loop around every UTCHMMA
__device__ __forceinline__ void umma_bf16_elect(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_c),
      "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint32_t bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
      "}" ::"r"(bar)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 16-byte async copy global -> shared; src_bytes == 0 zero-fills the destination
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// SWIZZLE_128B descriptors (cute::UMMA::SmemDescriptor): version 1 at [46,48), layout type 2 at [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int KIND, int STAGES>
__global__ void __launch_bounds__(NT, (STAGES == 1) ? 4 : ((STAGES == 2) ? 2 : 1)) tc2_gemm_kernel(const Tc2P p) {
  constexpr bool MN = (KIND == K2_DENSE_MN || KIND == K2_IM2COL_MN);
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) unsigned long long mma_done[STAGES];
  __shared__ uint32_t tmem_base_s;
  __shared__ int s_y[(KIND == K2_IM2COL_K) ? BM : 1];
  __shared__ int s_x[(KIND == K2_IM2COL_K) ? BM : 1];
  __shared__ int s_n[(KIND == K2_IM2COL_K) ? BM : 1];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int zb = blockIdx.z / p.splitk;
  const int zs = blockIdx.z - zb * p.splitk;
  const int kbeg = zs * p.kper;
  const int kend = min(p.K, kbeg + p.kper);
  const __nv_bfloat16* __restrict__ Ahi = p.Ahi + (long long)zb * p.sA;
  const __nv_bfloat16* __restrict__ Alo = p.Alo + (long long)zb * p.sA;
  const __nv_bfloat16* __restrict__ Bhi = p.Bhi + (long long)zb * p.sB;
  const __nv_bfloat16* __restrict__ Blo = p.Blo + (long long)zb * p.sB;
  float* __restrict__ C = p.C + (long long)zb * p.sC;
  const float* __restrict__ bias = p.bias ? p.bias + (long long)zb * p.sBias : nullptr;
  const int HW = p.cH * p.cW;
  int nvalid = p.N - n0;
  if (nvalid > BN) nvalid = BN;
  const int umma_n = (nvalid + 15) & ~15;
  const int nk = (kend > kbeg) ? (kend - kbeg + BK - 1) / BK : 0;
  const uint32_t smem_base = smem_u32(smem);
  const bool single = (p.flags & F_BF16) != 0;   // bf16 mode: hi plane only, one MMA per k-step

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(smem_u32(&mma_done[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"((uint32_t)BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (KIND == K2_IM2COL_K) {
    for (int r = tid; r < BM; r += NT) {
      int gm = m0 + r;
      if (gm < p.M) {
        unsigned n = fd_div((unsigned)gm, p.fdHW);
        unsigned rem = (unsigned)gm - n * (unsigned)HW;
        unsigned y = fd_div(rem, p.fdW);
        s_y[r] = (int)y;
        s_x[r] = (int)(rem - y * (unsigned)p.cW);
        s_n[r] = (int)(n * (unsigned)HW);
      } else {
        s_y[r] = -1000000;
        s_x[r] = 0;
        s_n[r] = 0;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  // ------------------------------------------------------------------ tile loader (cp.async, 16 B chunks)
  auto issue_loads = [&](int kt) {
    const int kb = kbeg + kt * BK;
    const uint32_t st = smem_base + (uint32_t)((kt % STAGES) * STAGE_BYTES);
    const uint32_t a_hi = st, a_lo = st + A_PLANE, b_hi = st + 2 * A_PLANE, b_lo = b_hi + B_PLANE;
    if (!MN) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int id = tid + i * NT;
        const int row = id >> 3, ch = id & 7;
        const int gk = kb + ch * 8;
        long long off = 0;
        bool ok = gk < kend;
        if (KIND == K2_DENSE_K) {
          const int gm = m0 + row;
          ok = ok && gm < p.M;
          off = (long long)gm * p.lda + gk;
        } else {
          unsigned tap = fd_div((unsigned)gk, p.fdC);
          int ci = gk - (int)tap * p.cC;
          unsigned ky = fd_div(tap, p.fdKW);
          int kx = (int)tap - (int)ky * p.KW;
          int iy = s_y[row] + (int)ky - p.padH, ix = s_x[row] + kx - p.padW;
          ok = ok && iy >= 0 && iy < p.cH && ix >= 0 && ix < p.cW;
          off = ((long long)(s_n[row] + iy * p.cW + ix)) * p.cC + ci;
        }
        if (!ok) off = 0;
        const uint32_t d = (uint32_t)(row * 128 + ((ch ^ (row & 7)) << 4));
        cp_async16(a_hi + d, Ahi + off, ok ? 16 : 0);
        if (!single) cp_async16(a_lo + d, Alo + off, ok ? 16 : 0);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int id = tid + i * NT;
        const int row = id >> 3, ch = id & 7;
        const int gk = kb + ch * 8, gn = n0 + row;
        const bool ok = gk < kend && gn < p.N;
        const long long off = ok ? (long long)gn * p.ldb + gk : 0;
        const uint32_t d = (uint32_t)(row * 128 + ((ch ^ (row & 7)) << 4));
        cp_async16(b_hi + d, Bhi + off, ok ? 16 : 0);
        if (!single) cp_async16(b_lo + d, Blo + off, ok ? 16 : 0);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int id = tid + i * NT;
        const int ch = id & 7, k = (id >> 3) & 63, mm = id >> 9;
        const int gk = kb + k;
        bool ok = gk < kend;
        long long off = 0;
        if (KIND == K2_DENSE_MN) {
          const int gm = m0 + mm * 64 + ch * 8;
          ok = ok && gm < p.M;
          off = (long long)gk * p.lda + gm;
        } else {
          const int mb = (m0 >> 6) + mm;          // 64-channel block index over (tap, ci/64)
          ok = ok && mb * 64 < p.M;
          unsigned tap = fd_div((unsigned)mb, p.fdCB);
          int ci0 = (mb - (int)tap * (int)p.fdCB.d) * 64;
          unsigned ky = fd_div(tap, p.fdKW);
          int kx = (int)tap - (int)ky * p.KW;
          unsigned n = fd_div((unsigned)gk, p.fdHW);
          unsigned rem = (unsigned)gk - n * (unsigned)HW;
          unsigned y = fd_div(rem, p.fdW);
          int x = (int)(rem - y * (unsigned)p.cW);
          int iy = (int)y + (int)ky - p.padH, ix = x + kx - p.padW;
          ok = ok && iy >= 0 && iy < p.cH && ix >= 0 && ix < p.cW;
          off = ((long long)((int)(n * (unsigned)HW) + iy * p.cW + ix)) * p.cC + ci0 + ch * 8;
        }
        if (!ok) off = 0;
        const uint32_t d = (uint32_t)(mm * 8192 + k * 128 + ((ch ^ (k & 7)) << 4));
        cp_async16(a_hi + d, Ahi + off, ok ? 16 : 0);
        if (!single) cp_async16(a_lo + d, Alo + off, ok ? 16 : 0);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int id = tid + i * NT;
        const int ch = id & 7, k = id >> 3;
        const int gk = kb + k, gn = n0 + ch * 8;
        const bool ok = gk < kend && gn < p.N;
        const long long off = ok ? (long long)gk * p.ldb + gn : 0;
        const uint32_t d = (uint32_t)(k * 128 + ((ch ^ (k & 7)) << 4));
        cp_async16(b_hi + d, Bhi + off, ok ? 16 : 0);
        if (!single) cp_async16(b_lo + d, Blo + off, ok ? 16 : 0);
      }
    }
  };

  // ------------------------------------------------------------------ main loop
  uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(umma_n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  if (MN) idesc |= (1u << 15) | (1u << 16);
  if (STAGES == 1) {
    // single-stage variant for K <= 64 (one k-tile): 49 KB of shared memory -> 4 CTAs per SM hide the latency
    for (int kt = 0; kt < nk; ++kt) {
      if (kt >= 1) mbar_wait(smem_u32(&mma_done[0]), (uint32_t)((kt - 1) & 1));
      issue_loads(kt);
      cp_async_commit();
      cp_async_wait<0>();
      fence_proxy_async();
      __syncthreads();
      if (warp == 0) {
        tc_fence_after();
        const uint32_t st = smem_base;
        const uint32_t a_hi = st, a_lo = st + A_PLANE, b_hi = st + 2 * A_PLANE, b_lo = b_hi + B_PLANE;
#pragma unroll
        for (int k16 = 0; k16 < BK / 16; ++k16) {
          uint64_t dah, dal, dbh, dbl;
          if (!MN) {
            const uint32_t ko = k16 * 32;
            dah = make_desc(a_hi + ko, 16, 1024);
            dal = make_desc(a_lo + ko, 16, 1024);
            dbh = make_desc(b_hi + ko, 16, 1024);
            dbl = make_desc(b_lo + ko, 16, 1024);
          } else {
            const uint32_t ko = k16 * 2048;
            dah = make_desc(a_hi + ko, 8192, 1024);
            dal = make_desc(a_lo + ko, 8192, 1024);
            dbh = make_desc(b_hi + ko, 8192, 1024);
            dbl = make_desc(b_lo + ko, 8192, 1024);
          }
          if (!single) {
            umma_bf16_elect(tmem_base, dal, dbh, idesc, (kt > 0 || k16 > 0) ? 1u : 0u);
            umma_bf16_elect(tmem_base, dah, dbl, idesc, 1u);
            umma_bf16_elect(tmem_base, dah, dbh, idesc, 1u);
          } else {
            umma_bf16_elect(tmem_base, dah, dbh, idesc, (kt > 0 || k16 > 0) ? 1u : 0u);
          }
        }
        umma_commit_elect(smem_u32(&mma_done[0]));
      }
    }
  } else {
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
      if (s < nk) issue_loads(s);
      cp_async_commit();
    }
    for (int kt = 0; kt < nk; ++kt) {
      const int s = kt % STAGES;
      cp_async_wait<STAGES - 2>();
      fence_proxy_async();
      __syncthreads();
      if (warp == 0) {
        tc_fence_after();
        const uint32_t st = smem_base + (uint32_t)(s * STAGE_BYTES);
        const uint32_t a_hi = st, a_lo = st + A_PLANE, b_hi = st + 2 * A_PLANE, b_lo = b_hi + B_PLANE;
  #pragma unroll
        for (int k16 = 0; k16 < BK / 16; ++k16) {
          uint64_t dah, dal, dbh, dbl;
          if (!MN) {
            const uint32_t ko = k16 * 32;      // 16 bf16 along K inside the 128-byte swizzle atom
            dah = make_desc(a_hi + ko, 16, 1024);
            dal = make_desc(a_lo + ko, 16, 1024);
            dbh = make_desc(b_hi + ko, 16, 1024);
            dbl = make_desc(b_lo + ko, 16, 1024);
          } else {
            const uint32_t ko = k16 * 2048;    // 16 k-rows = two 8-row groups of 1024 B
            dah = make_desc(a_hi + ko, 8192, 1024);
            dal = make_desc(a_lo + ko, 8192, 1024);
            dbh = make_desc(b_hi + ko, 8192, 1024);
            dbl = make_desc(b_lo + ko, 8192, 1024);
          }
          if (!single) {
            umma_bf16_elect(tmem_base, dal, dbh, idesc, (kt > 0 || k16 > 0) ? 1u : 0u);
            umma_bf16_elect(tmem_base, dah, dbl, idesc, 1u);
            umma_bf16_elect(tmem_base, dah, dbh, idesc, 1u);
          } else {
            umma_bf16_elect(tmem_base, dah, dbh, idesc, (kt > 0 || k16 > 0) ? 1u : 0u);
          }
        }
        umma_commit_elect(smem_u32(&mma_done[s]));
      }
      // refill the stage consumed at iteration kt-1 (its MMAs were committed to mma_done[(kt-1)%STAGES])
      const int nxt = kt + STAGES - 1;
      if (nxt < nk) {
        if (kt >= 1) {
          const int ps = (kt - 1) % STAGES;
          mbar_wait(smem_u32(&mma_done[ps]), (uint32_t)(((kt - 1) / STAGES) & 1));
        }
        issue_loads(nxt);
      }
      cp_async_commit();
    }
  }

  // ------------------------------------------------------------------ epilogue
  if (nk > 0) {
    const int sl = (nk - 1) % STAGES;
    mbar_wait(smem_u32(&mma_done[sl]), (uint32_t)(((nk - 1) / STAGES) & 1));
  }
  tc_fence_after();
  {
    const bool atomic = (p.flags & F_ATOMIC) != 0, accum = (p.flags & F_ACCUM) != 0, relu = (p.flags & F_RELU) != 0;
    const bool vecC = (p.flags & F_VECC) != 0;
    const int lane_base = (warp & 3) * 32;
    const int gm = m0 + lane_base + lane;
    const int half = warp >> 2;
    for (int c16 = half; c16 * 16 < umma_n; c16 += 2) {
      uint32_t r[16];
      if (nk > 0) {
        tmem_ld16(tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)(c16 * 16), r);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) r[j] = 0u;
      }
      if (gm < p.M) {
        const int gn0 = n0 + c16 * 16;
        float* dst = C + (long long)gm * p.ldc + gn0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int gn = gn0 + 4 * q + j;
            v[j] = __uint_as_float(r[4 * q + j]);
            if (bias && zs == 0 && gn < p.N) v[j] += __ldg(bias + gn);
          }
          const int gq = gn0 + 4 * q;
          if (atomic && p.partial) {
            float* pd = p.partial + (((long long)zs * p.batch + zb) * p.M + gm) * p.N + gq;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (gq + j < p.N) pd[j] = v[j];
          } else if (atomic) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (gq + j < p.N) atomicAdd(dst + 4 * q + j, v[j]);
          } else if (vecC && gq + 3 < p.N) {
            float4 o = make_float4(v[0], v[1], v[2], v[3]);
            if (accum) {
              float4 old = *reinterpret_cast<const float4*>(dst + 4 * q);
              o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
            }
            if (relu) {
              o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
            }
            *reinterpret_cast<float4*>(dst + 4 * q) = o;
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (gq + j < p.N) {
                float o = v[j];
                if (accum) o += dst[4 * q + j];
                if (relu) o = fmaxf(o, 0.f);
                dst[4 * q + j] = o;
              }
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN) : "memory");
  }
}

// ---------------------------------------------------------------------------------- split kernels
__device__ __forceinline__ void split1(float x, __nv_bfloat16& h, __nv_bfloat16& l) {
  h = __float2bfloat16_rn(x);
  l = __float2bfloat16_rn(x - __bfloat162float(h));
}
// planes[r][c] for c < colsP (zero padded), source row stride ld
__global__ void split_dense_kernel(const float* __restrict__ src, long long ld, long long rows, int cols, int colsP,
                                   __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int vec) {
  const long long n4 = rows * (long long)(colsP / 4);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / (colsP / 4);
    const int c = (int)(i - r * (colsP / 4)) * 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    const float* s = src + r * ld + c;
    if (vec && c + 3 < cols) {
      float4 t = __ldg(reinterpret_cast<const float4*>(s));
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (c + j < cols) v[j] = __ldg(s + j);
    }
    __nv_bfloat16 h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split1(v[j], h[j], l[j]);
    uint2 hv, lv;
    hv.x = (uint32_t)__bfloat16_as_ushort(h[0]) | ((uint32_t)__bfloat16_as_ushort(h[1]) << 16);
    hv.y = (uint32_t)__bfloat16_as_ushort(h[2]) | ((uint32_t)__bfloat16_as_ushort(h[3]) << 16);
    lv.x = (uint32_t)__bfloat16_as_ushort(l[0]) | ((uint32_t)__bfloat16_as_ushort(l[1]) << 16);
    lv.y = (uint32_t)__bfloat16_as_ushort(l[2]) | ((uint32_t)__bfloat16_as_ushort(l[3]) << 16);
    *reinterpret_cast<uint2*>(hi + r * colsP + c) = hv;
    *reinterpret_cast<uint2*>(lo + r * colsP + c) = lv;
  }
}
// same split, plus out[c] += sum_r src[r][c] (the bias gradient).  blockDim.x is a multiple of colsP/4, so every
// thread keeps the same 4 columns while it grid-strides over rows: partial sums stay in registers.
__global__ void split_colsum_kernel(const float* __restrict__ src, long long ld, long long rows, int cols, int colsP,
                                    __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                    float* __restrict__ colsum, int vec) {
  extern __shared__ float red[];          // [colsP]
  const int tx = colsP / 4;
  const int cg = threadIdx.x % tx;
  const int c = cg * 4;
  const long long rstep = ((long long)gridDim.x * blockDim.x) / tx;
  long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / tx;
  for (int i = threadIdx.x; i < colsP; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
  for (; r < rows; r += rstep) {
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    const float* s = src + r * ld + c;
    if (vec && c + 3 < cols) {
      float4 t = __ldg(reinterpret_cast<const float4*>(s));
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (c + j < cols) v[j] = __ldg(s + j);
    }
    __nv_bfloat16 h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      split1(v[j], h[j], l[j]);
      a[j] += v[j];
    }
    uint2 hv, lv;
    hv.x = (uint32_t)__bfloat16_as_ushort(h[0]) | ((uint32_t)__bfloat16_as_ushort(h[1]) << 16);
    hv.y = (uint32_t)__bfloat16_as_ushort(h[2]) | ((uint32_t)__bfloat16_as_ushort(h[3]) << 16);
    lv.x = (uint32_t)__bfloat16_as_ushort(l[0]) | ((uint32_t)__bfloat16_as_ushort(l[1]) << 16);
    lv.y = (uint32_t)__bfloat16_as_ushort(l[2]) | ((uint32_t)__bfloat16_as_ushort(l[3]) << 16);
    *reinterpret_cast<uint2*>(hi + r * colsP + c) = hv;
    *reinterpret_cast<uint2*>(lo + r * colsP + c) = lv;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) atomicAdd(&red[c + j], a[j]);
  __syncthreads();
  for (int i = threadIdx.x; i < cols; i += blockDim.x) atomicAdd(colsum + i, red[i]);
}
// src [K][N] (row stride ld) -> planes [N][Kp]   (small weight matrices)
__global__ void split_transpose_kernel(const float* __restrict__ src, long long ld, int K, int N, int Kp,
                                       __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const long long total = (long long)N * Kp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % Kp);
    const int n = (int)(i / Kp);
    float v = (k < K) ? __ldg(src + (long long)k * ld + n) : 0.f;
    __nv_bfloat16 h, l;
    split1(v, h, l);
    hi[i] = h;
    lo[i] = l;
  }
}

static inline bool aligned16(const void* p) { return (((uintptr_t)p) & 15) == 0; }
static inline long long rup8(long long x) { return (x + 7) & ~7LL; }

static int split_dense(const float* src, long long ld, long long rows, int cols, int colsP, __nv_bfloat16* hi,
                       __nv_bfloat16* lo, cudaStream_t st) {
  if (rows <= 0) return 0;
  long long n4 = rows * (colsP / 4);
  int blocks = (int)((n4 + 255) / 256);
  if (blocks > 148 * 32) blocks = 148 * 32;
  int vec = (ld % 4 == 0 && aligned16(src)) ? 1 : 0;
  split_dense_kernel<<<blocks, 256, 0, st>>>(src, ld, rows, cols, colsP, hi, lo, vec);
  TATT_LAUNCH_CHECK("split_dense_kernel");
  return 0;
}
static int split_transpose(const float* src, long long ld, int K, int N, int Kp, __nv_bfloat16* hi,
                           __nv_bfloat16* lo, cudaStream_t st) {
  long long total = (long long)N * Kp;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 32) blocks = 148 * 32;
  split_transpose_kernel<<<blocks, 256, 0, st>>>(src, ld, K, N, Kp, hi, lo);
  TATT_LAUNCH_CHECK("split_transpose_kernel");
  return 0;
}

template <int KIND, int STAGES>
static int launch_kind_s(const Tc2P& q, cudaStream_t st) {
  const int smem = STAGES * STAGE_BYTES + 1024;
  dim3 grid(ceil_div(q.M, BM), ceil_div(q.N, BN), q.batch * q.splitk);
  TATT_CUDA(cudaFuncSetAttribute(tc2_gemm_kernel<KIND, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  tc2_gemm_kernel<KIND, STAGES><<<grid, NT, smem, st>>>(q);
  TATT_LAUNCH_CHECK("tc2_gemm_kernel");
  return 0;
}
template <int KIND>
static int launch_kind(const Tc2P& q, cudaStream_t st) {
  // pipeline depth: 2 stages x 2 CTAs/SM (default) or 4 stages x 1 CTA/SM (TATT_TC2_STAGES=4)
  static const int stages = []() {
    const char* e = getenv("TATT_TC2_STAGES");
    return (e && e[0] == '4') ? 4 : 2;
  }();
  // weight-gradient kinds stream 27-68 k-tiles per CTA: deeper pipeline (TATT_TC2_STAGES_MN=2 to disable)
  static const int stages_mn = []() {
    const char* e = getenv("TATT_TC2_STAGES_MN");
    return (e && e[0] == '2') ? 2 : 4;
  }();
  constexpr bool kMN = (KIND == K2_DENSE_MN || KIND == K2_IM2COL_MN);
  const bool long_k = (KIND == K2_DENSE_K) && (q.kper / BK >= 12);      // e.g. the RPE recurrent GEMMs
  if (KIND == K2_DENSE_K && q.kper <= BK) return launch_kind_s<KIND, 1>(q, st);   // one k-tile: 4 CTAs/SM
  if (((kMN || long_k) ? stages_mn : stages) == 4) return launch_kind_s<KIND, 4>(q, st);
  return launch_kind_s<KIND, 2>(q, st);
}

}  // namespace

// C[b][m][n] += sum_s partial[s][b][m][n]
__global__ void splitk_reduce_kernel(const float* __restrict__ partial, float* __restrict__ C, int splitk, int batch,
                                     int M, int N, long long ldc, long long sC) {
  const long long per = (long long)batch * M * N;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= per) return;
  float a = 0.f;
  for (int s = 0; s < splitk; ++s) a += partial[(long long)s * per + i];
  const int n = (int)(i % N);
  const long long r = i / N;
  const int m = (int)(r % M);
  const int b = (int)(r / M);
  C[(long long)b * sC + (long long)m * ldc + n] += a;
}

// planes [rows][rup8(cols)] (transpose == 0) or [cols][rup8(rows)] (transpose == 1), zero padded
int tatt_tc2_split(const float* src, long long ld, long long rows, int cols, int transpose, void* hi, void* lo,
                   float* colsum, cudaStream_t st) {
  if (!transpose && colsum) {
    const int colsP = (int)rup8(cols);
    TATT_REQUIRE(colsP <= 1024, "split_bf16: fused column sums need cols <= 1024");
    TATT_CUDA(cudaMemsetAsync(colsum, 0, sizeof(float) * cols, st));
    if (rows <= 0) return 0;
    const int tx = colsP / 4;
    int nthr = (256 / tx) * tx;
    if (nthr < tx) nthr = tx;
    long long want = (rows * tx + nthr - 1) / nthr;
    int blocks = (int)(want < 148LL * 8 ? want : 148LL * 8);
    if (blocks < 1) blocks = 1;
    int vec = (ld % 4 == 0 && aligned16(src)) ? 1 : 0;
    split_colsum_kernel<<<blocks, nthr, sizeof(float) * colsP, st>>>(src, ld, rows, cols, colsP, (__nv_bfloat16*)hi,
                                                                   (__nv_bfloat16*)lo, colsum, vec);
    TATT_LAUNCH_CHECK("split_colsum_kernel");
    return 0;
  }
  TATT_REQUIRE(!colsum, "split_bf16: column sums are not available for the transposed split");
  if (!transpose)
    return split_dense(src, ld, rows, cols, (int)rup8(cols), (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, st);
  return split_transpose(src, ld, (int)rows, cols, (int)rup8(rows), (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, st);
}

// Returns 0 on success, 1 on error, -1 when the shape is not eligible or the workspace is too small
int tatt_tc2_gemm_launch(GemmP p, int amode, int bmode, bool want_split, void* ws, long long ws_bytes,
                         cudaStream_t st) {
  const bool a_pl = (p.flags & F_APLANES) != 0, b_pl = (p.flags & F_BPLANES) != 0;
  if ((ws == nullptr && !(a_pl && b_pl)) || p.N <= 4 || p.K < 32 || p.M < 32) {
    if (a_pl || b_pl) return tatt_set_error("tc2_gemm: pre-split operand planes given for an ineligible shape");
    return -1;
  }
  if (a_pl && ((amode != A_ROW && amode != A_COL) || p.lda % 8))
    return tatt_set_error("tc2_gemm: A planes need amode row/col and lda %% 8 == 0");
  if (b_pl && (!((amode == A_ROW && bmode == B_NK) || (amode == A_COL && bmode == B_KN)) || p.ldb % 8))
    return tatt_set_error("tc2_gemm: B planes need (row,NK) or (col,KN) operands and ldb %% 8 == 0");
  // single-pass small GEMMs (one N tile, one or two k-tiles): the in-loop split of v1 moves fewer bytes than a
  // separate split pass
  if (!a_pl && !b_pl && amode == A_ROW && p.N <= 64 && p.K <= 128 && p.batch == 1) return -1;
  const bool mn = (amode == A_COL || amode == A_IM2COL_T);
  if (mn && bmode != B_KN) return -1;
  Tc2P q = {};
  long long nA, nB;   // plane sizes in elements (per plane, all batches)
  int kind;
  if (amode == A_ROW) {
    if (p.K % 8) return -1;
    kind = K2_DENSE_K;
    q.lda = a_pl ? p.lda : p.K;
    q.sA = a_pl ? p.sA : (long long)p.M * p.K;
    nA = a_pl ? 0 : q.sA * p.batch;
  } else if (amode == A_IM2COL) {
    if (p.cC % 8 || p.batch != 1) return -1;
    kind = K2_IM2COL_K;
    nA = (long long)(p.M) * p.cC;          // M = nimg*H*W pixels
  } else if (amode == A_COL) {
    kind = K2_DENSE_MN;
    q.lda = a_pl ? p.lda : rup8(p.M);
    q.sA = a_pl ? p.sA : (long long)p.K * q.lda;
    nA = a_pl ? 0 : q.sA * p.batch;
  } else {
    if (p.cC % 64 || p.batch != 1) return -1;
    kind = K2_IM2COL_MN;
    nA = (long long)p.K * p.cC;            // K = pixels
  }
  if (b_pl) {
    q.ldb = p.ldb;
    q.sB = p.sB;
  } else if (!mn) {
    q.ldb = rup8(p.K);
    q.sB = (long long)p.N * q.ldb;
  } else {
    q.ldb = rup8(p.N);
    q.sB = (long long)p.K * q.ldb;
  }
  nB = b_pl ? 0 : q.sB * p.batch;
  const long long nA8 = rup8(nA), nB8 = rup8(nB);
  if (nA8 + nB8 > 0 && (ws == nullptr || (long long)sizeof(__nv_bfloat16) * 2 * (nA8 + nB8) > ws_bytes)) {
    if (a_pl || b_pl) return tatt_set_error("tc2_gemm: workspace too small");
    return -1;
  }
  __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(ws);
  if (nA8 + nB8 > 0 && !aligned16(base)) return -1;
  __nv_bfloat16 *Ahi = base, *Alo = base + nA8, *Bhi = base + 2 * nA8, *Blo = base + 2 * nA8 + nB8;
  if (a_pl) {
    Ahi = reinterpret_cast<__nv_bfloat16*>(const_cast<float*>(p.A));
    Alo = Ahi + p.loA;
  }
  if (b_pl) {
    Bhi = reinterpret_cast<__nv_bfloat16*>(const_cast<float*>(p.B));
    Blo = Bhi + p.loB;
  }

  // ---- split passes
  for (int b = 0; b < p.batch; ++b) {
    const float* Ab = p.A + (long long)b * p.sA;
    const float* Bb = p.B + (long long)b * p.sB;
    int rc = 0;
    if (a_pl) {
    } else if (kind == K2_DENSE_K)
      rc = split_dense(Ab, p.lda, p.M, p.K, p.K, Ahi + b * q.sA, Alo + b * q.sA, st);
    else if (kind == K2_DENSE_MN)
      rc = split_dense(Ab, p.lda, p.K, p.M, (int)q.lda, Ahi + b * q.sA, Alo + b * q.sA, st);
    else if (b == 0 && !(p.flags & F_A_VALID))       // F_A_VALID: the workspace still holds these planes
      rc = split_dense(Ab, p.cC, nA / p.cC, p.cC, p.cC, Ahi, Alo, st);
    if (rc) return rc;
    if (b_pl) {
    } else if (!mn) {
      if (bmode == B_NK)
        rc = split_dense(Bb, p.ldb, p.N, p.K, (int)q.ldb, Bhi + b * q.sB, Blo + b * q.sB, st);
      else
        rc = split_transpose(Bb, p.ldb, p.K, p.N, (int)q.ldb, Bhi + b * q.sB, Blo + b * q.sB, st);
    } else {
      rc = split_dense(Bb, p.ldb, p.K, p.N, (int)q.ldb, Bhi + b * q.sB, Blo + b * q.sB, st);
    }
    if (rc) return rc;
  }

  q.Ahi = Ahi; q.Alo = Alo; q.Bhi = Bhi; q.Blo = Blo;
  q.C = p.C; q.bias = p.bias;
  q.M = p.M; q.N = p.N; q.K = p.K;
  q.ldc = p.ldc; q.sC = p.sC; q.sBias = p.sBias;
  q.batch = p.batch;
  q.flags = p.flags & (F_ACCUM | F_RELU | F_VECC | F_BF16);
  q.cH = p.cH; q.cW = p.cW; q.cC = p.cC; q.KH = p.KH; q.KW = p.KW; q.padH = p.padH; q.padW = p.padW;
  q.fdHW = p.fdHW; q.fdW = p.fdW; q.fdC = p.fdC; q.fdKW = p.fdKW;
  q.fdCB = make_fd(p.cC >= 64 ? p.cC / 64 : 1);

  // ---- split-K policy (same as v1)
  q.splitk = 1;
  q.kper = ((p.K + BK - 1) / BK) * BK;
  const long long tiles = (long long)ceil_div(p.M, BM) * ceil_div(p.N, BN) * p.batch;
  bool split = want_split;
  int sk = 1;
  if (want_split) {
    // the MN (weight-gradient) kinds run the 4-stage / 1-CTA-per-SM pipeline: aim at exactly one wave
    long long target = mn ? 148LL : 148LL * 2;
    sk = (int)(target / tiles);
    if (sk < 1) sk = 1;
    int maxsk = ceil_div(p.K, BK * 4);
    if (sk > maxsk) sk = maxsk;
  } else if (tiles < 74 && p.K >= 512 && !(q.flags & F_RELU)) {
    sk = (int)(148 / tiles);            // at most one CTA per SM: a partial second wave would double the tail SMs' time
    if (sk < 1) sk = 1;
    int maxsk = p.K / 256;
    if (sk > maxsk) sk = maxsk;
    if (sk > 1) {
      split = true;
      if (!(q.flags & F_ACCUM)) {
        for (int b = 0; b < p.batch; ++b)
          TATT_CUDA(cudaMemset2DAsync(p.C + (long long)b * p.sC, sizeof(float) * p.ldc, 0, sizeof(float) * p.N,
                                      (size_t)p.M, st));
      }
    }
  }
  if (split) {
    if (sk < 1) sk = 1;
    int kper = ceil_div(p.K, sk);
    kper = ((kper + BK - 1) / BK) * BK;
    q.splitk = ceil_div(p.K, kper);
    q.kper = kper;
    q.flags |= F_ATOMIC;
    q.flags &= ~(F_RELU | F_ACCUM);
  }
  // split-K results: per-split partial tiles in the tail of the workspace + one reduction, instead of splitk-way
  // same-address atomics (the caller has zeroed / pre-filled C either way)
  q.partial = nullptr;
  // measured: not a win for the generic engine at these sizes (36.7 vs 36.0 ms/step) -> opt-in only
  static const bool partial_on = []() {
    const char* e = getenv("TATT_SPLITK_PARTIAL");
    return e && e[0] == '1';
  }();
  if (partial_on && want_split && q.splitk > 1 && ws != nullptr) {
    const long long used = (long long)sizeof(__nv_bfloat16) * 2 * (nA8 + nB8);
    const long long need = (long long)q.splitk * p.batch * p.M * p.N * (long long)sizeof(float);
    const long long off = (used + 255) & ~255LL;
    if (p.M * (long long)p.N <= 1 << 20 && off + need <= ws_bytes)
      q.partial = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(ws) + off);
  }
  int rc;
  switch (kind) {
    case K2_DENSE_K: rc = launch_kind<K2_DENSE_K>(q, st); break;
    case K2_IM2COL_K: rc = launch_kind<K2_IM2COL_K>(q, st); break;
    case K2_DENSE_MN: rc = launch_kind<K2_DENSE_MN>(q, st); break;
    default: rc = launch_kind<K2_IM2COL_MN>(q, st); break;
  }
  if (rc == 0 && q.partial) {
    const long long per = (long long)p.batch * p.M * p.N;
    splitk_reduce_kernel<<<(int)((per + 255) / 256), 256, 0, st>>>(q.partial, p.C, q.splitk, p.batch, p.M, p.N, p.ldc,
                                                                   p.sC);
    TATT_LAUNCH_CHECK("splitk_reduce_kernel");
  }
  return rc;
}
