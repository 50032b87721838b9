// tcgen05 (5th-gen tensor core) GEMM / implicit-GEMM convolution engine for sm_100a.
//
// Same problem family and loader modes as gemm.cu (A_ROW / A_COL / A_IM2COL / A_IM2COL_T, B_KN / B_NK), but
// the math runs on tcgen05.mma with the accumulator in TMEM:
//   * CTA tile 128 (M) x BN (N, 64 or 256; the instruction's N is the runtime valid width rounded to 16),
//     BK = 64, 256 threads, 2-stage shared-memory pipeline tracked by mbarriers (tcgen05.commit).
//   * Operands are fp32 in HBM.  To stay inside the reference's fp32 tolerance (1e-3; single-pass bf16 or tf32
//     is not enough -- SURVEY 8c) every fp32 value x is split in the loader into two bf16 numbers
//     hi = bf16(x), lo = bf16(x - hi) and each k-step issues three MMAs  hi*hi + hi*lo + lo*hi  into the same
//     fp32 TMEM accumulator (relative error ~2^-16 per product).  The split happens in registers on the way from
//     global to shared memory, so no extra HBM traffic is generated.
//   * Shared-memory operand tiles use the canonical K-major SWIZZLE_128B UMMA layout (row = 64 bf16 = 128 B,
//     16-byte chunk c of row r stored at chunk position c ^ (r & 7), 8-row groups 1024 B apart); sources that are
//     contiguous along M/N instead of K (weight-gradient GEMMs, packed conv weights) are transposed in registers
//     by loading 8(k) x 4(m) micro-blocks.
//   * Epilogue: tcgen05.ld (32x32b.x16) TMEM -> registers -> bias / ReLU / accumulate / split-K atomics -> HBM.
#include <cuda_bf16.h>
#include "common.cuh"
#include "gemm_params.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int NT = 256;
constexpr int STAGES = 2;

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
// warp-uniform issue (see tc2_gemm.cu): warp 0 runs the issue code, one elected lane executes the tcgen05 instruction
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
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
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

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout, sm_100):
// [0,14) start>>4 | [16,30) LBO>>4 (=1, unused for swizzled K-major) | [32,46) SBO>>4 (1024 B between 8-row groups)
// | [46,48) version=1 | [61,64) layout type 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 (1<<4), a=b=BF16 (1<<7, 1<<10),
// K-major A and B (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29)
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// split 8 fp32 into 8 bf16 hi + 8 bf16 lo, store both 16-byte chunks at (row, chunk) of a SW128 K-major tile
__device__ __forceinline__ void split_store(unsigned char* hi_tile, unsigned char* lo_tile, int row, int chunk,
                                            const float (&v)[8]) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * i]), h1 = __float2bfloat16_rn(v[2 * i + 1]);
    __nv_bfloat16 l0 = __float2bfloat16_rn(v[2 * i] - __bfloat162float(h0));
    __nv_bfloat16 l1 = __float2bfloat16_rn(v[2 * i + 1] - __bfloat162float(h1));
    h[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    l[i] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
  }
  const int off = row * 128 + ((chunk ^ (row & 7)) << 4);
  *reinterpret_cast<uint4*>(hi_tile + off) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(lo_tile + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }
// 4 consecutive elements along the contiguous axis, limit-guarded
__device__ __forceinline__ float4 ld4_guard(const float* src, int i0, int limit, bool vec) {
  if (i0 >= limit) return zero4();
  if (vec && i0 + 3 < limit) return ldg4(src);
  float4 v = zero4();
  v.x = __ldg(src);
  if (i0 + 1 < limit) v.y = __ldg(src + 1);
  if (i0 + 2 < limit) v.z = __ldg(src + 2);
  if (i0 + 3 < limit) v.w = __ldg(src + 3);
  return v;
}

template <int BN, int AMODE, int BMODE>
__global__ void __launch_bounds__(NT, (BN == 64) ? 2 : 1) tc_gemm_kernel(const GemmP p) {
  constexpr int A_PLANE = BM * 128;           // bytes: 128 rows x 64 bf16
  constexpr int B_PLANE = BN * 128;
  constexpr int STAGE_BYTES = 2 * A_PLANE + 2 * B_PLANE;
  constexpr bool A_KCONT = (AMODE == A_ROW || AMODE == A_IM2COL);
  constexpr bool B_KCONT = (BMODE == B_NK);
  constexpr int A_CH = A_KCONT ? (BM * 8 / NT) : 1;                 // chunks (K-contig) or micro-blocks per thread
  constexpr int B_CH = B_KCONT ? (BN * 8 / NT) : ((BN / 4) * 8 + NT - 1) / NT;

  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) unsigned long long mma_done[STAGES];
  __shared__ uint32_t tmem_base_s;
  __shared__ int s_y[(AMODE == A_IM2COL) ? BM : 1];
  __shared__ int s_x[(AMODE == A_IM2COL) ? BM : 1];
  __shared__ int s_n[(AMODE == A_IM2COL) ? BM : 1];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int zb = blockIdx.z / p.splitk;
  const int zs = blockIdx.z - zb * p.splitk;
  const int kbeg = zs * p.kper;
  const int kend = min(p.K, kbeg + p.kper);
  const float* __restrict__ A = p.A + (long long)zb * p.sA;
  const float* __restrict__ B = p.B + (long long)zb * p.sB;
  float* __restrict__ C = p.C + (long long)zb * p.sC;
  const float* __restrict__ bias = p.bias ? p.bias + (long long)zb * p.sBias : nullptr;
  const bool vecA = (p.flags & F_VECA) != 0, vecB = (p.flags & F_VECB) != 0;
  const int HW = p.cH * p.cW;
  int nvalid = p.N - n0;
  if (nvalid > BN) nvalid = BN;
  const int umma_n = (nvalid + 15) & ~15;
  const int nk = (kend > kbeg) ? (kend - kbeg + BK - 1) / BK : 0;

  // ---------------- one-time setup
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
  if (AMODE == A_IM2COL) {
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

  // A_IM2COL_T: this thread's 4 consecutive (tap, ci) rows are fixed over the K loop
  int t_ci = 0, t_dy = 0, t_dx = 0;
  bool t_ok = false;
  if (AMODE == A_IM2COL_T) {
    int gi = m0 + 4 * (tid >> 3);
    t_ok = gi < p.M;
    if (t_ok) {
      unsigned tap = fd_div((unsigned)gi, p.fdC);
      t_ci = gi - (int)tap * p.cC;
      unsigned ky = fd_div(tap, p.fdKW);
      t_dy = (int)ky - p.padH;
      t_dx = ((int)tap - (int)ky * p.KW) - p.padW;
    }
  }

  float4 ra[A_CH * 2 * (A_KCONT ? 1 : 4)];   // K-contig: 2 float4 per chunk; M-contig: 8 float4 per micro-block
  float4 rb[B_CH * 2 * (B_KCONT ? 1 : 4)];

  auto load_regs = [&](int kb) {
    // ------------------------------------------------ A
    if (A_KCONT) {
#pragma unroll
      for (int i = 0; i < A_CH; ++i) {
        const int id = tid + i * NT;
        const int row = id >> 3, ch = id & 7;
        const int gm = m0 + row, gk = kb + ch * 8;
        float4 v0 = zero4(), v1 = zero4();
        if (AMODE == A_ROW) {
          if (gm < p.M) {
            const float* src = A + (long long)gm * p.lda + gk;
            v0 = ld4_guard(src, gk, kend, vecA);
            v1 = ld4_guard(src + 4, gk + 4, kend, vecA);
          }
        } else {  // A_IM2COL: each float4 (4 channels of one tap) decodes its own tap
          const int y = s_y[row], x = s_x[row], nb = s_n[row];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int k4 = gk + 4 * h;
            float4 v = zero4();
            if (k4 < kend) {
              unsigned tap = fd_div((unsigned)k4, p.fdC);
              int ci = k4 - (int)tap * p.cC;
              unsigned ky = fd_div(tap, p.fdKW);
              int kx = (int)tap - (int)ky * p.KW;
              int iy = y + (int)ky - p.padH, ix = x + kx - p.padW;
              if (iy >= 0 && iy < p.cH && ix >= 0 && ix < p.cW)
                v = ldg4(A + ((long long)(nb + iy * p.cW + ix)) * p.cC + ci);
            }
            if (h == 0) v0 = v; else v1 = v;
          }
        }
        ra[2 * i] = v0;
        ra[2 * i + 1] = v1;
      }
    } else {
      // micro-block: rows 4*rg .. 4*rg+3 (contiguous in memory), k = kb + kc*8 + j
      const int kc = tid & 7, rg = tid >> 3;
      const int gm = m0 + 4 * rg;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int gk = kb + kc * 8 + j;
        float4 v = zero4();
        if (gk < kend) {
          if (AMODE == A_COL) {
            if (gm < p.M) v = ld4_guard(A + (long long)gk * p.lda + gm, gm, p.M, vecA);
          } else {  // A_IM2COL_T: k index = pixel
            if (t_ok) {
              unsigned n = fd_div((unsigned)gk, p.fdHW);
              unsigned rem = (unsigned)gk - n * (unsigned)HW;
              unsigned y = fd_div(rem, p.fdW);
              int x = (int)(rem - y * (unsigned)p.cW);
              int iy = (int)y + t_dy, ix = x + t_dx;
              if (iy >= 0 && iy < p.cH && ix >= 0 && ix < p.cW)
                v = ldg4(A + ((long long)((int)(n * (unsigned)HW) + iy * p.cW + ix)) * p.cC + t_ci);
            }
          }
        }
        ra[j] = v;
      }
    }
    // ------------------------------------------------ B
    if (B_KCONT) {
#pragma unroll
      for (int i = 0; i < B_CH; ++i) {
        const int id = tid + i * NT;
        const int row = id >> 3, ch = id & 7;
        const int gn = n0 + row, gk = kb + ch * 8;
        float4 v0 = zero4(), v1 = zero4();
        if (gn < p.N) {
          const float* src = B + (long long)gn * p.ldb + gk;
          v0 = ld4_guard(src, gk, kend, vecB);
          v1 = ld4_guard(src + 4, gk + 4, kend, vecB);
        }
        rb[2 * i] = v0;
        rb[2 * i + 1] = v1;
      }
    } else {
#pragma unroll
      for (int i = 0; i < B_CH; ++i) {
        const int id = tid + i * NT;
        const int kc = id & 7, rg = id >> 3;
        const int gn = n0 + 4 * rg;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int gk = kb + kc * 8 + j;
          float4 v = zero4();
          if (rg < BN / 4 && gk < kend && gn < p.N) v = ld4_guard(B + (long long)gk * p.ldb + gn, gn, p.N, vecB);
          rb[i * 8 + j] = v;
        }
      }
    }
  };

  auto store_smem = [&](int s) {
    unsigned char* a_hi = smem + s * STAGE_BYTES;
    unsigned char* a_lo = a_hi + A_PLANE;
    unsigned char* b_hi = a_lo + A_PLANE;
    unsigned char* b_lo = b_hi + B_PLANE;
    if (A_KCONT) {
#pragma unroll
      for (int i = 0; i < A_CH; ++i) {
        const int id = tid + i * NT;
        const float v[8] = {ra[2 * i].x, ra[2 * i].y, ra[2 * i].z, ra[2 * i].w,
                            ra[2 * i + 1].x, ra[2 * i + 1].y, ra[2 * i + 1].z, ra[2 * i + 1].w};
        split_store(a_hi, a_lo, id >> 3, id & 7, v);
      }
    } else {
      const int kc = tid & 7, rg = tid >> 3;
      {
        const float v0[8] = {ra[0].x, ra[1].x, ra[2].x, ra[3].x, ra[4].x, ra[5].x, ra[6].x, ra[7].x};
        split_store(a_hi, a_lo, 4 * rg + 0, kc, v0);
        const float v1[8] = {ra[0].y, ra[1].y, ra[2].y, ra[3].y, ra[4].y, ra[5].y, ra[6].y, ra[7].y};
        split_store(a_hi, a_lo, 4 * rg + 1, kc, v1);
        const float v2[8] = {ra[0].z, ra[1].z, ra[2].z, ra[3].z, ra[4].z, ra[5].z, ra[6].z, ra[7].z};
        split_store(a_hi, a_lo, 4 * rg + 2, kc, v2);
        const float v3[8] = {ra[0].w, ra[1].w, ra[2].w, ra[3].w, ra[4].w, ra[5].w, ra[6].w, ra[7].w};
        split_store(a_hi, a_lo, 4 * rg + 3, kc, v3);
      }
    }
    if (B_KCONT) {
#pragma unroll
      for (int i = 0; i < B_CH; ++i) {
        const int id = tid + i * NT;
        const float v[8] = {rb[2 * i].x, rb[2 * i].y, rb[2 * i].z, rb[2 * i].w,
                            rb[2 * i + 1].x, rb[2 * i + 1].y, rb[2 * i + 1].z, rb[2 * i + 1].w};
        split_store(b_hi, b_lo, id >> 3, id & 7, v);
      }
    } else {
#pragma unroll
      for (int i = 0; i < B_CH; ++i) {
        const int id = tid + i * NT;
        const int kc = id & 7, rg = id >> 3;
        if (rg < BN / 4) {
          const float4* r = &rb[i * 8];
          const float v0[8] = {r[0].x, r[1].x, r[2].x, r[3].x, r[4].x, r[5].x, r[6].x, r[7].x};
          split_store(b_hi, b_lo, 4 * rg + 0, kc, v0);
          const float v1[8] = {r[0].y, r[1].y, r[2].y, r[3].y, r[4].y, r[5].y, r[6].y, r[7].y};
          split_store(b_hi, b_lo, 4 * rg + 1, kc, v1);
          const float v2[8] = {r[0].z, r[1].z, r[2].z, r[3].z, r[4].z, r[5].z, r[6].z, r[7].z};
          split_store(b_hi, b_lo, 4 * rg + 2, kc, v2);
          const float v3[8] = {r[0].w, r[1].w, r[2].w, r[3].w, r[4].w, r[5].w, r[6].w, r[7].w};
          split_store(b_hi, b_lo, 4 * rg + 3, kc, v3);
        }
      }
    }
  };

  // ---------------- main loop
  const uint32_t idesc = make_idesc(umma_n);
  if (nk > 0) load_regs(kbeg);
  for (int kt = 0; kt < nk; ++kt) {
    const int s = kt % STAGES;
    if (kt >= STAGES) mbar_wait(smem_u32(&mma_done[s]), (uint32_t)(((kt / STAGES) - 1) & 1));
    store_smem(s);
    if (kt + 1 < nk) load_regs(kbeg + (kt + 1) * BK);
    fence_proxy_async();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      const uint32_t a_hi = smem_u32(smem + s * STAGE_BYTES);
      const uint32_t a_lo = a_hi + A_PLANE;
      const uint32_t b_hi = a_lo + A_PLANE;
      const uint32_t b_lo = b_hi + B_PLANE;
#pragma unroll
      for (int k16 = 0; k16 < BK / 16; ++k16) {
        const uint32_t ko = k16 * 32;   // 16 bf16 = 32 bytes along K inside the 128-byte swizzle atom
        const uint64_t dah = make_desc(a_hi + ko), dal = make_desc(a_lo + ko);
        const uint64_t dbh = make_desc(b_hi + ko), dbl = make_desc(b_lo + ko);
        if (!(p.flags & F_BF16)) {
          umma_bf16_elect(tmem_base, dal, dbh, idesc, (kt > 0 || k16 > 0) ? 1u : 0u);   // small terms first
          umma_bf16_elect(tmem_base, dah, dbl, idesc, 1u);
          umma_bf16_elect(tmem_base, dah, dbh, idesc, 1u);
        } else {
          umma_bf16_elect(tmem_base, dah, dbh, idesc, (kt > 0 || k16 > 0) ? 1u : 0u);   // bf16 mode
        }
      }
      umma_commit_elect(smem_u32(&mma_done[s]));
    }
  }

  // ---------------- epilogue
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
          if (atomic) {
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

template <int BN>
static int launch_bn(const GemmP& p, int amode, int bmode, cudaStream_t st) {
  constexpr int STAGE_BYTES = 2 * BM * 128 + 2 * BN * 128;
  const int smem = STAGES * STAGE_BYTES + 1024;
  dim3 grid(ceil_div(p.M, BM), ceil_div(p.N, BN), p.batch * p.splitk);
#define TATT_TC_CASE(AM, BMo)                                                                                  \
  if (amode == AM && bmode == BMo) {                                                                           \
    TATT_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<BN, AM, BMo>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                   smem));                                                                     \
    tc_gemm_kernel<BN, AM, BMo><<<grid, NT, smem, st>>>(p);                                                    \
    TATT_LAUNCH_CHECK("tc_gemm_kernel");                                                                       \
    return 0;                                                                                                  \
  }
  TATT_TC_CASE(A_ROW, B_NK)
  TATT_TC_CASE(A_ROW, B_KN)
  TATT_TC_CASE(A_COL, B_KN)
  TATT_TC_CASE(A_IM2COL, B_KN)
  TATT_TC_CASE(A_IM2COL_T, B_KN)
#undef TATT_TC_CASE
  return -1;
}

}  // namespace

// p.flags carries F_ACCUM / F_RELU / F_VEC* already resolved by the caller (run_gemm in gemm.cu)
int tatt_tc_gemm_launch(GemmP p, int amode, int bmode, bool want_split, cudaStream_t st) {
  if (p.N <= 4 || p.K < 32 || p.M < 32) return -1;
  int BNc = (p.N <= 64) ? 64 : 256;
  // few-tile problems (the RPE recurrent GEMMs: M = 128): prefer narrow N tiles so more SMs get work
  if (BNc == 256 && (long long)ceil_div(p.M, BM) * ceil_div(p.N, 256) * p.batch < 100) BNc = 64;
  p.splitk = 1;
  p.kper = ((p.K + BK - 1) / BK) * BK;
  const long long tiles = (long long)ceil_div(p.M, BM) * ceil_div(p.N, BNc) * p.batch;
  bool split = want_split;
  int sk = 1;
  if (want_split) {
    long long target = 148LL * 2;
    sk = (int)((target + tiles - 1) / tiles);
    int maxsk = ceil_div(p.K, BK * 4);
    if (sk > maxsk) sk = maxsk;
  } else if (tiles < 74 && p.K >= 512 && !(p.flags & F_RELU)) {
    // automatic split-K: partial sums are atomically added on top of C (zeroed first unless accumulating)
    sk = (int)((148 + tiles - 1) / tiles);
    int maxsk = p.K / 256;
    if (sk > maxsk) sk = maxsk;
    if (sk > 1) {
      split = true;
      if (!(p.flags & F_ACCUM)) {
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
    p.splitk = ceil_div(p.K, kper);
    p.kper = kper;
    p.flags |= F_ATOMIC;
    p.flags &= ~(F_RELU | F_ACCUM);
  }
  if (BNc == 64) return launch_bn<64>(p, amode, bmode, st);
  return launch_bn<256>(p, amode, bmode, st);
}
