// Row-panel GEMMs of the per-pixel linear layers with the fp32 -> bf16 hi/lo operand split FUSED into the loader.
//
// Every 1x1 convolution / nn.Linear / GRU input projection of the path (model/tsrn.py:1070,1075 GruBlock.conv1,
// nn.GRU weight_ih; model/transformer_v2.py:453-458,785-790 attention in/out projections and FFN) is a GEMM
// Y[P][N] = X[P][K] W[N][K]^T over P = N*H*W pixel rows with K, N in {64, 128, 192}: ~1 FLOP per byte, i.e. bound
// by HBM, not by the tensor cores.  The v2 engine (tc2_gemm.cu) needs its operands pre-split into bf16 planes, which
// costs one extra read + write of every activation map (split_dense / split_colsum passes: ~20 % of a training step,
// profiles/r1b_launches_by_kernel.txt).  Here the fp32 rows are read ONCE, split in registers on their way into the
// canonical SWIZZLE_128B shared-memory tiles, and consumed by tcgen05.mma (3 MMAs per k-step: lo*hi + hi*lo + hi*hi,
// fp32 accumulate in TMEM -- the same fp32-parity scheme as the rest of the engine):
//
//   rows_gemm_kernel   forward / data-gradient:  Y (=|+=) act(sum_kb X_kb[P][64] W[:, 64kb:64kb+64]^T + b)
//                      persistent over 128-row tiles, the (pre-split) weights stay resident in shared memory, the
//                      next tile is prefetched into registers while the current one is in the tensor pipe / epilogue,
//                      the epilogue goes TMEM -> registers -> swizzled staging -> 128-byte coalesced row stores.
//                      Up to three 64-column K blocks may come from different tensors (the channel concatenation of
//                      tsrn.py:902 never exists in memory).
//   rows_wgrad_kernel  weight gradient:  D[64][64*NB] = A[P][64]^T B[P][64*NB] (reduction over the pixel rows, both
//                      operands MN-major: a shared-memory row is one pixel's 64 channels), column sums of A or B (the
//                      bias gradient) accumulated in the loader registers, per-CTA partial tiles + a reduction kernel
//                      (deterministic, no same-address atomics).
//   rows_wgrad_ws_kernel  the same product for large P (the default there): warp-specialised, fp32 slabs by TMA into a
//                      7-buffer shared-memory ring, split IN PLACE into the hi | lo planes, one M = 128 x N = 128 MMA per
//                      k-step ([A_hi ; A_lo] x [B_hi | B_lo]: all four hi/lo products in the accumulator's quadrants).
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>
#include "common.cuh"
#include "gemm_params.cuh"

namespace {

constexpr int RB = 128;            // pixel rows per tile / slab
constexpr int NT = 256;
constexpr int PLANE = RB * 128;    // bytes of one bf16 plane of a [128 rows][64 ch] slab
constexpr int WG_LD = 192;         // row stride of a per-CTA partial tile
constexpr int WG_PART = 64 * WG_LD + WG_LD;   // floats: D[64][192] + colsum[192]

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
// warp-uniform issue: the whole warp runs the (uniform) surrounding code, one elected lane issues -- keeps descriptors in
// uniform registers instead of a per-lane waterfall loop around every UTCHMMA
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
// SWIZZLE_128B shared-memory descriptor (cute::UMMA::SmemDescriptor): version 1 at [46,48), layout type 2 at [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// 4 consecutive fp32 -> 4 bf16 hi (8 B) + 4 bf16 lo (8 B); element 0 at the lowest address
__device__ __forceinline__ uint32_t pack_bf16x2_rn(float lo_elem, float hi_elem) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi_elem), "f"(lo_elem));
  return d;
}
__device__ __forceinline__ void split4(const float4 a, uint2& hi, uint2& lo) {
  // hi = round-to-nearest bf16, lo = bf16(x - hi); the packed hi pair is unpacked with one shift and one mask
  const uint32_t h0 = pack_bf16x2_rn(a.x, a.y), h1 = pack_bf16x2_rn(a.z, a.w);
  const float l0x = a.x - __uint_as_float(h0 << 16), l0y = a.y - __uint_as_float(h0 & 0xffff0000u);
  const float l1x = a.z - __uint_as_float(h1 << 16), l1y = a.w - __uint_as_float(h1 & 0xffff0000u);
  hi = make_uint2(h0, h1);
  lo = make_uint2(pack_bf16x2_rn(l0x, l0y), pack_bf16x2_rn(l1x, l1y));
}
__device__ __forceinline__ void split8(const float4 a, const float4 b, uint4& hi, uint4& lo) {
  uint2 h0, l0, h1, l1;
  split4(a, h0, l0);
  split4(b, h1, l1);
  hi = make_uint4(h0.x, h0.y, h1.x, h1.y);
  lo = make_uint4(l0.x, l0.y, l1.x, l1.y);
}

// One [128 rows][64 cols] fp32 slab in flight per CTA: thread `tid` owns float4 c4 = tid & 15 (columns 4 c4 .. 4 c4 + 3)
// of rows (tid >> 4) + 16 i, i = 0..7 -- a warp reads two full 256-byte rows per instruction (fully coalesced).
struct Slab {
  float4 v[8];
};
// q = address of the thread's float4 in its first row, step = 16 row strides, rows_left = valid rows from that row on
__device__ __forceinline__ void slab_load(Slab& s, const float* __restrict__ q, long long step, int rows_left) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (16 * i < rows_left) {
      s.v[i] = __ldg(reinterpret_cast<const float4*>(q + i * step));
    } else {
      s.v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}
// split + store into a SWIZZLE_128B tile whose rows are the slab rows (128 B = 64 bf16 per row).  `hi` / `lo` already
// include the thread's offset r0 * 128 + ((c4 / 2) ^ (r0 & 7)) * 16 + (c4 & 1) * 8; rows r0 + 16 i share r0's swizzle, so
// the eight 8-byte stores use immediate offsets.  A warp writes two full rows per instruction (conflict-free).
__device__ __forceinline__ void slab_store(const Slab& s, unsigned char* hi, unsigned char* lo, bool single) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    uint2 h, l;
    split4(s.v[i], h, l);
    *reinterpret_cast<uint2*>(hi + i * 2048) = h;
    if (!single) *reinterpret_cast<uint2*>(lo + i * 2048) = l;
  }
}
__device__ __forceinline__ int slab_st_off(int tid) {
  const int c4 = tid & 15, r0 = tid >> 4;
  return r0 * 128 + (((c4 >> 1) ^ (r0 & 7)) << 4) + ((c4 & 1) << 3);
}
// running column sums of the slabs a thread has loaded (its 4 columns are the same in every slab)
__device__ __forceinline__ void slab_colsum(const Slab& s, float (&c)[4]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    c[0] += s.v[i].x; c[1] += s.v[i].y; c[2] += s.v[i].z; c[3] += s.v[i].w;
  }
}
// one warp polls the mbarrier, everyone else parks on the CTA barrier (spinning threads would steal issue slots)
__device__ __forceinline__ void cta_wait(uint32_t bar, uint32_t parity, int warp) {
  if (warp == 0) mbar_wait(bar, parity);
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------ forward / dgrad
struct RowsP {
  const float *X0, *X1, *X2;   // K blocks (64 columns each)
  long long ldx0, ldx1, ldx2;
  const float* W;              // W[N][K] (wtrans = 0) or W[K][N] (wtrans = 1), row stride ldw
  long long ldw;
  int wtrans;
  const float* bias;
  float* Y;
  long long ldy;
  long long M;
  int flags, ntiles;
};

template <int NB, int KB>
__global__ void __launch_bounds__(NT, (NB == 1 && KB == 1) ? 3 : 2) rows_gemm_kernel(const RowsP p) {
  constexpr int N = 64 * NB;
  constexpr int WT = N * 128;                 // bytes of one K block of one W plane ([N rows][128 B])
  constexpr uint32_t TCOLS = (NB == 1) ? 64u : ((NB == 2) ? 128u : 256u);
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool single = (p.flags & F_BF16) != 0;
  unsigned char* w_hi = smem;
  unsigned char* w_lo = smem + KB * WT;
  unsigned char* a_hi = smem + 2 * KB * WT;
  unsigned char* a_lo = a_hi + PLANE;

  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(TCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // weights: split once, resident for the whole kernel (K-major tiles, row n = output channel n)
  for (int idx = tid; idx < KB * N * 8; idx += NT) {
    const int ch = idx & 7;
    const int rn = idx >> 3;
    const int kb = rn / N, n = rn - kb * N;
    float4 a, b;
    if (!p.wtrans) {
      const float4* q = reinterpret_cast<const float4*>(p.W + (long long)n * p.ldw + kb * 64 + ch * 8);
      a = __ldg(q);
      b = __ldg(q + 1);
    } else {
      const float* q = p.W + (long long)(kb * 64 + ch * 8) * p.ldw + n;
      a = make_float4(__ldg(q), __ldg(q + p.ldw), __ldg(q + 2 * p.ldw), __ldg(q + 3 * p.ldw));
      b = make_float4(__ldg(q + 4 * p.ldw), __ldg(q + 5 * p.ldw), __ldg(q + 6 * p.ldw), __ldg(q + 7 * p.ldw));
    }
    uint4 h, l;
    split8(a, b, h, l);
    const int off = kb * WT + n * 128 + ((ch ^ (n & 7)) << 4);
    *reinterpret_cast<uint4*>(w_hi + off) = h;
    if (!single) *reinterpret_cast<uint4*>(w_lo + off) = l;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t bar_a = smem_u32(&bar);
  constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(RB >> 4) << 24);
  const bool accum = (p.flags & F_ACCUM) != 0, relu = (p.flags & F_RELU) != 0;

  // loader geometry (thread constants)
  const int r0 = tid >> 4, col = (tid & 15) * 4;
  const int st_off = slab_st_off(tid);
  // epilogue geometry: warp = (TMEM lane group lg, 32-column half); the operand stage ([128 rows][2 halves][128 B])
  // doubles as a warp-private transposition buffer, 16-byte chunk c of row r at chunk position c ^ (r & 7)
  const int lg = warp & 3, half = warp >> 2;
  const int erow = lg * 32 + lane;
  unsigned char* stw = a_hi + erow * 256 + half * 128;
  const int swz = (erow & 7) << 4;
  const int rr = lane >> 3, cc = lane & 7;          // copy-out: rows lg*32 + 4 i + rr, float4 column cc
  const unsigned char* str_e = a_hi + (lg * 32 + rr) * 256 + half * 128 + ((cc ^ rr) << 4);        // even i
  const unsigned char* str_o = a_hi + (lg * 32 + rr) * 256 + half * 128 + ((cc ^ (rr + 4)) << 4);  // odd i
  float4 bb[NB];
#pragma unroll
  for (int j = 0; j < NB; ++j)
    bb[j] = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + j * 64 + half * 32 + cc * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
  const uint32_t tacc = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(half * 32);
  const long long ystep = 4 * p.ldy;

  Slab pre;
  int tile = blockIdx.x, kb = 0;
  uint32_t ph = 0;
  bool pending = false;
  if (tile < p.ntiles) {
    const long long row0 = (long long)tile * RB + r0;
    slab_load(pre, p.X0 + row0 * p.ldx0 + col, 16 * p.ldx0, (int)min((long long)RB, p.M - row0));
  }
  while (tile < p.ntiles) {
    if (KB > 1 && pending) {     // MMAs of the previous K block still read the stage
      cta_wait(bar_a, ph, warp);
      ph ^= 1;
      pending = false;
    }
    slab_store(pre, a_hi + st_off, a_lo + st_off, single);
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo);
      const uint32_t bh = smem_u32(w_hi) + (uint32_t)(kb * WT), bl = smem_u32(w_lo) + (uint32_t)(kb * WT);
#pragma unroll
      for (int k16 = 0; k16 < 4; ++k16) {
        const uint32_t ko = k16 * 32;   // 16 bf16 along K inside the 128-byte swizzle atom
        const uint64_t dah = make_desc(ah + ko, 16, 1024), dal = make_desc(al + ko, 16, 1024);
        const uint64_t dbh = make_desc(bh + ko, 16, 1024), dbl = make_desc(bl + ko, 16, 1024);
        const uint32_t acc = (kb > 0 || k16 > 0) ? 1u : 0u;
        if (!single) {
          umma_bf16(tmem_base, dal, dbh, idesc, acc);
          umma_bf16(tmem_base, dah, dbl, idesc, 1u);
          umma_bf16(tmem_base, dah, dbh, idesc, 1u);
        } else {
          umma_bf16(tmem_base, dah, dbh, idesc, acc);
        }
      }
      umma_commit(bar_a);
    }
    pending = true;
    // prefetch the next (tile, K block) into registers: in flight during the MMAs and the epilogue
    int ntile = tile, nkb = kb + 1;
    if (nkb == KB) {
      nkb = 0;
      ntile = tile + gridDim.x;
    }
    if (ntile < p.ntiles) {
      const float* xs = (KB == 1 || nkb == 0) ? p.X0 : ((nkb == 1) ? p.X1 : p.X2);
      const long long ls = (KB == 1 || nkb == 0) ? p.ldx0 : ((nkb == 1) ? p.ldx1 : p.ldx2);
      const long long row0 = (long long)ntile * RB + r0;
      slab_load(pre, xs + row0 * ls + col, 16 * ls, (int)min((long long)RB, p.M - row0));
    }
    if (kb == KB - 1) {
      cta_wait(bar_a, ph, warp);
      ph ^= 1;
      pending = false;
      tc_fence_after();
      const long long erow0 = (long long)tile * RB + lg * 32 + rr;
      const int rows_left = (int)min((long long)RB, p.M - erow0);
      float* yrow = p.Y + erow0 * p.ldy + half * 32 + cc * 4;
#pragma unroll
      for (int j = 0; j < NB; ++j) {
#pragma unroll
        for (int c16 = 0; c16 < 2; ++c16) {
          uint32_t v[16];
          tmem_ld16(tacc + (uint32_t)(j * 64 + 16 * c16), v);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<float4*>(stw + ((((4 * c16 + q) << 4)) ^ swz)) =
                make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                            __uint_as_float(v[4 * q + 3]));
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 o = *reinterpret_cast<const float4*>(((i & 1) ? str_o : str_e) + i * 1024);
          o.x += bb[j].x; o.y += bb[j].y; o.z += bb[j].z; o.w += bb[j].w;
          if (4 * i < rows_left) {
            float4* dst = reinterpret_cast<float4*>(yrow + i * ystep + j * 64);
            if (accum) {
              const float4 old = *dst;
              o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
            }
            if (relu) {
              o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
            }
            *dst = o;
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncthreads();          // stage + accumulator free for the next tile
    }
    tile = ntile;
    kb = nkb;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TCOLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ weight gradient
struct WgP {
  const float* A;              // [M][lda]; columns a_off0..+32 and a_off1..+32 form the 64 "A channels"
  long long lda;
  int a_off0, a_off1;
  const float *B0, *B1, *B2;   // NB slabs of 64 columns
  long long ldb0, ldb1, ldb2;
  long long M;
  int nblk;
  float* partial;              // [gridDim.x][WG_PART]
  int colsum_src;              // 0 none, 1 A (64 sums), 2 B (64 NB sums)
  int flags;
};

// one slab in flight per CTA (the 2-deep variant below spills at NB = 3 and was measured slower there)
template <int NB>
__global__ void __launch_bounds__(NT, (NB == 3) ? 2 : 3) rows_wgrad1_kernel(const WgP p) {
  constexpr uint32_t TCOLS = (NB == 1) ? 64u : ((NB == 2) ? 128u : 256u);
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float red[WG_LD];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool single = (p.flags & F_BF16) != 0;
  // [A hi][A lo][B hi][B lo], one 64-channel MN block each.  The MMA's M is 128: the second 64-channel block of the A
  // descriptors (leading-dimension byte offset = PLANE) aliases the NEXT plane -- finite bf16 data whose products land
  // in accumulator rows 64..127, which are never read.
  unsigned char* a_hi = smem;
  unsigned char* a_lo = smem + PLANE;
  unsigned char* b_hi = smem + 2 * PLANE;
  unsigned char* b_lo = smem + 3 * PLANE;

  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(TCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < 4 * PLANE / 16; i += NT) *reinterpret_cast<uint4*>(smem + i * 16) = make_uint4(0u, 0u, 0u, 0u);
  for (int i = tid; i < WG_LD; i += NT) red[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t bar_a = smem_u32(&bar);
  // M = 128, N = 64, A and B MN-major (bits 15, 16)
  constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) |
                             ((uint32_t)(128 >> 4) << 24);
  float cs[NB][4];
#pragma unroll
  for (int j = 0; j < NB; ++j)
#pragma unroll
    for (int q = 0; q < 4; ++q) cs[j][q] = 0.f;

  const int r0 = tid >> 4, c4 = tid & 15;
  const int st_off = slab_st_off(tid);
  const int acol = (c4 < 8) ? (p.a_off0 + c4 * 4) : (p.a_off1 + (c4 - 8) * 4);
  Slab pre;
  uint32_t ph = 0;
  bool pending = false;
  bool first = true;
  int blk = blockIdx.x;
  if (blk < p.nblk) {
    const long long row0 = (long long)blk * RB + r0;
    slab_load(pre, p.A + row0 * p.lda + acol, 16 * p.lda, (int)min((long long)RB, p.M - row0));
  }
  while (blk < p.nblk) {
    const int nblk_next = blk + gridDim.x;
    const long long row0 = (long long)blk * RB + r0;
    const int rows_left = (int)min((long long)RB, p.M - row0);
#pragma unroll
    for (int s = 0; s <= NB; ++s) {
      if (pending) {           // the previous MMA group still reads A / B
        cta_wait(bar_a, ph, warp);
        ph ^= 1;
        pending = false;
      }
      if (s == 0) {
        slab_store(pre, a_hi + st_off, a_lo + st_off, single);
        if (p.colsum_src == 1) slab_colsum(pre, cs[0]);
      } else {
        slab_store(pre, b_hi + st_off, b_lo + st_off, single);
        if (p.colsum_src == 2) slab_colsum(pre, cs[(s > 0) ? (s - 1) : 0]);
      }
      fence_proxy_async();
      __syncthreads();
      // prefetch the next slab: B_s of this block, or A of the CTA's next block
      if (s < NB) {
        const float* bs = (s == 0) ? p.B0 : ((s == 1) ? p.B1 : p.B2);
        const long long ls = (s == 0) ? p.ldb0 : ((s == 1) ? p.ldb1 : p.ldb2);
        slab_load(pre, bs + row0 * ls + c4 * 4, 16 * ls, rows_left);
      } else if (nblk_next < p.nblk) {
        const long long nrow0 = (long long)nblk_next * RB + r0;
        slab_load(pre, p.A + nrow0 * p.lda + acol, 16 * p.lda, (int)min((long long)RB, p.M - nrow0));
      }
      if (s >= 1) {
        if (tid == 0) {
          tc_fence_after();
          const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
          const uint32_t tm = tmem_base + (uint32_t)((s - 1) * 64);
#pragma unroll
          for (int k16 = 0; k16 < 8; ++k16) {
            const uint32_t ko = k16 * 2048;   // 16 pixel rows = two 8-row groups of 1024 B
            const uint64_t dah = make_desc(ah + ko, PLANE, 1024), dal = make_desc(al + ko, PLANE, 1024);
            const uint64_t dbh = make_desc(bh + ko, PLANE, 1024), dbl = make_desc(bl + ko, PLANE, 1024);
            const uint32_t acc = (!first || k16 > 0) ? 1u : 0u;
            if (!single) {
              umma_bf16(tm, dal, dbh, idesc, acc);
              umma_bf16(tm, dah, dbl, idesc, 1u);
              umma_bf16(tm, dah, dbh, idesc, 1u);
            } else {
              umma_bf16(tm, dah, dbh, idesc, acc);
            }
          }
          umma_commit(bar_a);
        }
        pending = true;
      }
    }
    first = false;
    blk = nblk_next;
  }
  if (pending) {
    cta_wait(bar_a, ph, warp);
    ph ^= 1;
  }
  tc_fence_after();
  float* part = p.partial + (long long)blockIdx.x * WG_PART;
  const int lg = warp & 3, half = warp >> 2;
  if (lg < 2) {                      // accumulator rows 64..127 are padding
    const int row = lg * 32 + lane;
#pragma unroll
    for (int j = 0; j < NB; ++j) {
#pragma unroll
      for (int c16 = 0; c16 < 2; ++c16) {
        uint32_t v[16];
        const int col0 = j * 64 + half * 32 + c16 * 16;
        tmem_ld16(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)col0, v);
        float4* dst = reinterpret_cast<float4*>(part + row * WG_LD + col0);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          dst[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                               __uint_as_float(v[4 * q + 3]));
      }
    }
  }
  if (p.colsum_src) {
#pragma unroll
    for (int j = 0; j < NB; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (j == 0 || p.colsum_src == 2) atomicAdd(&red[j * 64 + c4 * 4 + q], cs[j][q]);
    __syncthreads();
    for (int i = tid; i < WG_LD; i += NT) part[64 * WG_LD + i] = red[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TCOLS) : "memory");
  }
}

template <int NB>
__global__ void __launch_bounds__(NT, 2) rows_wgrad_kernel(const WgP p) {
  constexpr uint32_t TCOLS = (NB == 1) ? 64u : ((NB == 2) ? 128u : 256u);
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float red[WG_LD];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool single = (p.flags & F_BF16) != 0;
  // [A hi][A lo][B hi][B lo], one 64-channel MN block each.  The MMA's M is 128: the second 64-channel block of the A
  // descriptors (leading-dimension byte offset = PLANE) aliases the NEXT plane -- finite bf16 data whose products land
  // in accumulator rows 64..127, which are never read.
  unsigned char* a_hi = smem;
  unsigned char* a_lo = smem + PLANE;
  unsigned char* b_hi = smem + 2 * PLANE;
  unsigned char* b_lo = smem + 3 * PLANE;

  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(TCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < 4 * PLANE / 16; i += NT) *reinterpret_cast<uint4*>(smem + i * 16) = make_uint4(0u, 0u, 0u, 0u);
  for (int i = tid; i < WG_LD; i += NT) red[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t bar_a = smem_u32(&bar);
  // M = 128, N = 64, A and B MN-major (bits 15, 16)
  constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) |
                             ((uint32_t)(128 >> 4) << 24);
  float cs[NB][4];
#pragma unroll
  for (int j = 0; j < NB; ++j)
#pragma unroll
    for (int q = 0; q < 4; ++q) cs[j][q] = 0.f;

  const int r0 = tid >> 4, c4 = tid & 15;
  const int st_off = slab_st_off(tid);
  const int acol = (c4 < 8) ? (p.a_off0 + c4 * 4) : (p.a_off1 + (c4 - 8) * 4);
  // The slab stream of this CTA (per 128-row block: A, B_0 .. B_{NB-1}) is double-buffered in registers with a prefetch
  // distance of TWO slabs (64 KB in flight per CTA: with ~2.8 us of loaded HBM latency one slab per CTA caps the kernel
  // at ~3.4 TB/s).  Blocks are processed in pairs so that the buffer of every stream position is a compile-time choice.
  Slab buf0, buf1;
  uint32_t ph = 0;
  bool pending = false;
  bool first = true;
  const int g = gridDim.x;
  int blkA = blockIdx.x;
  auto sload = [&](Slab& d, int jj) {          // stream position jj relative to the current pair (may reach into the next)
    int blk = blkA;
    if (jj >= 2 * (NB + 1)) {
      blk += 2 * g;
      jj -= 2 * (NB + 1);
    }
    if (jj > NB) {
      blk += g;
      jj -= NB + 1;
    }
    const long long row0 = (long long)blk * RB + r0;
    const int rows_left = (blk < p.nblk) ? (int)min((long long)RB, p.M - row0) : 0;
    if (jj == 0) {
      slab_load(d, p.A + row0 * p.lda + acol, 16 * p.lda, rows_left);
    } else {
      const float* bs = (jj == 1) ? p.B0 : ((jj == 2) ? p.B1 : p.B2);
      const long long ls = (jj == 1) ? p.ldb0 : ((jj == 2) ? p.ldb1 : p.ldb2);
      slab_load(d, bs + row0 * ls + c4 * 4, 16 * ls, rows_left);
    }
  };
  sload(buf0, 0);
  sload(buf1, 1);
  while (blkA < p.nblk) {
#pragma unroll
    for (int j = 0; j < 2 * (NB + 1); ++j) {
      const int s = j % (NB + 1);
      Slab& cur = (j & 1) ? buf1 : buf0;
      if (pending) {           // the previous MMA group still reads A / B
        cta_wait(bar_a, ph, warp);
        ph ^= 1;
        pending = false;
      }
      if (s == 0) {
        slab_store(cur, a_hi + st_off, a_lo + st_off, single);
        if (p.colsum_src == 1) slab_colsum(cur, cs[0]);
      } else {
        slab_store(cur, b_hi + st_off, b_lo + st_off, single);
        if (p.colsum_src == 2) slab_colsum(cur, cs[(s > 0) ? (s - 1) : 0]);
      }
      fence_proxy_async();
      __syncthreads();
      sload(cur, j + 2);       // two positions ahead, into the buffer just drained
      if (s >= 1) {
        if (warp == 0) {
          tc_fence_after();
          const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
          const uint32_t tm = tmem_base + (uint32_t)((s - 1) * 64);
#pragma unroll
          for (int k16 = 0; k16 < 8; ++k16) {
            const uint32_t ko = k16 * 2048;   // 16 pixel rows = two 8-row groups of 1024 B
            const uint64_t dah = make_desc(ah + ko, PLANE, 1024), dal = make_desc(al + ko, PLANE, 1024);
            const uint64_t dbh = make_desc(bh + ko, PLANE, 1024), dbl = make_desc(bl + ko, PLANE, 1024);
            const uint32_t acc = ((first && j <= NB) && k16 == 0) ? 0u : 1u;
            if (!single) {
              umma_bf16_elect(tm, dal, dbh, idesc, acc);
              umma_bf16_elect(tm, dah, dbl, idesc, 1u);
              umma_bf16_elect(tm, dah, dbh, idesc, 1u);
            } else {
              umma_bf16_elect(tm, dah, dbh, idesc, acc);
            }
          }
          umma_commit_elect(bar_a);
        }
        pending = true;
      }
    }
    first = false;
    blkA += 2 * g;
  }
  if (pending) {
    cta_wait(bar_a, ph, warp);
    ph ^= 1;
  }
  tc_fence_after();
  float* part = p.partial + (long long)blockIdx.x * WG_PART;
  const int lg = warp & 3, half = warp >> 2;
  if (lg < 2) {                      // accumulator rows 64..127 are padding
    const int row = lg * 32 + lane;
#pragma unroll
    for (int j = 0; j < NB; ++j) {
#pragma unroll
      for (int c16 = 0; c16 < 2; ++c16) {
        uint32_t v[16];
        const int col0 = j * 64 + half * 32 + c16 * 16;
        tmem_ld16(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)col0, v);
        float4* dst = reinterpret_cast<float4*>(part + row * WG_LD + col0);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          dst[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                               __uint_as_float(v[4 * q + 3]));
      }
    }
  }
  if (p.colsum_src) {
#pragma unroll
    for (int j = 0; j < NB; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (j == 0 || p.colsum_src == 2) atomicAdd(&red[j * 64 + c4 * 4 + q], cs[j][q]);
    __syncthreads();
    for (int i = tid; i < WG_LD; i += NT) part[64 * WG_LD + i] = red[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TCOLS) : "memory");
  }
}

// Warp-specialised variant (the default for large M): ONE CTA per SM = 8 converter warps + 1 MMA warp + 1 TMA warp.
//   * The two kernels above prefetch through registers.  All LDGs of a warp share ONE scoreboard (checked in the SASS:
//     every LDG of the 4-deep register ring of an earlier version of this kernel carried write-barrier 5), so waiting
//     for the oldest slab waits for the newest one too: the prefetch depth is 1 whatever the source says, and both
//     kernels sit at 3.0 - 3.7 TB/s = what <= 64 KB per SM in flight buys at the loaded HBM latency.
//   * Here the fp32 slabs arrive by TMA (completion on mbarriers, any depth) into a ring of SEVEN 32 KB buffers.  A
//     buffer goes  empty -> [TMA] fp32 slab [128][64] -> [converters, IN PLACE: all 256 threads read their 8 float4, one
//     named barrier, then write the bf16 hi | lo planes (2 x 16 KB) over them] -> [MMA warp] -> tcgen05.commit -> empty.
//     Typically 3 - 4 buffers are filling (96 - 128 KB per SM in flight), one is being converted, two or three (the
//     block's A slab + the B slabs in the tensor pipe) are being read.  Rows past M are zero-filled by the TMA unit.
//   * ONE MMA per k-step instead of three: the second 64-channel block of both descriptors (leading-dimension byte
//     offset = PLANE) is the LO plane of the same slab, so a single M = 128, N = 128 MMA [A_hi ; A_lo] x [B_hi | B_lo]
//     leaves hi*hi, hi*lo, lo*hi and lo*lo in the four 64 x 64 quadrants of the accumulator (8 KB of operand fetch per
//     k-step instead of 3 x 6 KB; these MN-major MMAs are bound by shared-memory operand fetch, ~100 cycles each).  The
//     epilogue adds the quadrants.  (The older kernels issue three M = 128, N = 64 MMAs and throw the upper halves away.)
//   Stream order per CTA: for each of its 128-row blocks k: A_k, B0_k .. B(NB-1)_k; slab n lives in buffer n % 7.
constexpr int WS_CONV_WARPS = 8;
constexpr int WS_NT = 32 * (WS_CONV_WARPS + 2);
constexpr int WS_NBUF = 7;
constexpr int WS_BUF = 2 * PLANE;               // 32 KB: one fp32 slab = its two bf16 planes
constexpr int WS_SMEM = WS_NBUF * WS_BUF + 1024;

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts64(uint32_t addr, uint2 v) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(v.x), "r"(v.y) : "memory");
}

struct WsMaps {
  CUtensorMap a0, a1;          // A columns a_off0..+32 / a_off1..+32: [M rows][32] fp32, box [128][32]
  CUtensorMap b[3];            // B slabs: [M rows][64] fp32, box [128][64]
};

template <int NB>
__global__ void __launch_bounds__(WS_NT, 1) rows_wgrad_ws_kernel(const __grid_constant__ WsMaps tm, const WgP p) {
  constexpr uint32_t TCOLS = (NB == 1) ? 128u : ((NB == 2) ? 256u : 512u);   // 128 accumulator columns per B slab
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) unsigned long long full_raw[WS_NBUF], full_pl[WS_NBUF], empty[WS_NBUF], done;
  __shared__ uint32_t tmem_base_s;
  __shared__ float red[WG_LD];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool single = (p.flags & F_BF16) != 0;
  const uint32_t sbase = smem_u32(smem);

  if (tid == 0) {
    for (int i = 0; i < WS_NBUF; ++i) {
      mbar_init(smem_u32(&full_raw[i]), 1);
      mbar_init(smem_u32(&full_pl[i]), 32 * WS_CONV_WARPS);
      mbar_init(smem_u32(&empty[i]), 1);
    }
    mbar_init(smem_u32(&done), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (warp == WS_CONV_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(TCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < WG_LD; i += WS_NT) red[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int g = gridDim.x;
  const int nk = (p.nblk - (int)blockIdx.x + g - 1) / g;     // blocks of this CTA: blockIdx.x + k g
  float cs[NB][4];
#pragma unroll
  for (int j = 0; j < NB; ++j)
#pragma unroll
    for (int q = 0; q < 4; ++q) cs[j][q] = 0.f;
  const int c4 = tid & 15;

  if (warp < WS_CONV_WARPS) {
    // ------------------------------------------------------------ converters: fp32 slab -> bf16 hi | lo planes, in place
    const int r0 = tid >> 4;
    const uint32_t st_off = (uint32_t)slab_st_off(tid);
    const uint32_t src_a = (uint32_t)((c4 >> 3) * PLANE + r0 * 128 + (c4 & 7) * 16);   // two [128][32] boxes
    const uint32_t src_b = (uint32_t)(r0 * 256 + c4 * 16);                              // one [128][64] box
    int b = 0;
    uint32_t par = 0;
    for (int k = 0; k < nk; ++k) {
#pragma unroll
      for (int sj = 0; sj <= NB; ++sj) {
        const uint32_t buf = sbase + (uint32_t)(b * WS_BUF);
        mbar_wait(smem_u32(&full_raw[b]), par);
        Slab cur;
        if (sj == 0) {
#pragma unroll
          for (int i = 0; i < 8; ++i) cur.v[i] = lds128(buf + src_a + i * 2048);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) cur.v[i] = lds128(buf + src_b + i * 4096);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");   // every converter holds its part: the planes may overwrite the slab
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          uint2 h, l;
          split4(cur.v[i], h, l);
          sts64(buf + st_off + i * 2048, h);
          if (!single) sts64(buf + PLANE + st_off + i * 2048, l);
        }
        fence_proxy_async();
        mbar_arrive(smem_u32(&full_pl[b]));
        if (sj == 0) {
          if (p.colsum_src == 1) slab_colsum(cur, cs[0]);
        } else {
          if (p.colsum_src == 2) slab_colsum(cur, cs[(sj > 0) ? (sj - 1) : 0]);
        }
        if (++b == WS_NBUF) {
          b = 0;
          par ^= 1;
        }
      }
    }
  } else if (warp == WS_CONV_WARPS) {
    // ------------------------------------------------------------ MMA issuer (whole warp, one elected lane issues)
    // M = 128, N = 128 (bf16 mode: 64, the lo planes do not exist), A and B MN-major (bits 15, 16)
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                           ((uint32_t)((single ? 64 : 128) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    int b = 0;
    uint32_t par = 0;
    for (int k = 0; k < nk; ++k) {
      const int ba = b;
      mbar_wait(smem_u32(&full_pl[ba]), par);
      const uint32_t ah = sbase + (uint32_t)(ba * WS_BUF);
      if (++b == WS_NBUF) {
        b = 0;
        par ^= 1;
      }
#pragma unroll
      for (int sj = 0; sj < NB; ++sj) {
        mbar_wait(smem_u32(&full_pl[b]), par);
        tc_fence_after();
        const uint32_t bh = sbase + (uint32_t)(b * WS_BUF);
        const uint32_t tmc = tmem_base + (uint32_t)(sj * 128);
#pragma unroll
        for (int k16 = 0; k16 < 8; ++k16) {
          const uint32_t ko = k16 * 2048;   // 16 pixel rows = two 8-row groups of 1024 B
          umma_bf16_elect(tmc, make_desc(ah + ko, PLANE, 1024), make_desc(bh + ko, PLANE, 1024), idesc,
                          (k == 0 && k16 == 0) ? 0u : 1u);
        }
        umma_commit_elect(smem_u32(&empty[b]));
        if (++b == WS_NBUF) {
          b = 0;
          par ^= 1;
        }
      }
      umma_commit_elect(smem_u32(&empty[ba]));
    }
    umma_commit_elect(smem_u32(&done));
  } else {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int b = 0;
      uint32_t par = 0;
      bool wrapped = false;
      for (int k = 0; k < nk; ++k) {
        const int row0 = ((int)blockIdx.x + k * g) * RB;
#pragma unroll
        for (int sj = 0; sj <= NB; ++sj) {
          if (wrapped) mbar_wait(smem_u32(&empty[b]), par ^ 1u);
          const uint32_t buf = sbase + (uint32_t)(b * WS_BUF), bar = smem_u32(&full_raw[b]);
          mbar_expect_tx(bar, (uint32_t)WS_BUF);
          if (sj == 0) {
            tma_load_2d(buf, &tm.a0, bar, 0, row0);
            tma_load_2d(buf + PLANE, &tm.a1, bar, 0, row0);
          } else {
            tma_load_2d(buf, &tm.b[sj - 1], bar, 0, row0);
          }
          if (++b == WS_NBUF) {
            b = 0;
            par ^= 1;
            wrapped = true;
          }
        }
      }
    }
  }
  mbar_wait(smem_u32(&done), 0u);
  tc_fence_after();
  float* part = p.partial + (long long)blockIdx.x * WG_PART;
  const int lg = warp & 3, half = warp >> 2;
  if (warp < WS_CONV_WARPS) {
    // rows 64..127 (the A_lo products; TMEM lane groups 2, 3) go through shared memory (the ring is dead) to the warps
    // that own rows 0..63.  bf16 mode: the lo planes were never written (stale slab bytes), rows 64..127 are ignored.
    constexpr int SLD = 196;                      // floats per staged row: 16-byte aligned, rows 4 banks apart
    float* stg = reinterpret_cast<float*>(smem);
    const int row = (lg & 1) * 32 + lane;
    uint32_t v[NB][2][16];
#pragma unroll
    for (int j = 0; j < NB; ++j)
#pragma unroll
      for (int c16 = 0; c16 < 2; ++c16) {
        const uint32_t tcol = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(j * 128 + half * 32 + c16 * 16);
        tmem_ld16(tcol, v[j][c16]);
        if (!single) {                              // + the B_lo column block
          uint32_t v2[16];
          tmem_ld16(tcol + 64, v2);
#pragma unroll
          for (int q = 0; q < 16; ++q) v[j][c16][q] = __float_as_uint(__uint_as_float(v[j][c16][q]) + __uint_as_float(v2[q]));
        }
      }
    if (lg >= 2) {
#pragma unroll
      for (int j = 0; j < NB; ++j)
#pragma unroll
        for (int c16 = 0; c16 < 2; ++c16) {
          float4* dst = reinterpret_cast<float4*>(stg + row * SLD + j * 64 + half * 32 + c16 * 16);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            dst[q] = make_float4(__uint_as_float(v[j][c16][4 * q]), __uint_as_float(v[j][c16][4 * q + 1]),
                                 __uint_as_float(v[j][c16][4 * q + 2]), __uint_as_float(v[j][c16][4 * q + 3]));
        }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");   // the eight converter warps
    if (lg < 2) {
#pragma unroll
      for (int j = 0; j < NB; ++j)
#pragma unroll
        for (int c16 = 0; c16 < 2; ++c16) {
          const int col0 = j * 64 + half * 32 + c16 * 16;
          const float4* lo4 = reinterpret_cast<const float4*>(stg + row * SLD + col0);
          float4* dst = reinterpret_cast<float4*>(part + row * WG_LD + col0);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float4 l = lo4[q];
            if (single) l = make_float4(0.f, 0.f, 0.f, 0.f);
            dst[q] = make_float4(__uint_as_float(v[j][c16][4 * q]) + l.x, __uint_as_float(v[j][c16][4 * q + 1]) + l.y,
                                 __uint_as_float(v[j][c16][4 * q + 2]) + l.z, __uint_as_float(v[j][c16][4 * q + 3]) + l.w);
          }
        }
    }
  }
  if (p.colsum_src) {
    if (warp < WS_CONV_WARPS) {
#pragma unroll
      for (int j = 0; j < NB; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (j == 0 || p.colsum_src == 2) atomicAdd(&red[j * 64 + c4 * 4 + q], cs[j][q]);
    }
    __syncthreads();
    for (int i = tid; i < WG_LD; i += WS_NT) part[64 * WG_LD + i] = red[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == WS_CONV_WARPS) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TCOLS) : "memory");
  }
}

// D[r][c] = sum over the per-CTA partial tiles (coalesced reads, 16 part groups per element, fixed summation order),
// then scattered to out[b][i][j] with D[b rb + (T ? j : i)][b cb + (T ? i : j)];  rows >= 64 of the index space are the
// column sums -> dbias
__global__ void __launch_bounds__(1024) rows_wgrad_reduce_kernel(const float* __restrict__ partial, int nparts,
                                                                float* __restrict__ out, int nb, int ni, int nj, int T,
                                                                int rb, int cb, float* __restrict__ dbias, int nbias,
                                                                int ncols) {
  // 16 part groups per element (a 4-group version spent 11 us per call on ~40 dependent-latency loads per thread)
  __shared__ float sm[16][64];
  const int e = threadIdx.x & 63, g = threadIdx.x >> 6;
  const int row = blockIdx.x / 3, c = (blockIdx.x % 3) * 64 + e;       // row 64 = the column sums
  float a = 0.f;
  if (c < ncols) {
    const float* src = partial + (long long)row * WG_LD + c;
#pragma unroll 4
    for (int q = g; q < nparts; q += 16) a += src[(long long)q * WG_PART];
  }
  sm[g][e] = a;
  __syncthreads();
  if (g != 0 || c >= ncols) return;
  a = 0.f;
#pragma unroll
  for (int q = 0; q < 16; q += 4) a += (sm[q][e] + sm[q + 1][e]) + (sm[q + 2][e] + sm[q + 3][e]);
  if (row == 64) {
    if (c < nbias) dbias[c] = a;
    return;
  }
  for (int b = 0; b < nb; ++b) {
    const int rr = row - b * rb, cc = c - b * cb;
    const int i = T ? cc : rr, j = T ? rr : cc;
    if (i >= 0 && i < ni && j >= 0 && j < nj) out[((long long)b * ni + i) * nj + j] = a;
  }
}

static int num_sms() {
  int dev = 0, n = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n > 0 ? n : 148;
}
static bool al16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

template <int NB, int KB>
static int launch_rows_gemm(const RowsP& p, cudaStream_t st) {
  const int smem = 2 * KB * NB * 64 * 128 + 2 * PLANE + 1024;
  int grid = ((NB == 1 && KB == 1) ? 3 : 2) * num_sms();
  if (grid > p.ntiles) grid = p.ntiles;
  TATT_CUDA(cudaFuncSetAttribute(rows_gemm_kernel<NB, KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  rows_gemm_kernel<NB, KB><<<grid, NT, smem, st>>>(p);
  TATT_LAUNCH_CHECK("rows_gemm_kernel");
  return 0;
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
    if (q != cudaDriverEntryPointSuccess) return nullptr;
    return (EncodeTiledFn)f;
  }();
  return fn;
}
// fp32 [M rows][cols] view with row stride ld (elements), box [128 rows][cols]; rows past M read as zeros
static bool encode_slab_map(CUtensorMap* m, const float* base, long long ld, long long M, int cols) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)M};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)cols, (cuuint32_t)RB};
  cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// -1: the TMA path cannot serve this call (no driver entry point / descriptor rejected): the caller falls back
template <int NB>
static int launch_rows_wgrad_ws(const WgP& p, int grid, cudaStream_t st) {
  WsMaps tm;
  if (!encode_slab_map(&tm.a0, p.A + p.a_off0, p.lda, p.M, 32) || !encode_slab_map(&tm.a1, p.A + p.a_off1, p.lda, p.M, 32))
    return -1;
  const float* bs[3] = {p.B0, p.B1, p.B2};
  const long long ls[3] = {p.ldb0, p.ldb1, p.ldb2};
  for (int j = 0; j < 3; ++j)
    if (!encode_slab_map(&tm.b[j], bs[j < NB ? j : 0], ls[j < NB ? j : 0], p.M, 64)) return -1;
  TATT_CUDA(cudaFuncSetAttribute(rows_wgrad_ws_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM));
  rows_wgrad_ws_kernel<NB><<<grid, WS_NT, WS_SMEM, st>>>(tm, p);
  TATT_LAUNCH_CHECK("rows_wgrad_ws_kernel");
  return 0;
}
template <int NB>
static int launch_rows_wgrad(const WgP& p, int grid, cudaStream_t st) {
  const int smem = 4 * PLANE + 1024;
  if (NB == 3) {
    TATT_CUDA(cudaFuncSetAttribute(rows_wgrad1_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    rows_wgrad1_kernel<NB><<<grid, NT, smem, st>>>(p);
    TATT_LAUNCH_CHECK("rows_wgrad1_kernel");
    return 0;
  }
  TATT_CUDA(cudaFuncSetAttribute(rows_wgrad_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  rows_wgrad_kernel<NB><<<grid, NT, smem, st>>>(p);
  TATT_LAUNCH_CHECK("rows_wgrad_kernel");
  return 0;
}
constexpr int WG_MAX_GRID = 3 * 148;

}  // namespace

extern "C" {

int tatt_rows_gemm(const float* X0, long long ldx0, const float* X1, long long ldx1, const float* X2, long long ldx2,
                   const float* W, long long ldw, int wtrans, const float* bias, float* Y, long long ldy, long long M,
                   int N, int KB, int flags, void* stream) {
  TATT_REQUIRE(M >= 1 && (N == 64 || N == 128 || N == 192) && KB >= 1 && KB <= 3 && (KB == 1 || N == 64),
               "tatt_rows_gemm: unsupported shape M=%lld N=%d K=%d", M, N, 64 * KB);
  TATT_REQUIRE(X0 && W && Y && (KB < 2 || X1) && (KB < 3 || X2), "tatt_rows_gemm: null operand");
  TATT_REQUIRE(al16(X0) && al16(X1) && al16(X2) && al16(Y) && al16(bias) && (wtrans || al16(W)) && ldx0 % 4 == 0 &&
                   ldx1 % 4 == 0 && ldx2 % 4 == 0 && ldy % 4 == 0 && (wtrans || ldw % 4 == 0),
               "tatt_rows_gemm: operands must be 16-byte aligned with row strides that are multiples of 4");
  RowsP p;
  p.X0 = X0; p.X1 = X1; p.X2 = X2;
  p.ldx0 = ldx0; p.ldx1 = ldx1; p.ldx2 = ldx2;
  p.W = W; p.ldw = ldw; p.wtrans = wtrans;
  p.bias = bias; p.Y = Y; p.ldy = ldy;
  p.M = M; p.flags = flags;
  p.ntiles = (int)((M + RB - 1) / RB);
  cudaStream_t st = (cudaStream_t)stream;
  if (KB == 1) {
    if (N == 64) return launch_rows_gemm<1, 1>(p, st);
    if (N == 128) return launch_rows_gemm<2, 1>(p, st);
    return launch_rows_gemm<3, 1>(p, st);
  }
  if (KB == 2) return launch_rows_gemm<1, 2>(p, st);
  return launch_rows_gemm<1, 3>(p, st);
}

int tatt_rows_wgrad_ws_bytes(void) { return WG_MAX_GRID * WG_PART * (int)sizeof(float) + 256; }

int tatt_rows_wgrad(const float* A, long long lda, int a_off0, int a_off1, const float* B0, long long ldb0,
                    const float* B1, long long ldb1, const float* B2, long long ldb2, int NB, long long M,
                    int colsum_src, float* out, int nb, int ni, int nj, int transpose, int rb, int cb, float* dbias,
                    int nbias, void* ws, long long ws_bytes, int flags, void* stream) {
  TATT_REQUIRE(M >= 1 && NB >= 1 && NB <= 3, "tatt_rows_wgrad: unsupported shape M=%lld NB=%d", M, NB);
  TATT_REQUIRE(A && B0 && (NB < 2 || B1) && (NB < 3 || B2) && out && ws, "tatt_rows_wgrad: null operand");
  TATT_REQUIRE(al16(A) && al16(B0) && al16(B1) && al16(B2) && al16(ws) && lda % 4 == 0 && ldb0 % 4 == 0 &&
                   ldb1 % 4 == 0 && ldb2 % 4 == 0 && a_off0 % 4 == 0 && a_off1 % 4 == 0,
               "tatt_rows_wgrad: operands must be 16-byte aligned with strides / offsets that are multiples of 4");
  TATT_REQUIRE(colsum_src >= 0 && colsum_src <= 2 && (colsum_src == 0 || dbias) && nbias >= 0 &&
                   nbias <= ((colsum_src == 1) ? 64 : 64 * NB),
               "tatt_rows_wgrad: bad column-sum request");
  {
    const int rmax = (nb - 1) * rb + (transpose ? nj : ni), cmax = (nb - 1) * cb + (transpose ? ni : nj);
    TATT_REQUIRE(nb >= 1 && ni >= 1 && nj >= 1 && rmax <= 64 && cmax <= 64 * NB,
                 "tatt_rows_wgrad: output map [%d][%d][%d] exceeds the [64][%d] product", nb, ni, nj, 64 * NB);
  }
  WgP p;
  p.A = A; p.lda = lda; p.a_off0 = a_off0; p.a_off1 = a_off1;
  p.B0 = B0; p.B1 = B1; p.B2 = B2;
  p.ldb0 = ldb0; p.ldb1 = ldb1; p.ldb2 = ldb2;
  p.M = M;
  p.nblk = (int)((M + RB - 1) / RB);
  p.colsum_src = colsum_src; p.flags = flags;
  // TATT_WG_WS=0: the register-prefetch kernels at two CTAs per SM (A/B timing); the warp-specialised TMA kernel runs one
  // CTA per SM and needs a few blocks per CTA to fill its ring
  static const int ws_on = []() {
    const char* e = getenv("TATT_WG_WS");
    return e ? atoi(e) : 1;
  }();
  cudaStream_t st = (cudaStream_t)stream;
  p.partial = reinterpret_cast<float*>(ws);
  int rc = -1, nparts = 0;
  if (ws_on && p.nblk >= 2 * num_sms()) {
    const int grid = num_sms() < WG_MAX_GRID ? num_sms() : WG_MAX_GRID;
    TATT_REQUIRE((long long)grid * WG_PART * (long long)sizeof(float) <= ws_bytes,
                 "tatt_rows_wgrad: workspace too small (%lld bytes, need %lld)", ws_bytes,
                 (long long)grid * WG_PART * (long long)sizeof(float));
    rc = (NB == 1) ? launch_rows_wgrad_ws<1>(p, grid, st)
                   : ((NB == 2) ? launch_rows_wgrad_ws<2>(p, grid, st) : launch_rows_wgrad_ws<3>(p, grid, st));
    nparts = grid;
  }
  if (rc < 0) {
    int grid = 2 * num_sms();
    if (grid > WG_MAX_GRID) grid = WG_MAX_GRID;
    if (grid > p.nblk) grid = p.nblk;
    TATT_REQUIRE((long long)grid * WG_PART * (long long)sizeof(float) <= ws_bytes,
                 "tatt_rows_wgrad: workspace too small (%lld bytes, need %lld)", ws_bytes,
                 (long long)grid * WG_PART * (long long)sizeof(float));
    rc = (NB == 1) ? launch_rows_wgrad<1>(p, grid, st)
                   : ((NB == 2) ? launch_rows_wgrad<2>(p, grid, st) : launch_rows_wgrad<3>(p, grid, st));
    nparts = grid;
  }
  if (rc) return rc;
  rows_wgrad_reduce_kernel<<<(64 + (colsum_src ? 1 : 0)) * 3, 1024, 0, st>>>(
      p.partial, nparts, out, nb, ni, nj, transpose, rb, cb, dbias, colsum_src ? nbias : 0, 64 * NB);
  TATT_LAUNCH_CHECK("rows_wgrad_reduce_kernel");
  return 0;
}

}  // extern "C"
