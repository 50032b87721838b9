// 3x3 / 64->64 convolution (stride 1, pad 1, NHWC) on tcgen05 with TMA-staged halo tiles.
//
// The implicit-GEMM kernels of tc2_gemm.cu re-read the activation tile once per tap (9x) and the weight tile once
// per CTA per tap: 885 MB of L2->SMEM traffic per conv at the bench shape, which bounds them (ncu: tensor pipe 25 %,
// profiles/r1_ncu_full_tc2_conv3x3.csv).  Here one CTA owns R output rows x 128 output columns of one image:
//   * ONE TMA box load per bf16 plane (hi, lo) brings the (R+2) x 130 pixel halo x 64 channels into shared memory
//     (out-of-image coordinates are zero-filled by the TMA unit = the conv padding), 128B-swizzled;
//   * the A operand of tap (dy,dx) for output row r is simply the 128 consecutive halo pixels starting at
//     (r+dy+1)*130 + (dx+1): a shifted UMMA descriptor into the same tile -- no re-load, no im2col address math;
//   * the 9 weight taps ([64 co][64 ci] bf16 hi+lo = 16 KB each) stream through a 3-stage TMA ring;
//   * warp-specialised: warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer (3 MMAs per k-step: lo*hi, hi*lo,
//     hi*hi into fp32 TMEM accumulators, one 64-column accumulator per output row), warps 2-5 = epilogue
//     (tcgen05.ld -> +bias -> HBM).
// L2->SMEM traffic per conv drops ~3x (A: (R+2)/R reads per pixel instead of 9; weights once per R rows).
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>
#include "common.cuh"
#include "gemm_params.cuh"

namespace {

constexpr int TW = 128;        // output columns per CTA
constexpr int HALO_W = TW + 2;
constexpr int NSTAGE = 5;      // weight-tap ring (5 x 16 KB: load latency hidden behind 4 taps of MMAs)
constexpr int B_TAP_BYTES = 2 * 64 * 128;   // hi + lo

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
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// K-major SWIZZLE_128B descriptor; `base_off` = (start >> 7) & 7 when the start is not on a 1024-byte boundary
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t base_off) {
  uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int R>
__global__ void __launch_bounds__(192, 1)
conv3x3_tma_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                   const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                   float* __restrict__ Y, const float* __restrict__ bias, int H, int W, int base_off_mode,
                   int single) {
  constexpr int A_ROWS = (R + 2) * HALO_W;
  constexpr int A_PLANE = ((A_ROWS * 128 + 1023) / 1024) * 1024;
  constexpr int TMEM_COLS = (R * 64 <= 64) ? 64 : ((R * 64 <= 128) ? 128 : 256);
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) unsigned long long bar_a, bar_done, bar_full[NSTAGE], bar_empty[NSTAGE];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nxt = W / TW, nyb = H / R;
  const int xb = blockIdx.x % nxt;
  const int yb = (blockIdx.x / nxt) % nyb;
  const int n = blockIdx.x / (nxt * nyb);
  const int x0 = xb * TW, y0 = yb * R;
  const uint32_t a_hi = smem_u32(smem), a_lo = a_hi + A_PLANE, b_ring = a_lo + A_PLANE;

  if (tid == 0) {
    mbar_init(smem_u32(&bar_a), 1);
    mbar_init(smem_u32(&bar_done), 1);
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_expect_tx(smem_u32(&bar_a), (single ? 1u : 2u) * (uint32_t)(A_ROWS * 128));
      tma_load_4d(a_hi, &tmAh, smem_u32(&bar_a), 0, x0 - 1, y0 - 1, n);
      if (!single) tma_load_4d(a_lo, &tmAl, smem_u32(&bar_a), 0, x0 - 1, y0 - 1, n);
      for (int tap = 0; tap < 9; ++tap) {
        const int s = tap % NSTAGE;
        if (tap >= NSTAGE) mbar_wait(smem_u32(&bar_empty[s]), (uint32_t)(((tap / NSTAGE) - 1) & 1));
        const uint32_t dst = b_ring + (uint32_t)(s * B_TAP_BYTES);
        mbar_expect_tx(smem_u32(&bar_full[s]), (uint32_t)(single ? B_TAP_BYTES / 2 : B_TAP_BYTES));
        tma_load_2d(dst, &tmBh, smem_u32(&bar_full[s]), tap * 64, 0);
        if (!single) tma_load_2d(dst + 64 * 128, &tmBl, smem_u32(&bar_full[s]), tap * 64, 0);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      mbar_wait(smem_u32(&bar_a), 0);
      tc_fence_after();
      for (int tap = 0; tap < 9; ++tap) {
        const int s = tap % NSTAGE;
        mbar_wait(smem_u32(&bar_full[s]), (uint32_t)((tap / NSTAGE) & 1));
        tc_fence_after();
        const int dy = tap / 3 - 1, dx = tap % 3 - 1;
        const uint32_t b_hi = b_ring + (uint32_t)(s * B_TAP_BYTES), b_lo = b_hi + 64 * 128;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const uint32_t row0 = (uint32_t)((r + dy + 1) * HALO_W + (dx + 1)) * 128u;
#pragma unroll
          for (int k16 = 0; k16 < 4; ++k16) {
            const uint32_t ko = k16 * 32;
            const uint32_t sa_h = a_hi + row0 + ko, sa_l = a_lo + row0 + ko;
            const uint32_t bo_h = base_off_mode ? ((sa_h >> 7) & 7) : 0, bo_l = base_off_mode ? ((sa_l >> 7) & 7) : 0;
            const uint64_t dah = make_desc(sa_h, bo_h), dal = make_desc(sa_l, bo_l);
            const uint64_t dbh = make_desc(b_hi + ko, 0), dbl = make_desc(b_lo + ko, 0);
            const uint32_t tm = tmem_base + (uint32_t)(r * 64);
            if (!single) {
              umma_bf16(tm, dal, dbh, idesc, (tap > 0 || k16 > 0) ? 1u : 0u);
              umma_bf16(tm, dah, dbl, idesc, 1u);
              umma_bf16(tm, dah, dbh, idesc, 1u);
            } else {
              umma_bf16(tm, dah, dbh, idesc, (tap > 0 || k16 > 0) ? 1u : 0u);
            }
          }
        }
        umma_commit(smem_u32(&bar_empty[s]));
      }
      umma_commit(smem_u32(&bar_done));
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5 -> TMEM lane groups 2,3,0,1)
    mbar_wait(smem_u32(&bar_done), 0);
    tc_fence_after();
    const int lg = warp & 3;
    const int px = x0 + lg * 32 + lane;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float* dst = Y + (((long long)n * H + (y0 + r)) * W + px) * 64;
#pragma unroll
      for (int c16 = 0; c16 < 4; ++c16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(r * 64 + c16 * 16), v);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float4 o;
          o.x = __uint_as_float(v[4 * q + 0]);
          o.y = __uint_as_float(v[4 * q + 1]);
          o.z = __uint_as_float(v[4 * q + 2]);
          o.w = __uint_as_float(v[4 * q + 3]);
          if (bias) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + c16 * 16 + 4 * q));
            o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
          }
          *reinterpret_cast<float4*>(dst + c16 * 16 + 4 * q) = o;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ persistent variant
// conv3x3_tma_kernel above runs one tile per CTA: 6.9 waves of {prologue, halo load, 216 MMAs, epilogue} that do not
// overlap (ncu: tensor pipe 26 %).  The rolling-halo variant below is persistent (one CTA per SM walks down a column of
// 2-row tiles) and keeps the tensor pipe fed:
//   * the 4 halo rows live in a circular buffer of two row PAIRS; consecutive tiles share two rows, so each tile only
//     loads its two NEW rows -- into the pair that went dead after the ky = 0, 1 taps of the previous tile, i.e. the
//     load is issued two thirds through the previous tile's MMAs and waited for just before this tile's ky = 1 taps;
//   * two 128-column TMEM accumulators: the epilogue of tile t overlaps the MMAs of tile t+1;
//   * the weight-tap ring keeps streaming across tiles;
//   * the MMA thread builds descriptors with one integer add each (row bases precomputed per tile).
// A pair is padded to a multiple of 1024 bytes so the TMA box (2 rows) and the UMMA descriptors agree on the swizzle.
constexpr int ROW_BYTES = HALO_W * 128;                       // 16640
constexpr int PAIR_BYTES = ((2 * ROW_BYTES + 1023) / 1024) * 1024;   // 34816
constexpr int RL_A_PLANE = 2 * PAIR_BYTES;

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// K-major SWIZZLE_128B descriptors from their low words (start address >> 4, LBO field 1); high word: SBO 1024 B,
// version 1, layout type 2
// warp-uniform variants: the WHOLE warp executes the surrounding (uniform) control flow and one elected lane issues, so
// ptxas keeps descriptors / addresses in uniform registers instead of emitting a per-lane "waterfall" loop around
// every UTCHMMA (which is what `if (lane == 0) { ... }` around the issue loop compiles to)
__device__ __forceinline__ void umma_lo_elect(uint32_t tmem_c, uint32_t a_lo32, uint32_t b_lo32, uint32_t idesc,
                                              uint32_t acc) {
  constexpr uint32_t HI = (1024u >> 4) | (1u << 14) | (2u << 29);
  asm volatile(
      "{\n\t"
      ".reg .pred p, e;\n\t"
      ".reg .b64 da, db;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t"
      "}" ::"r"(tmem_c),
      "r"(a_lo32), "r"(b_lo32), "r"(idesc), "r"(acc), "r"(HI)
      : "memory");
}
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

constexpr int RL_NSTAGE = 4;                                  // weight-tap ring (4 x 16 KB) + 16 KB of epilogue staging
constexpr int RL_STG_BYTES = 4 * 32 * 128;                    // per epilogue warp: [32 px][32 ch] fp32

template <bool single, bool cat, bool multi_in>
__global__ void __launch_bounds__(192, 1)
conv3x3_roll_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                    const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                    float* __restrict__ Y, const float* __restrict__ bias, int H, int W, int ntiles, int dbg, int ldy,
                    int ngroups, float* __restrict__ stats, int ncin_rt) {
  const int ncin = multi_in ? ncin_rt : 1;      // compile-time 1 for the 64-input-channel layers (the tuned path)
  // C_in = 64 * ncin (ncin > 1 only with one output group): every output tile accumulates ncin x 9 taps in its TMEM
  // accumulator, input group gi = channels [64 gi, 64 gi + 64) of the [P][C_in] planes = k columns tap * C_in + 64 gi of
  // the weight planes.  The halo rows of a group cannot roll into the next tile (the four row slots are reloaded per
  // group), so every (tile, group) is a "fresh" tile; the loads still hide behind the MMAs: rows 0, 1 of the next group
  // are requested after tap 5, rows 2, 3 after tap 8 and are first needed at the next group's tap 3.
  // C_out = 64 * ngroups: work item w = g * ntiles + t is the 64-channel output group g of pixel tile t (the same halo
  // tile is re-read per group, mostly from L2; every group is the 64 -> 64 problem with its own weight slice)
  constexpr uint32_t acc_stride = cat ? 256u : 128u, row_stride = cat ? 128u : 64u;
  constexpr uint32_t tmem_cols = cat ? 512u : 256u;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) unsigned long long pair_full[2], pair_empty[2], w_full[RL_NSTAGE], w_empty[RL_NSTAGE], acc_full[2],
      acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float stat_red[4 * 128];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nyb = H / 2, nxt = W / TW;
  const int t0 = (int)((long long)blockIdx.x * ntiles * ngroups / gridDim.x);
  const int t1 = (int)((long long)(blockIdx.x + 1) * ntiles * ngroups / gridDim.x);
  const uint32_t a_hi = smem_u32(smem), a_lo = a_hi + RL_A_PLANE, b_ring = a_lo + RL_A_PLANE;
  unsigned char* stg_base = smem + 2 * RL_A_PLANE + RL_NSTAGE * B_TAP_BYTES;

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&pair_full[i]), 1);
      mbar_init(smem_u32(&pair_empty[i]), 1);
      mbar_init(smem_u32(&acc_full[i]), 1);
      mbar_init(smem_u32(&acc_empty[i]), 4);
    }
    for (int s = 0; s < RL_NSTAGE; ++s) {
      mbar_init(smem_u32(&w_full[s]), 1);
      mbar_init(smem_u32(&w_empty[s]), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  // cat: the weight tap is ONE [128 x 64] K-major tile (rows 0..63 = hi plane, 64..127 = lo plane, contiguous in the
  // ring), so A_hi x [B_hi | B_lo] is a single N = 128 MMA into two 64-column blocks (hi*hi | hi*lo) and A_lo x B_hi a
  // second, N = 64 one: 14 KB of shared-memory operand fetch per k-step and row instead of 18 KB, 2 issues instead of
  // 3.  The epilogue adds the two column blocks.  Accumulators: 2 rows x 128 columns, ping-pong = all 512 TMEM columns.

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const uint32_t pair_bytes = (single ? 1u : 2u) * (uint32_t)(2 * ROW_BYTES);
      int nload0 = 0, nload1 = 0;                 // loads issued into pair 0 / 1
      int ws = 0;                                 // weight ring stage
      uint32_t wpar = 0;                          // parity of the NEXT wait on w_empty[ws] (valid once the ring wrapped)
      bool wrapped = false;
      for (int w = t0; w < t1; ++w) {
        const int g = w / ntiles, t = w - g * ntiles;
        const int yb = t % nyb, xb = (t / nyb) % nxt, n = t / (nyb * nxt);
        const int y0 = 2 * yb, x0 = xb * TW;
        const bool fresh = (w == t0) || (yb == 0) || (ncin > 1);
        const int p0 = (y0 >> 1) & 1;
        for (int gi = 0; gi < ncin; ++gi) {
        auto load_pair = [&](int k) {
          const int pr = (k == 0) ? p0 : (p0 ^ 1);
          const int nl = pr ? nload1 : nload0;
          if (nl > 0) mbar_wait(smem_u32(&pair_empty[pr]), (uint32_t)((nl - 1) & 1));
          const uint32_t bar = smem_u32(&pair_full[pr]);
          mbar_expect_tx(bar, pair_bytes);
          const int yy = (k == 0) ? (y0 - 1) : (y0 + 1);
          tma_load_4d(a_hi + (uint32_t)(pr * PAIR_BYTES), &tmAh, bar, gi * 64, x0 - 1, yy, n);
          if (!single) tma_load_4d(a_lo + (uint32_t)(pr * PAIR_BYTES), &tmAl, bar, gi * 64, x0 - 1, yy, n);
          if (pr) ++nload1; else ++nload0;
        };
        if (fresh) load_pair(0);                   // else rows y0-1, y0 are already there (previous tile's rows 2, 3)
        // rows y0+1, y0+2 are first read at tap 3.  With input groups their slots are released only when the previous
        // group is done, so the wait sits behind the first three weight taps (else it would hold those back too)
        if (ncin == 1) load_pair(1);
        for (int tap = 0; tap < 9; ++tap) {
          if ((dbg & 2) && (w > t0 || tap >= RL_NSTAGE)) break;
          if (ncin > 1 && tap == 3) load_pair(1);
          if (wrapped) mbar_wait(smem_u32(&w_empty[ws]), wpar);
          const uint32_t dst = b_ring + (uint32_t)(ws * B_TAP_BYTES);
          mbar_expect_tx(smem_u32(&w_full[ws]), (uint32_t)(single ? B_TAP_BYTES / 2 : B_TAP_BYTES));
          tma_load_2d(dst, &tmBh, smem_u32(&w_full[ws]), (tap * ncin + gi) * 64, g * 64);
          if (!single) tma_load_2d(dst + 64 * 128, &tmBl, smem_u32(&w_full[ws]), (tap * ncin + gi) * 64, g * 64);
          if (++ws == RL_NSTAGE) {
            ws = 0;
            if (wrapped) wpar ^= 1;
            wrapped = true;
          }
        }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    {                                             // whole warp, uniform control flow; one elected lane issues
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      constexpr uint32_t idesc128 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      constexpr uint32_t LBO1 = 1u << 16;
      int nfull0 = 0, nfull1 = 0;                 // pair loads consumed
      int ws = 0;
      uint32_t wpar = 0;
      int it = 0;
      for (int w = t0; w < t1; ++w, ++it) {
        const int t = w % ntiles;
        const int yb = t % nyb;
        const int y0 = 2 * yb;
        const bool fresh = (w == t0) || (yb == 0) || (ncin > 1);
        const int p0 = (y0 >> 1) & 1, p1 = p0 ^ 1;
        if (it >= 2) mbar_wait(smem_u32(&acc_empty[it & 1]), (uint32_t)(((it >> 1) - 1) & 1));
        for (int gi = 0; gi < ncin; ++gi) {
        const bool next_fresh = (ncin > 1) ? (gi + 1 < ncin || w + 1 < t1) : ((w + 1 < t1) && (((t + 1) % nyb) == 0));
        if (fresh) {
          mbar_wait(smem_u32(&pair_full[p0]), (uint32_t)((p0 ? nfull1 : nfull0) & 1));
          if (p0) ++nfull1; else ++nfull0;
        }
        // halo row j (0..3) of this tile lives in circular slot (y0 + j) & 3 = pair (slot >> 1), row (slot & 1)
        uint32_t rowh[4], rowl[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int slot = (y0 + j) & 3;
          const uint32_t off = (uint32_t)((slot >> 1) * PAIR_BYTES + (slot & 1) * ROW_BYTES);
          rowh[j] = ((a_hi + off) >> 4) | LBO1;
          rowl[j] = ((a_lo + off) >> 4) | LBO1;
        }
        const uint32_t tacc = tmem_base + (uint32_t)(it & 1) * acc_stride;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int ky = tap / 3, kx = tap % 3;
          if (tap == 3) {                          // rows 2, 3 (the pair loaded for this tile) are first touched here
            mbar_wait(smem_u32(&pair_full[p1]), (uint32_t)((p1 ? nfull1 : nfull0) & 1));
            if (p1) ++nfull1; else ++nfull0;
          }
          if (!((dbg & 2) && (w > t0 || tap >= RL_NSTAGE))) mbar_wait(smem_u32(&w_full[ws]), wpar);
          tc_fence_after();
          const uint32_t bh = ((b_ring + (uint32_t)(ws * B_TAP_BYTES)) >> 4) | LBO1, bl = bh + ((64 * 128) >> 4);
          // per output row: lo*hi, hi*lo, hi*hi back to back (consecutive MMAs that share an operand are cheaper: the
          // N = 64 MMAs are bound by shared-memory operand fetch, ~80 cycles each instead of the 32-cycle math floor;
          // alternating the two rows' accumulators was measured 15 % slower)
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const uint32_t ah = rowh[r + ky] + (uint32_t)(kx * 8), al = rowl[r + ky] + (uint32_t)(kx * 8);
#pragma unroll
            for (int k16 = 0; k16 < 4; ++k16) {
              const uint32_t acc = (gi > 0 || tap > 0 || k16 > 0) ? 1u : 0u;
              const uint32_t tm = tacc + (uint32_t)r * row_stride;
              if (cat) {
                umma_lo_elect(tm, ah + 2 * k16, bh + 2 * k16, idesc128, acc);   // [hi*hi | hi*lo]
                umma_lo_elect(tm, al + 2 * k16, bh + 2 * k16, idesc, 1u);       // lo*hi
              } else if (!single) {
                umma_lo_elect(tm, al + 2 * k16, bh + 2 * k16, idesc, acc);
                umma_lo_elect(tm, ah + 2 * k16, bl + 2 * k16, idesc, 1u);
                umma_lo_elect(tm, ah + 2 * k16, bh + 2 * k16, idesc, 1u);
              } else {
                umma_lo_elect(tm, ah + 2 * k16, bh + 2 * k16, idesc, acc);
              }
            }
          }
          umma_commit_elect(smem_u32(&w_empty[ws]));
          if (++ws == RL_NSTAGE) {
            ws = 0;
            wpar ^= 1;
          }
          if (tap == 5) umma_commit_elect(smem_u32(&pair_empty[p0]));   // rows 0, 1 are dead: the next tile's new rows go here
        }
        if (next_fresh) umma_commit_elect(smem_u32(&pair_empty[p1]));   // the next tile reloads both pairs
        }
        umma_commit_elect(smem_u32(&acc_full[it & 1]));
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5 -> TMEM lane groups 2,3,0,1)
    // TMEM -> registers -> warp-private swizzled staging -> 128-byte coalesced row segments (a direct store would
    // scatter 16-byte pieces at a 256-byte stride: 32 sectors per instruction, which throttles the LSU)
    const int lg = warp & 3;
    unsigned char* stg = stg_base + lg * (32 * 128);
    const int pq = lane >> 3, cc = lane & 7;          // copy-out: pixel 4 i + pq, float4 column cc of the 32-channel half
    float4 bb[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
    // BatchNorm statistics of the output (stats != nullptr, one output-channel group): per-channel sum and sum of
    // squares of what is stored, accumulated per thread over its pixels (fixed 2 x 4 channels per lane), reduced over
    // the CTA in shared memory and written as ONE row of 128 floats per CTA (stats[blockIdx.x][{sum, sumsq}][64]);
    // tatt_bn_finalize adds the rows in double.  (A first version used fp64 atomics on the 128 addresses straight from
    // every warp: 75 k atomics on 128 addresses cost as much as the separate statistics pass it replaced.)
    float4 ssum[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)}, ssq[2] = {ssum[0], ssum[0]};
    auto flush_stats = [&](int g) {
      if (!stats || g < 0) return;
      float* red = reinterpret_cast<float*>(stat_red);          // [4 warps][128]
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float v[8] = {ssum[h].x, ssum[h].y, ssum[h].z, ssum[h].w, ssq[h].x, ssq[h].y, ssq[h].z, ssq[h].w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          v[k] += __shfl_xor_sync(0xffffffffu, v[k], 8);
          v[k] += __shfl_xor_sync(0xffffffffu, v[k], 16);
        }
        if (pq == 0) {
          const int c0 = h * 32 + cc * 4;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            red[lg * 128 + c0 + k] = v[k];
            red[lg * 128 + 64 + c0 + k] = v[4 + k];
          }
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");             // the four epilogue warps
      const int i = lg * 32 + lane;                               // 0..127
      stats[(long long)blockIdx.x * 128 + i] = red[i] + red[128 + i] + red[256 + i] + red[384 + i];
    };
    int it = 0, gcur = -1;
    for (int w = t0; w < t1; ++w, ++it) {
      const int g = w / ntiles, t = w - g * ntiles;
      if (g != gcur) {
        flush_stats(gcur);
        gcur = g;
        if (bias) {
#pragma unroll
          for (int h = 0; h < 2; ++h) bb[h] = __ldg(reinterpret_cast<const float4*>(bias + g * 64 + h * 32 + cc * 4));
        }
      }
      const int yb = t % nyb, xb = (t / nyb) % nxt, n = t / (nyb * nxt);
      const int y0 = 2 * yb, px0 = xb * TW + lg * 32;
      mbar_wait(smem_u32(&acc_full[it & 1]), (uint32_t)((it >> 1) & 1));
      tc_fence_after();
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float* dst = Y + (((long long)n * H + (y0 + r)) * W + px0 + pq) * ldy + g * 64 + cc * 4;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
          for (int c16 = 0; c16 < 2; ++c16) {
            uint32_t v[16];
            const uint32_t tcol = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(it & 1) * acc_stride +
                                  (uint32_t)r * row_stride + (uint32_t)(h * 32 + c16 * 16);
            tmem_ld16(tcol, v);
            if (cat) {                              // + the hi*lo column block
              uint32_t v2[16];
              tmem_ld16(tcol + 64, v2);
#pragma unroll
              for (int q = 0; q < 16; ++q) v[q] = __float_as_uint(__uint_as_float(v[q]) + __uint_as_float(v2[q]));
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
              *reinterpret_cast<float4*>(stg + lane * 128 + (((4 * c16 + q) ^ (lane & 7)) << 4)) =
                  make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                              __uint_as_float(v[4 * q + 3]));
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int p = 4 * i + pq;
            float4 o = *reinterpret_cast<const float4*>(stg + p * 128 + ((cc ^ (p & 7)) << 4));
            o.x += bb[h].x; o.y += bb[h].y; o.z += bb[h].z; o.w += bb[h].w;
            if (stats) {
              ssum[h].x += o.x; ssum[h].y += o.y; ssum[h].z += o.z; ssum[h].w += o.w;
              ssq[h].x = fmaf(o.x, o.x, ssq[h].x); ssq[h].y = fmaf(o.y, o.y, ssq[h].y);
              ssq[h].z = fmaf(o.z, o.z, ssq[h].z); ssq[h].w = fmaf(o.w, o.w, ssq[h].w);
            }
            if (!(dbg & 1)) *reinterpret_cast<float4*>(dst + (long long)(4 * i) * ldy + h * 32) = o;
          }
          __syncwarp();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&acc_empty[it & 1]));
    }
    flush_stats(gcur);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ weight gradient
// dWt[(tap,ci)][co] += sum_px X[px + tap][ci] * dY[px][co]  for the same 3x3 / 64->64 conv.  The reduction axis is the
// pixel axis, so both operands are MN-major tiles whose rows are pixels (128 B = 64 bf16 channels).  One k-tile =
// 64 pixels of one image row: ONE TMA halo load (3 rows x 66 px, hi+lo) + one dY tile serve all nine taps -- five
// 128-row M-tiles (taps 2i, 2i+1 as the two 64-channel MN blocks of a descriptor, LBO = byte distance between the two
// shifted halo windows), accumulated in five 64-column TMEM accumulators over the CTA's share of the pixels, then
// added to global memory with atomics (split-K over CTAs).
constexpr int WG_HALO_ROWS = 3 * 66;
constexpr int WG_X_PLANE = ((WG_HALO_ROWS * 128 + 1023) / 1024) * 1024;   // 25600
constexpr int WG_STAGE = 2 * WG_X_PLANE + 2 * 64 * 128;                   // X hi, X lo, dY hi, dY lo
constexpr int WG_NSTAGE = 3;

__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(192, 1)
conv3x3_wgrad_tma_kernel(const __grid_constant__ CUtensorMap tmXh, const __grid_constant__ CUtensorMap tmXl,
                         const __grid_constant__ CUtensorMap tmGh, const __grid_constant__ CUtensorMap tmGl,
                         float* __restrict__ dWt, float* __restrict__ partial, int H, int W, int total_tiles,
                         int tiles_per_cta, int single, int ngroups) {
  // C_out = 64 * ngroups: CTA b accumulates the 64-channel dY group g = b % ngroups over the pixel tiles of slot b / ngroups
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) unsigned long long bar_done, bar_full[WG_NSTAGE], bar_empty[WG_NSTAGE];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = blockIdx.x % ngroups, ldw = 64 * ngroups;
  const int t0 = (blockIdx.x / ngroups) * tiles_per_cta;
  int t1 = t0 + tiles_per_cta;
  if (t1 > total_tiles) t1 = total_tiles;
  const int nt = t1 - t0;
  const int nxs = W / 64;
  const uint32_t sbase = smem_u32(smem);

  if (tid == 0) {
    mbar_init(smem_u32(&bar_done), 1);
    for (int s = 0; s < WG_NSTAGE; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < nt; ++i) {
        const int s = i % WG_NSTAGE;
        if (i >= WG_NSTAGE) mbar_wait(smem_u32(&bar_empty[s]), (uint32_t)(((i / WG_NSTAGE) - 1) & 1));
        const int t = t0 + i;
        const int xs = t % nxs, y = (t / nxs) % H, n = t / (nxs * H);
        const uint32_t st = sbase + (uint32_t)(s * WG_STAGE);
        const uint32_t bar = smem_u32(&bar_full[s]);
        mbar_expect_tx(bar, (single ? 1u : 2u) * ((uint32_t)(WG_HALO_ROWS * 128) + 64u * 128u));
        tma_load_4d(st, &tmXh, bar, 0, xs * 64 - 1, y - 1, n);
        if (!single) tma_load_4d(st + WG_X_PLANE, &tmXl, bar, 0, xs * 64 - 1, y - 1, n);
        tma_load_4d(st + 2 * WG_X_PLANE, &tmGh, bar, grp * 64, xs * 64, y, n);
        if (!single) tma_load_4d(st + 2 * WG_X_PLANE + 64 * 128, &tmGl, bar, grp * 64, xs * 64, y, n);
      }
    }
  } else if (warp == 1) {
    {                                             // whole warp, uniform control flow; one elected lane issues
      // M = 128, N = 64, A and B MN-major (bits 15, 16)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                             ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      for (int i = 0; i < nt; ++i) {
        const int s = i % WG_NSTAGE;
        mbar_wait(smem_u32(&bar_full[s]), (uint32_t)((i / WG_NSTAGE) & 1));
        tc_fence_after();
        const uint32_t st = sbase + (uint32_t)(s * WG_STAGE);
        const uint32_t x_hi = st, x_lo = st + WG_X_PLANE, g_hi = st + 2 * WG_X_PLANE, g_lo = g_hi + 64 * 128;
#pragma unroll
        for (int mt = 0; mt < 5; ++mt) {
          const int ta = 2 * mt, tb = (2 * mt + 1 < 9) ? 2 * mt + 1 : 2 * mt;
          const uint32_t offa = (uint32_t)(((ta / 3) * 66 + (ta % 3)) * 128);
          const uint32_t offb = (uint32_t)(((tb / 3) * 66 + (tb % 3)) * 128);
          const uint32_t lbo = (tb != ta) ? (offb - offa) : 128u;
#pragma unroll
          for (int k16 = 0; k16 < 4; ++k16) {
            const uint32_t ko = k16 * 2048;   // 16 pixel rows
            const uint64_t dah = make_desc_mn(x_hi + offa + ko, lbo), dal = make_desc_mn(x_lo + offa + ko, lbo);
            const uint64_t dbh = make_desc_mn(g_hi + ko, 8192), dbl = make_desc_mn(g_lo + ko, 8192);
            const uint32_t tm = tmem_base + (uint32_t)(mt * 64);
            if (!single) {
              umma_bf16_elect(tm, dal, dbh, idesc, (i > 0 || k16 > 0) ? 1u : 0u);
              umma_bf16_elect(tm, dah, dbl, idesc, 1u);
              umma_bf16_elect(tm, dah, dbh, idesc, 1u);
            } else {
              umma_bf16_elect(tm, dah, dbh, idesc, (i > 0 || k16 > 0) ? 1u : 0u);
            }
          }
        }
        umma_commit_elect(smem_u32(&bar_empty[s]));
      }
      umma_commit_elect(smem_u32(&bar_done));
    }
  } else if (nt > 0) {
    mbar_wait(smem_u32(&bar_done), 0);
    tc_fence_after();
    const int lg = warp & 3;
    const int m = lg * 32 + lane;              // row of the 128-row M tile: (m >= 64) selects the second tap
#pragma unroll
    for (int mt = 0; mt < 5; ++mt) {
      const int tap = 2 * mt + (m >> 6);
#pragma unroll
      for (int c16 = 0; c16 < 4; ++c16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(mt * 64 + c16 * 16), v);
        if (tap < 9) {
          const long long off = ((long long)(tap * 64 + (m & 63))) * 64 + c16 * 16;       // inside a [576][64] tile
          if (partial) {       // per-CTA partial tile, summed by wgrad_reduce_kernel (no same-address atomics)
            float4* dst = reinterpret_cast<float4*>(partial + (long long)blockIdx.x * (576 * 64) + off);
#pragma unroll
            for (int q = 0; q < 4; ++q)
              dst[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                   __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
          } else {
            float* dst = dWt + ((long long)(tap * 64 + (m & 63))) * ldw + grp * 64 + c16 * 16;
#pragma unroll
            for (int j = 0; j < 16; ++j) atomicAdd(dst + j, __uint_as_float(v[j]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// out[k][64 g + c] = sum over the CTAs b = j * ngroups + g of partial[b][k][c]   (ngroups = 1: a plain sum of tiles)
// 64 consecutive elements per CTA x 8 slot groups (one thread walking all ~148 slots took 33 us per call); fixed
// summation order
__global__ void __launch_bounds__(512) wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ out,
                                                           int nslots, int ngroups) {
  __shared__ float sm[8][64];
  const int n = 576 * 64;
  const int e = threadIdx.x & 63, q = threadIdx.x >> 6;
  const int i = blockIdx.x * 64 + e;            // < n * ngroups: the grid is exactly n * ngroups / 64 CTAs
  const int g = i / n, r = i - g * n;
  float a = 0.f;
#pragma unroll 4
  for (int j = q; j < nslots; j += 8) a += partial[(long long)(j * ngroups + g) * n + r];
  sm[q][e] = a;
  __syncthreads();
  if (q != 0) return;
  a = ((sm[0][e] + sm[1][e]) + (sm[2][e] + sm[3][e])) + ((sm[4][e] + sm[5][e]) + (sm[6][e] + sm[7][e]));
  out[(long long)(r >> 6) * (64 * ngroups) + g * 64 + (r & 63)] = a;
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

static inline long long rup8(long long x) { return (x + 7) & ~7LL; }

}  // namespace

// conv3x3 64 -> Cout (64, 128, 192 or 256) through the TMA kernel.  Returns 0 ok, 1 error, -1 not eligible (caller
// falls through).  Cout > 64 needs the persistent rolling-halo kernel (mode 3).
int tatt_tc3_conv3x3_launch(const float* X, const float* Wt, const float* bias, float* Y, int nimg, int H, int W,
                            int Cin, int Cout, int single, int a_valid, void* ws, long long ws_bytes, float* stats,
                            cudaStream_t st) {
  static const int mode = []() {          // TATT_TMA: 0 = off, 1 / 2 = one tile per CTA (base_offset 0 / from address),
    const char* e = getenv("TATT_TMA");   //           3 = persistent rolling-halo kernel (default)
    return e ? atoi(e) : 3;
  }();
  if (mode == 0 || ws == nullptr || W % TW != 0 || H % 2 != 0) return -1;
  if (Cout % 64 != 0 || Cout < 64 || Cout > 256 || ((Cout != 64 || stats) && mode != 3) || (stats && Cout != 64)) return -1;
  static const int multi_in = []() {      // TATT_ROLL_CIN=0: C_in > 64 goes back to the im2col GEMM engine (A/B timing)
    const char* e = getenv("TATT_ROLL_CIN");
    return e ? atoi(e) : 1;
  }();
  if (Cin % 64 != 0 || Cin < 64 || Cin > 256 || (Cin != 64 && (mode != 3 || stats || !multi_in || Cout != 64))) return -1;
  const int ngroups = Cout / 64, K = 9 * Cin;
  EncodeTiledFn enc = get_encode();
  if (!enc) return -1;
  const long long P = (long long)nimg * H * W;
  const long long nA = rup8(P * Cin), nB = rup8((long long)Cout * K);
  if ((long long)sizeof(__nv_bfloat16) * 2 * (nA + nB) > ws_bytes || (((uintptr_t)ws) & 15)) return -1;
  __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(ws);
  __nv_bfloat16 *Ahi = base, *Alo = base + nA, *Bhi = base + 2 * nA, *Blo = base + 2 * nA + nB;
  int rc = a_valid ? 0 : tatt_tc2_split(X, Cin, P, Cin, 0, Ahi, Alo, nullptr, st);   // activations -> [P][Cin] planes
  if (rc) return rc;
  rc = tatt_tc2_split(Wt, Cout, K, Cout, 1, Bhi, Blo, nullptr, st);          // Wt[9 Cin][Cout] -> planes [Cout co][9 Cin k]
  if (rc) return rc;

  constexpr int R = 2;
  CUtensorMap tmAh, tmAl, tmBh, tmBl;
  {
    const cuuint64_t pxb = (cuuint64_t)Cin * 2;                 // bytes per pixel of a plane
    cuuint64_t gdim[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)nimg};
    cuuint64_t gstr[3] = {pxb, (cuuint64_t)W * pxb, (cuuint64_t)H * W * pxb};
    cuuint32_t box[4] = {64, HALO_W, (cuuint32_t)(mode == 3 ? 2 : R + 2), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    for (int pl = 0; pl < 2; ++pl) {
      CUresult r = enc(pl ? &tmAl : &tmAh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, pl ? (void*)Alo : (void*)Ahi, gdim, gstr,
                       box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return tatt_set_error("cuTensorMapEncodeTiled(A) failed: %d", (int)r);
    }
  }
  {
    cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)Cout};
    cuuint64_t gstr[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {64, 64};
    cuuint32_t estr[2] = {1, 1};
    for (int pl = 0; pl < 2; ++pl) {
      CUresult r = enc(pl ? &tmBl : &tmBh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, pl ? (void*)Blo : (void*)Bhi, gdim, gstr,
                       box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return tatt_set_error("cuTensorMapEncodeTiled(B) failed: %d", (int)r);
    }
  }
  if (mode == 3) {
    const int smem = 2 * RL_A_PLANE + RL_NSTAGE * B_TAP_BYTES + RL_STG_BYTES + 1024;
    const int ntiles = nimg * (H / 2) * (W / TW);
    int dev = 0, nsm = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    const int grid = ntiles * ngroups < nsm ? ntiles * ngroups : nsm;
    // timing-experiment switches of the kernel (1: skip the global stores, 2: load the weight taps once) produce WRONG
    // results by design; they are reachable only in builds with -DTATT_ROLL_EXPERIMENTS
#ifdef TATT_ROLL_EXPERIMENTS
    static const int dbg = []() {
      const char* e = getenv("TATT_ROLL_DBG");
      return e ? atoi(e) : 0;
    }();
#else
    const int dbg = 0;
#endif
    static const int cat_on = []() {               // TATT_ROLL_CAT=0: three N = 64 MMAs per k-step instead of N = 128 + N = 64
      const char* e = getenv("TATT_ROLL_CAT");
      return e ? atoi(e) : 1;
    }();
    auto* kern = (Cin > 64)
                     ? (single ? conv3x3_roll_kernel<true, false, true>
                               : (cat_on ? conv3x3_roll_kernel<false, true, true> : conv3x3_roll_kernel<false, false, true>))
                     : (single ? conv3x3_roll_kernel<true, false, false>
                               : (cat_on ? conv3x3_roll_kernel<false, true, false> : conv3x3_roll_kernel<false, false, false>));
    TATT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<grid, 192, smem, st>>>(tmAh, tmAl, tmBh, tmBl, Y, bias, H, W, ntiles, dbg, Cout, ngroups, stats, Cin / 64);
    TATT_LAUNCH_CHECK("conv3x3_roll_kernel");
    return 0;
  }
  constexpr int A_PLANE = ((((R + 2) * HALO_W) * 128 + 1023) / 1024) * 1024;
  const int smem = 2 * A_PLANE + NSTAGE * B_TAP_BYTES + 1024;
  TATT_CUDA(cudaFuncSetAttribute(conv3x3_tma_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int grid = nimg * (H / R) * (W / TW);
  conv3x3_tma_kernel<R><<<grid, 192, smem, st>>>(tmAh, tmAl, tmBh, tmBl, Y, bias, H, W, mode == 2 ? 1 : 0, single);
  TATT_LAUNCH_CHECK("conv3x3_tma_kernel");
  return 0;
}

// conv3x3 64->64 weight gradient through the TMA kernel; dWt must be zeroed by the caller.  Same return convention.
int tatt_tc3_conv3x3_wgrad_launch(const float* X, const float* dY, float* dWt, int nimg, int H, int W, int Cout,
                                  int single, int a_valid, int b_valid, void* ws, long long ws_bytes, cudaStream_t st) {
  static const int mode = []() {
    const char* e = getenv("TATT_TMA_WGRAD");
    return e ? atoi(e) : 1;
  }();
  if (mode == 0 || ws == nullptr || W % 64 != 0) return -1;
  if (Cout % 64 != 0 || Cout < 64 || Cout > 256) return -1;
  const int ngroups = Cout / 64;
  EncodeTiledFn enc = get_encode();
  if (!enc) return -1;
  const long long P = (long long)nimg * H * W;
  const long long nA = rup8(P * 64), nG = rup8(P * Cout);
  const long long plane_bytes = (long long)sizeof(__nv_bfloat16) * 2 * (nA + nG);
  if (plane_bytes > ws_bytes || (((uintptr_t)ws) & 15)) return -1;
  __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(ws);
  __nv_bfloat16 *Xh = base, *Xl = base + nA, *Gh = base + 2 * nA, *Gl = base + 2 * nA + nG;
  int rc = a_valid ? 0 : tatt_tc2_split(X, 64, P, 64, 0, Xh, Xl, nullptr, st);
  if (rc) return rc;
  rc = b_valid ? 0 : tatt_tc2_split(dY, Cout, P, Cout, 0, Gh, Gl, nullptr, st);     // b_valid: the caller split dY already
  if (rc) return rc;
  CUtensorMap tm[4];
  cuuint32_t estr[4] = {1, 1, 1, 1};
  cuuint32_t boxX[4] = {64, 66, 3, 1}, boxG[4] = {64, 64, 1, 1};
  void* ptrs[4] = {Xh, Xl, Gh, Gl};
  for (int i = 0; i < 4; ++i) {
    const cuuint64_t ch = i < 2 ? 64 : (cuuint64_t)Cout;
    cuuint64_t gdim[4] = {ch, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)nimg};
    cuuint64_t gstr[3] = {ch * 2, (cuuint64_t)W * ch * 2, (cuuint64_t)H * W * ch * 2};
    CUresult r = enc(&tm[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, ptrs[i], gdim, gstr, i < 2 ? boxX : boxG, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return tatt_set_error("cuTensorMapEncodeTiled(wgrad) failed: %d", (int)r);
  }
  const int total = (int)(P / 64);
  int dev = 0, nsm = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  int slots = nsm / ngroups;                       // CTAs per output-channel group; ngroups * slots <= #SMs
  if (slots > total) slots = total;
  const int per = (total + slots - 1) / slots;
  slots = (total + per - 1) / per;
  const int grid = slots * ngroups;
  const int smem = WG_NSTAGE * WG_STAGE + 1024;
  // per-CTA partial tiles live behind the planes when the workspace is large enough -- and behind the slot right after
  // the dY planes where the data-gradient pass of the same layer (which reuses the dY planes, F_A_VALID with the
  // workspace advanced past the X planes) puts its weight planes: that pass may run concurrently on another stream
  float* partial = nullptr;
  const long long part_off = ((plane_bytes + 4 * rup8(576LL * Cout) + 64 + 15) & ~15LL);
  const long long part_bytes = (long long)grid * 576 * 64 * sizeof(float);
  if (ws_bytes >= part_off + part_bytes + 16)
    partial = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(ws) + part_off);
  TATT_CUDA(cudaFuncSetAttribute(conv3x3_wgrad_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  conv3x3_wgrad_tma_kernel<<<grid, 192, smem, st>>>(tm[0], tm[1], tm[2], tm[3], dWt, partial, H, W, total, per, single,
                                                    ngroups);
  TATT_LAUNCH_CHECK("conv3x3_wgrad_tma_kernel");
  if (partial) {
    wgrad_reduce_kernel<<<576 * ngroups, 512, 0, st>>>(partial, dWt, slots, ngroups);
    TATT_LAUNCH_CHECK("wgrad_reduce_kernel");
  }
  return 0;
}
