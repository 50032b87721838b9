// Persistent recurrence kernel of the recurrent positional encoding (quirk Q1): InfoTransformer.forward,
// /root/reference/model/transformer_v2.py:177,201,215-221 -- a bidirectional nn.GRU(H*C -> H*C/2, batch_first) fed
// [W, N, H*C], i.e. T = N sequential steps over a "batch" of W rows.  The GRU input is identical at every step, so
// GI = W_ih x + b_ih is computed once outside; what remains is the chain
//     GH_s = H_{s-1} W_hh^T + b_hh ;  r,z = sigmoid(GI + GH) ;  n = tanh(GI_n + r * GH_n) ;  H_s = n + z (H_{s-1} - n)
// Round 1 ran it as 2 launches per step (a 128-CTA GEMM + a gate kernel): 128 launches per pass, 19.5 + 5.5 us per step,
// with both W_hh planes (25 MB) re-read from L2 at every step.
//
// Here ONE launch runs all N steps.  Grid = 2 directions x Hd/16 slices; CTA (d, c) owns hidden units [16c, 16c+16) of
// direction d, i.e. the 48 gate rows {r, z, n} x 16 units of W_hh, so the gate math is CTA-local:
//   * the bf16 HI plane of its W_hh slice ([48][Hd], 96 KB at Hd = 1024) stays resident in shared memory for all N steps;
//   * per step it streams the hidden state of its direction (bf16 hi/lo planes, [Wd][Hd], written by the 64 CTAs of the
//     direction in the previous step) and the LO plane of its W_hh slice through a 3-stage TMA ring (38 KB per 64-wide
//     k-block: A_hi, A_lo, W_lo), and issues hi*hi + hi*lo + lo*hi tcgen05 MMAs (M = 128 rows, N = 48) into ONE TMEM
//     accumulator -- the same fp32-parity scheme as the rest of the engine;
//   * 8 epilogue warps (thread = one batch row x 8 units, TMEM 32x32b layout) do the gate math with the hidden state of
//     their (row, units) register-resident across steps, and write H_s (fp32 + bf16 planes), the saved gates and the
//     output slice query_pos[b] directly;
//   * step barrier per direction: red.release.gpu on a per-step counter + ld.acquire spin by the TMA producer thread,
//     then fence.proxy.async before the TMA loads of the freshly written planes.  All CTAs are co-resident (cooperative
//     launch, grid <= #SMs, one CTA per SM); spins are bounded and raise a flag instead of hanging.
#include <cooperative_groups.h>
#include <stdlib.h>
#include "tc_prims.cuh"

using namespace tcp;

namespace {

constexpr int RU = 16;                    // hidden units per CTA
constexpr int RN = 3 * RU;                // gate rows per CTA = MMA N
constexpr int R_WT = RN * 128;            // one [48][64] bf16 K-major SWIZZLE_128B tile: 6144 B
constexpr int R_AT = 128 * 128;           // one [128][64] bf16 tile: 16384 B
constexpr int R_STAGE = 2 * R_AT + R_WT;  // A_hi | A_lo | W_lo
constexpr int R_NSTAGE = 3;
constexpr int R_THREADS = 64 + 256;       // warp 0 TMA producer, warp 1 MMA issuer, warps 2..9 epilogue
constexpr unsigned int R_SPIN_LIMIT = 1u << 24;

__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) {
  // tanh(x) = 1 - 2 / (e^{2x} + 1); exact limits at +-inf, abs error ~1e-7 (the path's tolerance is 1e-3)
  const float e = __expf(2.f * x);
  return 1.f - __fdividef(2.f, e + 1.f);
}

struct RpeFwdParams {
  const float* GI;      // [2][Wd][3Hd]   W_ih x + b_ih (constant over steps)
  const float* BHH;     // [2][3Hd]
  float* HALL;          // [2][N+1][Wd][Hd]  fp32 hidden states, HALL[:,0] == 0
  float* GATES;         // [N][2][Wd][4][Hd] r, z, n, W_hn h + b_hn (saved for backward) or nullptr
  float* QPOS;          // [N][Himg*Wd][C]
  __nv_bfloat16* HP;    // bf16 hi plane of HALL (same indexing); lo plane at HP + h_lo
  long long h_lo;
  unsigned int* sync;   // [2][N] step counters + [1] abort flag, zeroed by the launcher
  int N, Wd, Hd, C, Himg;
};

__global__ void __launch_bounds__(R_THREADS, 1)
rpe_fwd_persist_kernel(const __grid_constant__ CUtensorMap tmHh, const __grid_constant__ CUtensorMap tmHl,
                       const __grid_constant__ CUtensorMap tmWh, const __grid_constant__ CUtensorMap tmWl,
                       const RpeFwdParams p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) unsigned long long bar_full[R_NSTAGE], bar_empty[R_NSTAGE], bar_wres, bar_acc;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KB = p.Hd >> 6;                       // 64-wide k-blocks
  const int nslice = p.Hd / RU;                   // CTAs per direction
  const int dir = blockIdx.x / nslice, cs = blockIdx.x % nslice;
  const int u_base = cs * RU;
  const uint32_t w_res = smem_u32(smem);                              // KB resident W_hi tiles
  const uint32_t ring = w_res + (uint32_t)KB * R_WT;
  const int box_rows = p.Wd < 128 ? p.Wd : 128;
  unsigned int* cnt = p.sync + (size_t)dir * p.N;
  unsigned int* abort_flag = p.sync + 2 * (size_t)p.N;

  if (tid == 0) {
    for (int s = 0; s < R_NSTAGE; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    mbar_init(smem_u32(&bar_wres), 1);
    mbar_init(smem_u32(&bar_acc), 1);
    mbar_fence_init();
  }
  // rows >= Wd of the A tiles are never written by TMA: zero the ring once so the unused accumulator rows stay finite
  if (p.Wd < 128) {
    for (int i = tid; i < R_NSTAGE * R_STAGE / 16; i += R_THREADS)
      reinterpret_cast<uint4*>(smem + (size_t)KB * R_WT)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
  }
  __syncwarp();
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_s), 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const int wrow = dir * 3 * p.Hd + u_base;      // row of gate 0 of this slice in the [2*3Hd][Hd] weight planes
      mbar_expect_tx(smem_u32(&bar_wres), (uint32_t)(KB * R_WT));
      for (int kb = 0; kb < KB; ++kb)
#pragma unroll
        for (int g = 0; g < 3; ++g)
          tma_load_2d(w_res + (uint32_t)(kb * R_WT + g * RU * 128), &tmWh, smem_u32(&bar_wres), kb * 64, wrow + g * p.Hd);
      int st = 0;
      uint32_t epar = 0;
      bool wrapped = false;
      const unsigned int target = (unsigned int)nslice;
      for (int s = 0; s < p.N; ++s) {
        if (s > 0) {                                 // every CTA of this direction has written H_s (planes included)
          unsigned int spins = 0;
          while (ld_acquire(cnt + (s - 1)) < target) {
            if (++spins > R_SPIN_LIMIT || *reinterpret_cast<volatile unsigned int*>(abort_flag) != 0u) {
              atomicExch(abort_flag, 1u);            // never hang: results are invalid, the flag says so
              break;
            }
          }
          fence_proxy_async();                       // generic-proxy writes observed above -> async-proxy (TMA) reads below
        }
        const int hrow = (dir * (p.N + 1) + s) * p.Wd;
        for (int kb = 0; kb < KB; ++kb) {
          if (wrapped) mbar_wait(smem_u32(&bar_empty[st]), epar);
          const uint32_t dst = ring + (uint32_t)(st * R_STAGE), bar = smem_u32(&bar_full[st]);
          mbar_expect_tx(bar, (uint32_t)(2 * box_rows * 128 + R_WT));
          tma_load_2d(dst, &tmHh, bar, kb * 64, hrow);
          tma_load_2d(dst + R_AT, &tmHl, bar, kb * 64, hrow);
#pragma unroll
          for (int g = 0; g < 3; ++g)
            tma_load_2d(dst + 2 * R_AT + (uint32_t)(g * RU * 128), &tmWl, bar, kb * 64, wrow + g * p.Hd);
          if (++st == R_NSTAGE) {
            st = 0;
            if (wrapped) epar ^= 1;
            wrapped = true;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (whole warp, elected lane issues)
    constexpr uint32_t idesc = idesc_bf16(128, RN);
    mbar_wait(smem_u32(&bar_wres), 0);
    int st = 0;
    uint32_t fpar = 0;
    for (int s = 0; s < p.N; ++s) {
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(smem_u32(&bar_full[st]), fpar);
        tc_fence_after();
        const uint32_t ah = desc_lo(ring + (uint32_t)(st * R_STAGE)), al = ah + (R_AT >> 4), wl = al + (R_AT >> 4);
        const uint32_t wh = desc_lo(w_res + (uint32_t)(kb * R_WT));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          umma_elect(tmem_base, ah + 2 * k, wh + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);   // hi * hi
          umma_elect(tmem_base, ah + 2 * k, wl + 2 * k, idesc, 1u);                            // hi * lo
          umma_elect(tmem_base, al + 2 * k, wh + 2 * k, idesc, 1u);                            // lo * hi
        }
        umma_commit_elect(smem_u32(&bar_empty[st]));
        if (++st == R_NSTAGE) {
          st = 0;
          fpar ^= 1;
        }
      }
      umma_commit_elect(smem_u32(&bar_acc));
      // the next step's first MMA overwrites the accumulator: it cannot start before this CTA's epilogue has read it,
      // because the next step's A tiles only arrive after the step barrier, which needs this CTA's own arrival
    }
  } else {
    // ------------------------------------------------------------------ epilogue: gates (8 warps, 2 per TMEM lane group)
    const int ew = warp - 2;                       // 0..7
    const int lg = warp & 3;                       // TMEM lane group this warp may access
    const int half = (ew >> 2) & 1;                // units [8 half, 8 half + 8) of the slice  (warps 2..5 -> 0, 6..9 -> 1)
    const int w = lg * 32 + lane;                  // batch row
    const bool valid = w < p.Wd;
    const int j0 = u_base + half * 8;              // first hidden unit of this thread
    float gr[8], gz[8], gn[8], bn[8], hp[8];
    {
      const long long g3 = ((long long)dir * p.Wd + (valid ? w : 0)) * 3 * p.Hd + j0;
      const float* bh = p.BHH + (long long)dir * 3 * p.Hd + j0;
#pragma unroll
      for (int q = 0; q < 8; q += 4) {
        const float4 a = *reinterpret_cast<const float4*>(p.GI + g3 + q);
        const float4 b = *reinterpret_cast<const float4*>(p.GI + g3 + p.Hd + q);
        const float4 c = *reinterpret_cast<const float4*>(p.GI + g3 + 2 * p.Hd + q);
        const float4 ba = *reinterpret_cast<const float4*>(bh + q);
        const float4 bb = *reinterpret_cast<const float4*>(bh + p.Hd + q);
        const float4 bc = *reinterpret_cast<const float4*>(bh + 2 * p.Hd + q);
        gr[q] = a.x + ba.x; gr[q + 1] = a.y + ba.y; gr[q + 2] = a.z + ba.z; gr[q + 3] = a.w + ba.w;
        gz[q] = b.x + bb.x; gz[q + 1] = b.y + bb.y; gz[q + 2] = b.z + bb.z; gz[q + 3] = b.w + bb.w;
        gn[q] = c.x; gn[q + 1] = c.y; gn[q + 2] = c.z; gn[q + 3] = c.w;
        bn[q] = bc.x; bn[q + 1] = bc.y; bn[q + 2] = bc.z; bn[q + 3] = bc.w;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) hp[q] = 0.f;
    }
    const uint32_t trow = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(half * 8);
    const int f0 = dir * p.Hd + j0;                // feature index in the [H*C] output vector
    const int hh = f0 / p.C, cc = f0 % p.C;
    for (int s = 0; s < p.N; ++s) {
      mbar_wait(smem_u32(&bar_acc), (uint32_t)(s & 1));
      tc_fence_after();
      uint32_t ar[8], az[8], an[8];
      tmem_ld8_nowait(trow, ar);
      tmem_ld8_nowait(trow + RU, az);
      tmem_ld8_nowait(trow + 2 * RU, an);
      tmem_ld_wait();
      tc_fence_before();
      float rr[8], zz[8], nn[8], gh[8], hn[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        rr[q] = fast_sigmoid(gr[q] + __uint_as_float(ar[q]));
        zz[q] = fast_sigmoid(gz[q] + __uint_as_float(az[q]));
        gh[q] = __uint_as_float(an[q]) + bn[q];
        nn[q] = fast_tanh(gn[q] + rr[q] * gh[q]);
        hn[q] = nn[q] + zz[q] * (hp[q] - nn[q]);
        hp[q] = hn[q];
      }
      if (valid) {
        const long long hidx = (((long long)dir * (p.N + 1) + s + 1) * p.Wd + w) * p.Hd + j0;
        uint2 h0, l0, h1, l1;
        split4(hn[0], hn[1], hn[2], hn[3], h0, l0);
        split4(hn[4], hn[5], hn[6], hn[7], h1, l1);
        // planes first: they are what the other CTAs wait for
        *reinterpret_cast<uint4*>(p.HP + hidx) = make_uint4(h0.x, h0.y, h1.x, h1.y);
        *reinterpret_cast<uint4*>(p.HP + p.h_lo + hidx) = make_uint4(l0.x, l0.y, l1.x, l1.y);
      }
      // publish: every epilogue thread fences its own stores, the 256 threads meet, one arrives on the step counter
      fence_proxy_async();
      __threadfence();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (tid == 64) red_release_add(cnt + s, 1u);
      if (valid) {                                  // off the critical path: fp32 state, saved gates, output slice
        const long long hidx = (((long long)dir * (p.N + 1) + s + 1) * p.Wd + w) * p.Hd + j0;
        *reinterpret_cast<float4*>(p.HALL + hidx) = make_float4(hn[0], hn[1], hn[2], hn[3]);
        *reinterpret_cast<float4*>(p.HALL + hidx + 4) = make_float4(hn[4], hn[5], hn[6], hn[7]);
        if (p.GATES) {
          const long long gi = ((((long long)s * 2 + dir) * p.Wd + w) * 4) * p.Hd + j0;
          *reinterpret_cast<float4*>(p.GATES + gi) = make_float4(rr[0], rr[1], rr[2], rr[3]);
          *reinterpret_cast<float4*>(p.GATES + gi + 4) = make_float4(rr[4], rr[5], rr[6], rr[7]);
          *reinterpret_cast<float4*>(p.GATES + gi + p.Hd) = make_float4(zz[0], zz[1], zz[2], zz[3]);
          *reinterpret_cast<float4*>(p.GATES + gi + p.Hd + 4) = make_float4(zz[4], zz[5], zz[6], zz[7]);
          *reinterpret_cast<float4*>(p.GATES + gi + 2 * p.Hd) = make_float4(nn[0], nn[1], nn[2], nn[3]);
          *reinterpret_cast<float4*>(p.GATES + gi + 2 * p.Hd + 4) = make_float4(nn[4], nn[5], nn[6], nn[7]);
          *reinterpret_cast<float4*>(p.GATES + gi + 3 * p.Hd) = make_float4(gh[0], gh[1], gh[2], gh[3]);
          *reinterpret_cast<float4*>(p.GATES + gi + 3 * p.Hd + 4) = make_float4(gh[4], gh[5], gh[6], gh[7]);
        }
        const int b = dir == 0 ? s : p.N - 1 - s;
        float* q = p.QPOS + ((long long)b * p.Himg * p.Wd + (long long)hh * p.Wd + w) * p.C + cc;
        *reinterpret_cast<float4*>(q) = make_float4(hn[0], hn[1], hn[2], hn[3]);
        *reinterpret_cast<float4*>(q + 4) = make_float4(hn[4], hn[5], hn[6], hn[7]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 64);
}

}  // namespace

extern "C" {

// 0 when the persistent kernel supports the shape on this device, else non-zero (callers then use the per-step path)
int tatt_rpe_persist_supported(int N, int Wd, int Hd, int C) {
  if (N < 1 || Wd < 8 || Wd > 128 || Wd % 8 || Hd % 64 || Hd < 64 || Hd > 1024 || C % 8 || (Hd % C) || C < 8) return 1;
  int dev = 0, nsm = 0, coop = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 1;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  if (!coop || 2 * (Hd / RU) > nsm) return 1;
  static const int off = []() {
    const char* e = getenv("TATT_RPE_PERSIST");
    return (e && e[0] == '0') ? 1 : 0;
  }();
  return off;
}

int tatt_rpe_sync_bytes(int N) { return (int)sizeof(unsigned int) * (2 * N + 4); }

int tatt_rpe_fwd(const float* GI, const float* BHH, const void* WPL, long long w_lo, void* HPL, long long h_lo,
                 float* HALL, float* GATES, float* QPOS, void* sync, int N, int Wd, int Hd, int C, int Himg,
                 void* stream) {
  TATT_REQUIRE(tatt_rpe_persist_supported(N, Wd, Hd, C) == 0, "rpe_fwd: unsupported shape N=%d Wd=%d Hd=%d C=%d", N, Wd,
               Hd, C);
  TATT_REQUIRE(2 * Hd == Himg * C, "rpe_fwd: 2*Hd (%d) must equal Himg*C (%d)", 2 * Hd, Himg * C);
  TATT_REQUIRE(((uintptr_t)WPL & 15) == 0 && ((uintptr_t)HPL & 15) == 0 && (w_lo % 8) == 0 && (h_lo % 8) == 0,
               "rpe_fwd: planes must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const __nv_bfloat16* wp = reinterpret_cast<const __nv_bfloat16*>(WPL);
  __nv_bfloat16* hp = reinterpret_cast<__nv_bfloat16*>(HPL);
  CUtensorMap tmHh, tmHl, tmWh, tmWl;
  const int box_rows = Wd < 128 ? Wd : 128;
  const long long hrows = 2LL * (N + 1) * Wd;
  if (make_map_2d(&tmHh, hp, hrows, Hd, Hd, box_rows)) return 1;
  if (make_map_2d(&tmHl, hp + h_lo, hrows, Hd, Hd, box_rows)) return 1;
  if (make_map_2d(&tmWh, wp, 2LL * 3 * Hd, Hd, Hd, RU)) return 1;
  if (make_map_2d(&tmWl, wp + w_lo, 2LL * 3 * Hd, Hd, Hd, RU)) return 1;
  TATT_CUDA(cudaMemsetAsync(sync, 0, (size_t)tatt_rpe_sync_bytes(N), st));
  RpeFwdParams p{GI, BHH, HALL, GATES, QPOS, hp, h_lo, reinterpret_cast<unsigned int*>(sync), N, Wd, Hd, C, Himg};
  const int KB = Hd / 64;
  const int smem = KB * R_WT + R_NSTAGE * R_STAGE + 1024;
  TATT_CUDA(cudaFuncSetAttribute(rpe_fwd_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * (Hd / RU));
  cfg.blockDim = dim3(R_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;      // all CTAs co-resident or the launch fails (no silent deadlock)
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  TATT_CUDA(cudaLaunchKernelEx(&cfg, rpe_fwd_persist_kernel, tmHh, tmHl, tmWh, tmWl, p));
  return 0;
}

}  // extern "C"
