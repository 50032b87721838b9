// One TransformerDecoderLayer_TP forward (+ the decoder's shared final LayerNorm) as ONE tcgen05 kernel:
// /root/reference/model/transformer_v2.py:806-833 (forward_post: cross-attention only, quirk Q3), :380-390 (the
// per-layer `norm(output)` of return_intermediate), nn.MultiheadAttention(64, 4 heads) semantics as SURVEY 8a.
//
//   qin = tgt + query_pos ; Q = qin Wq^T + bq ; per head h: S_h = (Q_h / 4) K_h^T (26 keys) ; P_h = dropout(softmax(S_h))
//   O_h = P_h V_h ; y = O Wo^T + bo ; t1 = LN2(tgt + dropout2(y)) ; f = W2 dropout(relu(W1 t1 + b1)) + b2
//   out = LN3(t1 + dropout3(f)) ; inter = LN_final(out) ; attention weights = mean_h P_h (post-dropout, like torch)
//
// Round 1 ran this as ~14 launches per layer with every [tokens][64] intermediate making an HBM round trip, and the
// attention core (QK^T / softmax / PV) on CUDA cores.  Here a CTA owns two 128-token tiles (two 128-thread groups,
// each with its own operand tile and TMEM columns); every contraction -- the four 64x64 projections AND the per-head
// QK^T (M128 N32 K16) and PV (M128 N16 K32) -- is a tcgen05.mma with fp32 accumulation in TMEM and the repo's bf16
// hi/lo operand split (hi*hi + hi*lo + lo*hi: fp32 parity).  In the 32x32b TMEM layout one thread owns one token row, so
// softmax over the 26 keys, both LayerNorms, the residuals, biases, ReLU and all four dropouts are thread-local epilogue
// math between MMAs; the four weight matrices and the sample's projected K / V^T stay resident in shared memory.
// HBM traffic per layer: read tgt + query_pos, write out + inter (+ attention weights of the last layer); in training
// mode the tensors the (unfused) backward consumes are written as side outputs.  Dropout masks are bit-identical to
// tatt_dropout / tatt_mha64_* (same Philox counters), which regenerate them in the backward pass.
// No TMA multicast / clusters: nothing is shared between CTAs except 64 KB of weights (read once per CTA from L2).
#include <stdlib.h>
#include "tc_prims.cuh"

using namespace tcp;

namespace {

constexpr int DL_THREADS = 256;
constexpr int W_TILE = 64 * 128;            // [64][64] bf16 K-major SW128 = 8192 B
constexpr int K_TILE = 32 * 128;            // [32 keys][64 ch]
constexpr int A_TILE = 128 * 128;           // [128 rows][64 k]
constexpr int KV_BYTES = 2 * K_TILE + 2 * W_TILE;   // K hi, K lo, Vt hi, Vt lo = 24576
constexpr int OFF_W = 0;                              // Wq, Wo, W1, W2: each hi | lo
constexpr int OFF_KV = 4 * 2 * W_TILE;                // 65536
constexpr int OFF_A = OFF_KV + 2 * KV_BYTES;          // 114688
constexpr int OFF_STG = OFF_A + 2 * 2 * A_TILE;       // 180224
constexpr int DL_SMEM = OFF_STG + 8 * 4096 + 1024;    // 214016

// parameter vectors kept in shared memory (index * 64 floats)
enum { V_BQ = 0, V_BO, V_G2, V_B2N, V_B1, V_B2, V_G3, V_B3N, V_GF, V_BF, V_COUNT };

struct DecLayerParams {
  const float *tgt, *qpos, *Kp, *Vp;                  // [P][64], [P][64], [N][Lk][64], [N][Lk][64]
  const float *Wq, *Wo, *W1, *W2;                     // [64][64] each (row stride ldq for Wq: in_proj slice)
  const float* vec[V_COUNT];                          // bq, bo, ln2.g, ln2.b, b1, b2, ln3.g, ln3.b, lnf.g, lnf.b
  float *out, *inter, *aw;                            // [P][64], [P][64], [N][Lq][Lk] or null
  // training side outputs (all null in eval): qin, q, a, S2, t1, h1, h1d, S3 [P][64]; st2, st3, stf [2][P]
  float *s_qin, *s_q, *s_a, *s_S2, *s_t1, *s_h1, *s_h1d, *s_S3, *st2, *st3, *stf;
  const unsigned long long* rng;                      // {seed, counter} or null
  float p_attn, p2, p1, p3;                           // dropout probabilities (0 in eval)
  unsigned long long site_attn, site2, site1, site3;
  long long P;
  int N, Lq, Lk, ntiles;
};

// The tile loop of this kernel is one straight line of code that every warp executes once per tile; fully inlined it was
// 15.5 k SASS instructions (250 KB) against a 32 KB L1.5 instruction cache, and ncu showed "no instruction" as the top
// stall reason (9.5 stall cycles per issued instruction, SM issue slots 12 % busy).  Everything that repeats is therefore
// a real function call (Philox: 40 calls per token row; the coalesced half of the row load / store helpers) or a rolled
// loop (the four heads), so the body that streams from L2 is ~3x smaller and the callees stay cache-resident.
__device__ __noinline__ uint4 philox_call(uint2 key, uint4 ctr) { return philox4x32_10(key, ctr); }

__device__ __forceinline__ uint32_t sw128(int r, int k) {   // byte offset of bf16 element (r, k) in a K-major SW128 tile
  return (uint32_t)(r * 128 + ((((k >> 3) ^ (r & 7)) << 4) | ((k & 7) << 1)));
}

// warp-cooperative [32 rows][64] fp32 <-> one row per lane, through a warp-private 4 KB swizzled staging buffer
// (coalesced 128-byte global segments; conflict-free shared-memory accesses).  The global <-> staging halves are real
// calls (see the note on code size above); the register <-> staging halves must stay inline (register arrays).
__device__ __noinline__ void stage_in(const float* __restrict__ g, unsigned char* stg, int lane) {   // 32 rows x 32 floats
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int f = i * 32 + lane, r = f >> 3, c = f & 7;
    const float4 v = __ldg(reinterpret_cast<const float4*>(g + (long long)r * 64 + c * 4));
    *reinterpret_cast<float4*>(stg + r * 128 + ((c ^ (r & 7)) << 4)) = v;
  }
  __syncwarp();
}
__device__ __noinline__ void stage_out(float* __restrict__ g, const unsigned char* stg, int lane) {
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int f = i * 32 + lane, r = f >> 3, c = f & 7;
    *reinterpret_cast<float4*>(g + (long long)r * 64 + c * 4) =
        *reinterpret_cast<const float4*>(stg + r * 128 + ((c ^ (r & 7)) << 4));
  }
  __syncwarp();
}
__device__ __forceinline__ void load_rows(const float* __restrict__ g, float (&row)[64], unsigned char* stg, int lane) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    stage_in(g + h * 32, stg, lane);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float4 v = *reinterpret_cast<const float4*>(stg + lane * 128 + ((c ^ (lane & 7)) << 4));
      row[h * 32 + c * 4 + 0] = v.x; row[h * 32 + c * 4 + 1] = v.y; row[h * 32 + c * 4 + 2] = v.z; row[h * 32 + c * 4 + 3] = v.w;
    }
    __syncwarp();
  }
}
__device__ __forceinline__ void store_rows(float* __restrict__ g, const float (&row)[64], unsigned char* stg, int lane) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
#pragma unroll
    for (int c = 0; c < 8; ++c)
      *reinterpret_cast<float4*>(stg + lane * 128 + ((c ^ (lane & 7)) << 4)) =
          make_float4(row[h * 32 + c * 4], row[h * 32 + c * 4 + 1], row[h * 32 + c * 4 + 2], row[h * 32 + c * 4 + 3]);
    stage_out(g + h * 32, stg, lane);
  }
}

// 8 consecutive fp32 of a token row -> one 16-byte chunk of the bf16 hi tile and one of the lo tile.  Measured (N = 64,
// train + dropout): both helpers inline 368 us; projection issue as a call 379 us; this one as a call (56 sites per tile,
// 10 register arguments each) 409 us -- calls only pay where the callee is long (Philox) or off the register path
#ifndef DL_SPLIT_CALL
#define DL_SPLIT_CALL 0
#endif
#ifndef DL_PROJ_CALL
#define DL_PROJ_CALL 0
#endif
#if DL_SPLIT_CALL
#define DL_SPLIT_ATTR __noinline__
#else
#define DL_SPLIT_ATTR __forceinline__
#endif
#if DL_PROJ_CALL
#define DL_PROJ_ATTR __noinline__
#else
#define DL_PROJ_ATTR __forceinline__
#endif
__device__ DL_SPLIT_ATTR void split_store8(float x0, float x1, float x2, float x3, float x4, float x5, float x6, float x7,
                                          unsigned char* hi, unsigned char* lo) {
  uint2 h0, l0, h1, l1;
  split4(x0, x1, x2, x3, h0, l0);
  split4(x4, x5, x6, x7, h1, l1);
  *reinterpret_cast<uint4*>(hi) = make_uint4(h0.x, h0.y, h1.x, h1.y);
  *reinterpret_cast<uint4*>(lo) = make_uint4(l0.x, l0.y, l1.x, l1.y);
}

// one token row (64 fp32) -> bf16 hi / lo rows of the K-major SW128 operand tile
__device__ __forceinline__ void row_to_tile(const float (&row)[64], unsigned char* a_hi, unsigned char* a_lo, int r) {
#pragma unroll
  for (int c8 = 0; c8 < 8; ++c8) {
    const uint32_t off = (uint32_t)(r * 128 + ((c8 ^ (r & 7)) << 4));
    split_store8(row[c8 * 8], row[c8 * 8 + 1], row[c8 * 8 + 2], row[c8 * 8 + 3], row[c8 * 8 + 4], row[c8 * 8 + 5],
                 row[c8 * 8 + 6], row[c8 * 8 + 7], a_hi + off, a_lo + off);
  }
}

__device__ __forceinline__ void tmem_row64(uint32_t taddr, float (&row)[64]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t v[16];
    tmem_ld16_nowait(taddr + 16 * q, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) row[16 * q + i] = __uint_as_float(v[i]);
  }
}

// the same mask as dropout_kernel (elementwise.cu): element e of the flat [P][64] tensor uses the 16-bit lane (e & 7) of
// Philox counter e >> 3, i.e. 8 calls per token row
__device__ __forceinline__ void row_dropout(float (&row)[64], float p, unsigned long long seed, unsigned long long offset,
                                            long long t) {
  const float scale = 1.f / (1.f - p);
  const uint32_t thr = (uint32_t)(p * 65536.f + 0.5f);
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
#pragma unroll
  for (int c8 = 0; c8 < 8; ++c8) {
    const unsigned long long idx = (unsigned long long)(t * 8 + c8);
    const uint4 r = philox_call(key, make_uint4((uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)offset,
                                                (uint32_t)(offset >> 32)));
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      row[c8 * 8 + 2 * k] *= (w[k] & 0xffffu) >= thr ? scale : 0.f;
      row[c8 * 8 + 2 * k + 1] *= (w[k] >> 16) >= thr ? scale : 0.f;
    }
  }
}

// y = LN(row) * g + b over the 64 channels of one token (same formulas as ln_fwd_kernel, norm.cu)
__device__ __forceinline__ void row_layernorm(const float (&x)[64], float (&y)[64], const float* g, const float* b, float& mean,
                                              float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 64; ++i) s += x[i];
  const float m = s * (1.f / 64.f);
  float v = 0.f;
#pragma unroll
  for (int i = 0; i < 64; ++i) {
    const float d = x[i] - m;
    v = fmaf(d, d, v);
  }
  const float rs = 1.f / sqrtf(v * (1.f / 64.f) + 1e-5f);
#pragma unroll
  for (int i = 0; i < 64; ++i) y[i] = (x[i] - m) * rs * g[i] + b[i];
  mean = m;
  rstd = rs;
}

// a [64 x 64] projection: D[:, 0:64) = A W^T with the three hi/lo products (issued by one warp; a real call, see above)
__device__ DL_PROJ_ATTR void proj_issue(uint32_t tmem_g, uint32_t dA_hi, uint32_t dA_lo, uint32_t wh, uint32_t wl) {
  constexpr uint32_t id64 = idesc_bf16(128, 64);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    umma_elect(tmem_g, dA_hi + 2 * k, wh + 2 * k, id64, k > 0 ? 1u : 0u);
    umma_elect(tmem_g, dA_hi + 2 * k, wl + 2 * k, id64, 1u);
    umma_elect(tmem_g, dA_lo + 2 * k, wh + 2 * k, id64, 1u);
  }
}

template <bool TRAIN>
__global__ void __launch_bounds__(DL_THREADS, 1) tp_declayer_fwd_kernel(const DecLayerParams p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(16) float vec[V_COUNT][64];
  __shared__ __align__(8) unsigned long long mbar[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = warp >> 2, gw = warp & 3;         // group 0 / 1; warp within the group == TMEM lane group
  const int r = gw * 32 + lane;                     // token row of this thread inside its tile

  // ---- one-time setup: weights -> bf16 hi/lo K-major tiles, parameter vectors, barriers, TMEM
  {
    const float* Ws[4] = {p.Wq, p.Wo, p.W1, p.W2};
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      unsigned char* hi = smem + OFF_W + w * 2 * W_TILE;
      unsigned char* lo = hi + W_TILE;
      for (int i = tid; i < 64 * 16; i += DL_THREADS) {          // [n][c4] float4 pieces
        const int n = i >> 4, c4 = i & 15;
        const float4 v = __ldg(reinterpret_cast<const float4*>(Ws[w] + n * 64 + c4 * 4));
        uint2 h, l;
        split4(v.x, v.y, v.z, v.w, h, l);
        const uint32_t off = (uint32_t)(n * 128 + (((c4 >> 1) ^ (n & 7)) << 4) + (c4 & 1) * 8);
        *reinterpret_cast<uint2*>(hi + off) = h;
        *reinterpret_cast<uint2*>(lo + off) = l;
      }
    }
    for (int i = tid; i < V_COUNT * 64; i += DL_THREADS) vec[i >> 6][i & 63] = __ldg(p.vec[i >> 6] + (i & 63));
    for (int i = tid; i < 2 * KV_BYTES / 16; i += DL_THREADS)      // zero K / V^T tiles: padded keys must read as 0
      reinterpret_cast<uint4*>(smem + OFF_KV)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
      mbar_init(smem_u32(&mbar[0]), 1);
      mbar_init(smem_u32(&mbar[1]), 1);
      mbar_fence_init();
    }
    __syncwarp();
    if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 256);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  const uint32_t tmem_g = tmem_base_s + (uint32_t)grp * 128u;            // this group's 128 columns
  const uint32_t trow = tmem_g + ((uint32_t)(gw * 32) << 16);            // + this warp's lanes
  unsigned char* a_hi = smem + OFF_A + grp * 2 * A_TILE;
  unsigned char* a_lo = a_hi + A_TILE;
  unsigned char* kv = smem + OFF_KV + grp * KV_BYTES;                    // K hi | K lo | Vt hi | Vt lo
  unsigned char* stg = smem + OFF_STG + warp * 4096;
  const uint32_t bar = smem_u32(&mbar[grp]);
  uint32_t par = 0;
  const uint32_t dA_hi = desc_lo(smem_u32(a_hi)), dA_lo = desc_lo(smem_u32(a_lo));
  const uint32_t dK_hi = desc_lo(smem_u32(kv)), dK_lo = desc_lo(smem_u32(kv + K_TILE));
  const uint32_t dV_hi = desc_lo(smem_u32(kv + 2 * K_TILE)), dV_lo = desc_lo(smem_u32(kv + 2 * K_TILE + W_TILE));
  constexpr uint32_t id32 = idesc_bf16(128, 32), id16 = idesc_bf16(128, 16);
  unsigned long long seed = 0, ctr = 0;
  if (TRAIN && p.rng) {
    seed = p.rng[0];
    ctr = p.rng[1] * 65536ull;
  }
  int cur_n = -1;

#define GROUP_SYNC() asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory")
  // publish the operand tile written by this group's threads, let warp 0 of the group issue, wait for completion
#define PHASE_BEGIN()            \
  fence_proxy_async_smem();      \
  tc_fence_before();             \
  GROUP_SYNC();                  \
  if (gw == 0) {                 \
    tc_fence_after();
#define PHASE_END()              \
    umma_commit_elect(bar);      \
  }                              \
  mbar_wait(bar, par);           \
  par ^= 1;                      \
  tc_fence_after();

  auto proj = [&](int w) {
    const uint32_t wh = desc_lo(smem_u32(smem + OFF_W + w * 2 * W_TILE));
    proj_issue(tmem_g, dA_hi, dA_lo, wh, wh + (W_TILE >> 4));
  };

  for (int pair = blockIdx.x; 2 * pair + grp < p.ntiles; pair += gridDim.x) {
    const int tile = 2 * pair + grp;
    const long long t0 = (long long)tile * 128;                 // first token of the tile
    const long long t = t0 + r;                                  // this thread's token
    const int n = (int)(t0 / p.Lq);
    const float* g_tgt = p.tgt + (t0 + gw * 32) * 64;
    const long long woff = (t0 + gw * 32) * 64;                 // this warp's 32 rows in any [P][64] tensor

    // ---- projected keys / values of sample n -> K tile [32 keys][64 ch], V^T tile [64 ch][keys]
    if (n != cur_n) {
      cur_n = n;
      const float* Kn = p.Kp + (long long)n * p.Lk * 64;
      const float* Vn = p.Vp + (long long)n * p.Lk * 64;
      for (int i = r; i < p.Lk * 16; i += 128) {
        const int j = i >> 4, c4 = i & 15;
        const float4 kq = __ldg(reinterpret_cast<const float4*>(Kn + j * 64 + c4 * 4));
        uint2 h, l;
        split4(kq.x, kq.y, kq.z, kq.w, h, l);
        const uint32_t off = (uint32_t)(j * 128 + (((c4 >> 1) ^ (j & 7)) << 4) + (c4 & 1) * 8);
        *reinterpret_cast<uint2*>(kv + off) = h;
        *reinterpret_cast<uint2*>(kv + K_TILE + off) = l;
        const float4 vq = __ldg(reinterpret_cast<const float4*>(Vn + j * 64 + c4 * 4));
        const float vv[4] = {vq.x, vq.y, vq.z, vq.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int d = c4 * 4 + e;
          const __nv_bfloat16 vh = __float2bfloat16_rn(vv[e]);
          const __nv_bfloat16 vl = __float2bfloat16_rn(vv[e] - __bfloat162float(vh));
          const uint32_t o2 = sw128(d, j);
          *reinterpret_cast<__nv_bfloat16*>(kv + 2 * K_TILE + o2) = vh;
          *reinterpret_cast<__nv_bfloat16*>(kv + 2 * K_TILE + W_TILE + o2) = vl;
        }
      }
    }

    float cur[64];
    // ---- phase 1: qin = tgt + qpos ; Q = qin Wq^T
    {
      float q2[64];
      load_rows(g_tgt, cur, stg, lane);
      load_rows(p.qpos + woff, q2, stg, lane);
#pragma unroll
      for (int i = 0; i < 64; ++i) cur[i] += q2[i];
      if (TRAIN) store_rows(p.s_qin + woff, cur, stg, lane);
      row_to_tile(cur, a_hi, a_lo, r);
    }
    PHASE_BEGIN()
    proj(0);
    PHASE_END()
    tmem_row64(trow, cur);
#pragma unroll
    for (int i = 0; i < 64; ++i) cur[i] += vec[V_BQ][i];
    if (TRAIN) store_rows(p.s_q + woff, cur, stg, lane);
#pragma unroll
    for (int i = 0; i < 64; ++i) cur[i] *= 0.25f;                // 1 / sqrt(head_dim)
    row_to_tile(cur, a_hi, a_lo, r);

    // ---- phase 2: S_h = Q_h K_h^T for the 4 heads -> TMEM columns [32 h, 32 h + 32)
    PHASE_BEGIN()
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      umma_elect(tmem_g + 32 * h, dA_hi + 2 * h, dK_hi + 2 * h, id32, 0u);
      umma_elect(tmem_g + 32 * h, dA_hi + 2 * h, dK_lo + 2 * h, id32, 1u);
      umma_elect(tmem_g + 32 * h, dA_lo + 2 * h, dK_hi + 2 * h, id32, 1u);
    }
    PHASE_END()

    // ---- phases 3a / 3b: softmax (+ dropout) per head, P_h -> operand tile, O_h = P_h V_h (two heads per round)
    float aw[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) aw[j] = 0.f;
#pragma unroll 1
    for (int rd = 0; rd < 2; ++rd) {
#pragma unroll 1
      for (int hh = 0; hh < 2; ++hh) {
        const int h = 2 * rd + hh;
        float s[32];
        {
          uint32_t v0[16], v1[16];
          tmem_ld16_nowait(trow + 32 * h, v0);
          tmem_ld16_nowait(trow + 32 * h + 16, v1);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            s[j] = __uint_as_float(v0[j]);
            s[16 + j] = __uint_as_float(v1[j]);
          }
        }
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (j >= p.Lk) s[j] = -INFINITY;
          mx = fmaxf(mx, s[j]);
        }
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float e = (j < p.Lk) ? __expf(s[j] - mx) : 0.f;
          s[j] = e;
          sum += e;
        }
        const float inv = 1.f / sum;
#pragma unroll
        for (int j = 0; j < 32; ++j) s[j] *= inv;
        if (TRAIN && p.p_attn > 0.f) {                 // same bits as drop_scales() in attn.cu
          const float sc = 1.f / (1.f - p.p_attn);
          const uint32_t thr = (uint32_t)(p.p_attn * 65536.f + 0.5f);
          const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
          const unsigned long long offset = ctr + p.site_attn;
          const long long elem = ((long long)n * 4 + h) * p.Lq + (t - (long long)n * p.Lq);
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) {
            const unsigned long long idx = (unsigned long long)(elem * 4 + j8);
            const uint4 rr = philox_call(key, make_uint4((uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)offset,
                                                         (uint32_t)(offset >> 32)));
            const uint32_t w4[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              s[j8 * 8 + 2 * k] *= (w4[k] & 0xffffu) >= thr ? sc : 0.f;
              s[j8 * 8 + 2 * k + 1] *= (w4[k] >> 16) >= thr ? sc : 0.f;
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) aw[j] += 0.25f * s[j];
        // P_h -> k-columns [32 hh, 32 hh + 32) of the operand tile
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          const uint32_t off = (uint32_t)(r * 128 + (((hh * 4 + c8) ^ (r & 7)) << 4));
          split_store8(s[c8 * 8], s[c8 * 8 + 1], s[c8 * 8 + 2], s[c8 * 8 + 3], s[c8 * 8 + 4], s[c8 * 8 + 5], s[c8 * 8 + 6],
                       s[c8 * 8 + 7], a_hi + off, a_lo + off);
        }
      }
      PHASE_BEGIN()
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int h = 2 * rd + hh;
        const uint32_t vh = dV_hi + (uint32_t)((16 * h * 128) >> 4), vl = dV_lo + (uint32_t)((16 * h * 128) >> 4);
#pragma unroll
        for (int k = 0; k < 2; ++k) {                            // keys 0..15, 16..31
          const uint32_t ak = (uint32_t)(2 * (2 * hh + k));
          umma_elect(tmem_g + 16 * h, dA_hi + ak, vh + 2 * k, id16, k > 0 ? 1u : 0u);
          umma_elect(tmem_g + 16 * h, dA_hi + ak, vl + 2 * k, id16, 1u);
          umma_elect(tmem_g + 16 * h, dA_lo + ak, vh + 2 * k, id16, 1u);
        }
      }
      PHASE_END()
    }
    if (p.aw) {                                                  // head-averaged weights [N][Lq][Lk]: rows are contiguous
      float* sw = reinterpret_cast<float*>(stg);
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < p.Lk) sw[lane * p.Lk + j] = aw[j];
      __syncwarp();
      float* dst = p.aw + (t0 + gw * 32) * p.Lk;
      for (int i = lane; i < 32 * p.Lk; i += 32) dst[i] = sw[i];
      __syncwarp();
    }

    // ---- phase 4: y = O Wo^T + bo ; s2 = tgt + dropout2(y) ; t1 = LN2(s2)
    tmem_row64(trow, cur);
    if (TRAIN) store_rows(p.s_a + woff, cur, stg, lane);
    row_to_tile(cur, a_hi, a_lo, r);
    PHASE_BEGIN()
    proj(1);
    PHASE_END()
    float t1[64];
    {
      tmem_row64(trow, cur);
#pragma unroll
      for (int i = 0; i < 64; ++i) cur[i] += vec[V_BO][i];
      if (TRAIN && p.p2 > 0.f) row_dropout(cur, p.p2, seed, ctr + p.site2, t);
      load_rows(g_tgt, t1, stg, lane);                           // residual (L2-hot re-read instead of 64 live registers)
#pragma unroll
      for (int i = 0; i < 64; ++i) cur[i] += t1[i];
      if (TRAIN) store_rows(p.s_S2 + woff, cur, stg, lane);
      float m, rs;
      row_layernorm(cur, t1, vec[V_G2], vec[V_B2N], m, rs);
      if (TRAIN) {
        p.st2[t] = m;
        p.st2[p.P + t] = rs;
        store_rows(p.s_t1 + woff, t1, stg, lane);
      }
      row_to_tile(t1, a_hi, a_lo, r);
    }

    // ---- phase 5: h1 = relu(t1 W1^T + b1) ; phase 6: f = dropout(h1) W2^T + b2
    PHASE_BEGIN()
    proj(2);
    PHASE_END()
    tmem_row64(trow, cur);
#pragma unroll
    for (int i = 0; i < 64; ++i) cur[i] = fmaxf(cur[i] + vec[V_B1][i], 0.f);
    if (TRAIN) store_rows(p.s_h1 + woff, cur, stg, lane);
    if (TRAIN && p.p1 > 0.f) row_dropout(cur, p.p1, seed, ctr + p.site1, t);
    if (TRAIN) store_rows(p.s_h1d + woff, cur, stg, lane);
    row_to_tile(cur, a_hi, a_lo, r);
    PHASE_BEGIN()
    proj(3);
    PHASE_END()
    tmem_row64(trow, cur);
#pragma unroll
    for (int i = 0; i < 64; ++i) cur[i] += vec[V_B2][i];
    if (TRAIN && p.p3 > 0.f) row_dropout(cur, p.p3, seed, ctr + p.site3, t);
#pragma unroll
    for (int i = 0; i < 64; ++i) cur[i] += t1[i];
    if (TRAIN) store_rows(p.s_S3 + woff, cur, stg, lane);
    {
      float m, rs;
      row_layernorm(cur, t1, vec[V_G3], vec[V_B3N], m, rs);      // t1 := out
      if (TRAIN) {
        p.st3[t] = m;
        p.st3[p.P + t] = rs;
      }
      store_rows(p.out + woff, t1, stg, lane);
      row_layernorm(t1, cur, vec[V_GF], vec[V_BF], m, rs);       // cur := LN_final(out)
      if (TRAIN) {
        p.stf[t] = m;
        p.stf[p.P + t] = rs;
      }
      store_rows(p.inter + woff, cur, stg, lane);
    }
  }
#undef PHASE_BEGIN
#undef PHASE_END
#undef GROUP_SYNC
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base_s, 256);
}

}  // namespace

extern "C" {

// in[]  : tgt, qpos, Kp, Vp, Wq, Wo, W1, W2, bq, bo, ln2.g, ln2.b, b1, b2, ln3.g, ln3.b, lnf.g, lnf.b      (18 pointers)
// out[] : out, inter, aw|null, then (train == 1) qin, q, a, S2, t1, h1, h1d, S3, st2, st3, stf             (3 + 11 pointers)
// pdrop[4] / sites[4]: attention, dropout2 (after out-proj), dropout (FFN hidden), dropout3 (FFN out)
int tatt_tp_declayer_fwd(const float* const* in, float* const* out, int train, int N, int Lq, int Lk,
                         const float* pdrop, const unsigned long long* rng, const unsigned long long* sites,
                         void* stream) {
  TATT_REQUIRE(N >= 1 && Lq % 128 == 0 && Lk >= 1 && Lk <= 32, "tp_declayer_fwd: needs Lq %% 128 == 0 and Lk <= 32 (Lq=%d Lk=%d)",
               Lq, Lk);
  for (int i = 0; i < 18; ++i) TATT_REQUIRE(in[i] != nullptr, "tp_declayer_fwd: input %d is null", i);
  TATT_REQUIRE(out[0] && out[1], "tp_declayer_fwd: out / inter must be given");
  DecLayerParams p{};
  p.tgt = in[0]; p.qpos = in[1]; p.Kp = in[2]; p.Vp = in[3];
  p.Wq = in[4]; p.Wo = in[5]; p.W1 = in[6]; p.W2 = in[7];
  for (int i = 0; i < V_COUNT; ++i) p.vec[i] = in[8 + i];
  p.out = out[0]; p.inter = out[1]; p.aw = out[2];
  if (train) {
    for (int i = 3; i < 14; ++i) TATT_REQUIRE(out[i] != nullptr, "tp_declayer_fwd: training side output %d is null", i);
    p.s_qin = out[3]; p.s_q = out[4]; p.s_a = out[5]; p.s_S2 = out[6]; p.s_t1 = out[7]; p.s_h1 = out[8];
    p.s_h1d = out[9]; p.s_S3 = out[10]; p.st2 = out[11]; p.st3 = out[12]; p.stf = out[13];
    const bool any = pdrop && (pdrop[0] > 0.f || pdrop[1] > 0.f || pdrop[2] > 0.f || pdrop[3] > 0.f);
    TATT_REQUIRE(!any || (rng && sites), "tp_declayer_fwd: dropout needs the rng state and the 4 site ids");
    if (any) {
      p.rng = rng;
      p.p_attn = pdrop[0]; p.p2 = pdrop[1]; p.p1 = pdrop[2]; p.p3 = pdrop[3];
      p.site_attn = sites[0]; p.site2 = sites[1]; p.site1 = sites[2]; p.site3 = sites[3];
    }
  }
  p.N = N; p.Lq = Lq; p.Lk = Lk;
  p.P = (long long)N * Lq;
  p.ntiles = (int)(p.P / 128);
  int dev = 0, nsm = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  const int npairs = (p.ntiles + 1) / 2;
  const int grid = npairs < nsm ? npairs : nsm;
  cudaStream_t st = (cudaStream_t)stream;
  if (train) {
    TATT_CUDA(cudaFuncSetAttribute(tp_declayer_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DL_SMEM));
    tp_declayer_fwd_kernel<true><<<grid, DL_THREADS, DL_SMEM, st>>>(p);
  } else {
    TATT_CUDA(cudaFuncSetAttribute(tp_declayer_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, DL_SMEM));
    tp_declayer_fwd_kernel<false><<<grid, DL_THREADS, DL_SMEM, st>>>(p);
  }
  TATT_LAUNCH_CHECK("tp_declayer_fwd_kernel");
  return 0;
}

}  // extern "C"
