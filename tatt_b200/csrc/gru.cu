// GRU recurrences.
//  (1) SRB BiGRU, hidden 32 (GruBlock, model/tsrn.py:1067-1084): one warp per (sequence, direction),
//      h in registers, W_hh rows in registers, strided row addressing so that vertical (T=H) and
//      horizontal (T=W) scans run directly on the NHWC feature map -- the reference's
//      permute/contiguous copies (tsrn.py:1076-1083, 906) do not exist here.
//  (2) RPE BiGRU of the TP Interpreter (model/transformer_v2.py:177, 215-221; quirk Q1: recurrence
//      over the BATCH axis): per-step gate kernels around the batched recurrent GEMM (gemm.cu).
// Gate order r,z,n; n = tanh(gi_n + r * (W_hn h + b_hn)); h' = (1-z) n + z h   (torch.nn.GRU).
#include <cuda_bf16.h>
#include <stdlib.h>
#include "common.cuh"

namespace {

__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16* hi, __nv_bfloat16* lo, long long i) {
  __nv_bfloat16 h = __float2bfloat16_rn(x);
  hi[i] = h;
  lo[i] = __float2bfloat16_rn(x - __bfloat162float(h));
}

// fast activations for the recurrences: MUFU ex2/rcp based, abs error ~1e-6 (tolerance of the path: 1e-3)
__device__ __forceinline__ float fsigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float ftanh(float x) { return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x)); }

// GI    [rows][192]  = [dir][gate][32]   (x W_ih^T + b_ih, both directions)
// OUT   [rows][64]   = [dir][32]
// GATES [rows][320]  = [dir][r,z,n,ghn,hprev][32]   (saved for the backward; may be NULL)
// Each warp advances TWO sequences of the same direction in lock-step: the W_hh rows in registers are shared and
// the two independent dependency chains double the ILP of this latency-bound recurrence.
__global__ void __launch_bounds__(128)
gru32_scan_fwd_kernel(const float* __restrict__ GI, const float* __restrict__ Whh,
                      const float* __restrict__ bhh, float* __restrict__ OUT, float* __restrict__ GATES,
                      int nseq, int T, int s_inner, long long outer_stride, long long inner_stride,
                      long long t_stride) {
  __shared__ __align__(16) float hs[4][2][32];   // per-warp broadcast buffers for the two hidden states
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int dir = (int)(gw & 1);
  const long long seq0 = (gw >> 1) * 2;
  if (seq0 >= nseq) return;
  const bool two = seq0 + 1 < nseq;
  const long long seq1 = two ? seq0 + 1 : seq0;
  const long long base[2] = {(seq0 / s_inner) * outer_stride + (seq0 % s_inner) * inner_stride,
                             (seq1 / s_inner) * outer_stride + (seq1 % s_inner) * inner_stride};

  float wr[32], wz[32], wn[32];
  const float* W = Whh + dir * 96 * 32;
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    wr[k] = W[(lane)*32 + k];
    wz[k] = W[(32 + lane) * 32 + k];
    wn[k] = W[(64 + lane) * 32 + k];
  }
  const float br = bhh[dir * 96 + lane], bz = bhh[dir * 96 + 32 + lane], bn = bhh[dir * 96 + 64 + lane];

  float h[2] = {0.f, 0.f};
  long long off = (dir == 0 ? 0 : (long long)(T - 1)) * t_stride;     // row offset shared by both sequences
  const long long step_stride = dir == 0 ? t_stride : -t_stride;
  constexpr int PF = 4;
  float pr[2][PF], pz[2][PF], pn[2][PF];
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int j = 0; j < PF; ++j) {
      pr[u][j] = pz[u][j] = pn[u][j] = 0.f;
      if (j < T) {
        const float* gp = GI + (base[u] + off + j * step_stride) * 192 + dir * 96 + lane;
        pr[u][j] = gp[0];
        pz[u][j] = gp[32];
        pn[u][j] = gp[64];
      }
    }
  for (int s0 = 0; s0 < T; s0 += PF) {
#pragma unroll
    for (int j = 0; j < PF; ++j) {
      const int s = s0 + j;
      if (s < T) {
        float gr[2], gz[2], gn[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          gr[u] = pr[u][j];
          gz[u] = pz[u][j];
          gn[u] = pn[u][j];
          if (s + PF < T) {
            const float* np = GI + (base[u] + off + PF * step_stride) * 192 + dir * 96 + lane;
            pr[u][j] = np[0];
            pz[u][j] = np[32];
            pn[u][j] = np[64];
          }
        }
        __syncwarp();
        hs[wib][0][lane] = h[0];
        hs[wib][1][lane] = h[1];
        __syncwarp();
        float ar[2] = {br, br}, az[2] = {bz, bz}, an[2] = {bn, bn};
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
          const float4 a = *reinterpret_cast<const float4*>(&hs[wib][0][k4 * 4]);
          const float4 b = *reinterpret_cast<const float4*>(&hs[wib][1][k4 * 4]);
          const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            ar[0] = fmaf(wr[k4 * 4 + e], av[e], ar[0]);
            az[0] = fmaf(wz[k4 * 4 + e], av[e], az[0]);
            an[0] = fmaf(wn[k4 * 4 + e], av[e], an[0]);
            ar[1] = fmaf(wr[k4 * 4 + e], bv[e], ar[1]);
            az[1] = fmaf(wz[k4 * 4 + e], bv[e], az[1]);
            an[1] = fmaf(wn[k4 * 4 + e], bv[e], an[1]);
          }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          float r = fsigmoid(gr[u] + ar[u]);
          float z = fsigmoid(gz[u] + az[u]);
          float n = ftanh(gn[u] + r * an[u]);
          float hn = n + z * (h[u] - n);
          if (u == 0 || two) {
            const long long row = base[u] + off;
            OUT[row * 64 + dir * 32 + lane] = hn;
            if (GATES) {
              float* g = GATES + row * 320 + dir * 160 + lane;
              g[0] = r;
              g[32] = z;
              g[64] = n;
              g[96] = an[u];
              g[128] = h[u];
            }
          }
          h[u] = hn;
        }
        off += step_stride;
      }
    }
  }
}

// dGI [rows][192] (grad wrt the input projections), dGH [rows][192] (grad wrt W_hh h + b_hh); two sequences per warp
__global__ void __launch_bounds__(128)
gru32_scan_bwd_kernel(const float* __restrict__ dOUT, const float* __restrict__ GATES,
                      const float* __restrict__ Whh, float* __restrict__ dGI, float* __restrict__ dGH, int nseq,
                      int T, int s_inner, long long outer_stride, long long inner_stride, long long t_stride) {
  __shared__ __align__(16) float gs[4][2][96];   // per-warp broadcast buffers: dpr | dpz | dhn of both sequences
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int dir = (int)(gw & 1);
  const long long seq0 = (gw >> 1) * 2;
  if (seq0 >= nseq) return;
  const bool two = seq0 + 1 < nseq;
  const long long seq1 = two ? seq0 + 1 : seq0;
  const long long base[2] = {(seq0 / s_inner) * outer_stride + (seq0 % s_inner) * inner_stride,
                             (seq1 / s_inner) * outer_stride + (seq1 % s_inner) * inner_stride};

  float wc[96];  // column `lane` of W_hh[dir]
  const float* W = Whh + dir * 96 * 32;
#pragma unroll
  for (int i = 0; i < 96; ++i) wc[i] = W[i * 32 + lane];

  float dh[2] = {0.f, 0.f};
  long long off = (dir == 0 ? (long long)(T - 1) : 0) * t_stride;   // walk the recurrence backwards
  const long long step_stride = dir == 0 ? -t_stride : t_stride;
  constexpr int PF = 2;
  float qr[2][PF], qz[2][PF], qn[2][PF], qg[2][PF], qh[2][PF], qo[2][PF];
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int j = 0; j < PF; ++j) {
      qr[u][j] = qz[u][j] = qn[u][j] = qg[u][j] = qh[u][j] = qo[u][j] = 0.f;
      if (j < T) {
        const long long rj = base[u] + off + j * step_stride;
        const float* g = GATES + rj * 320 + dir * 160 + lane;
        qr[u][j] = g[0]; qz[u][j] = g[32]; qn[u][j] = g[64]; qg[u][j] = g[96]; qh[u][j] = g[128];
        qo[u][j] = dOUT[rj * 64 + dir * 32 + lane];
      }
    }
  for (int s0 = 0; s0 < T; s0 += PF) {
#pragma unroll
    for (int j = 0; j < PF; ++j) {
      const int s = s0 + j;
      if (s < T) {
        float gz_[2], go_[2];
        __syncwarp();
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const float r = qr[u][j], z = qz[u][j], n = qn[u][j], ghn = qg[u][j], hp = qh[u][j];
          const float go = qo[u][j] + dh[u];
          if (s + PF < T) {
            const long long rj = base[u] + off + PF * step_stride;
            const float* g = GATES + rj * 320 + dir * 160 + lane;
            qr[u][j] = g[0]; qz[u][j] = g[32]; qn[u][j] = g[64]; qg[u][j] = g[96]; qh[u][j] = g[128];
            qo[u][j] = dOUT[rj * 64 + dir * 32 + lane];
          }
          float dn = go * (1.f - z);
          float dz = go * (hp - n);
          float dpn = dn * (1.f - n * n);
          float dpr = dpn * ghn * r * (1.f - r);
          float dpz = dz * z * (1.f - z);
          float dhn = dpn * r;
          if (u == 0 || two) {
            const long long row = base[u] + off;
            float* o = dGI + row * 192 + dir * 96 + lane;
            o[0] = dpr;
            o[32] = dpz;
            o[64] = dpn;
            float* o2 = dGH + row * 192 + dir * 96 + lane;
            o2[0] = dpr;
            o2[32] = dpz;
            o2[64] = dhn;
          }
          gs[wib][u][lane] = dpr;
          gs[wib][u][32 + lane] = dpz;
          gs[wib][u][64 + lane] = dhn;
          gz_[u] = z;
          go_[u] = go;
        }
        __syncwarp();
        float acc[2] = {go_[0] * gz_[0], go_[1] * gz_[1]}, acc2[2] = {0.f, 0.f}, acc3[2] = {0.f, 0.f};
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const float4 a = *reinterpret_cast<const float4*>(&gs[wib][u][k4 * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&gs[wib][u][32 + k4 * 4]);
            const float4 c = *reinterpret_cast<const float4*>(&gs[wib][u][64 + k4 * 4]);
            acc[u] = fmaf(wc[k4 * 4 + 0], a.x, acc[u]);
            acc2[u] = fmaf(wc[32 + k4 * 4 + 0], b.x, acc2[u]);
            acc3[u] = fmaf(wc[64 + k4 * 4 + 0], c.x, acc3[u]);
            acc[u] = fmaf(wc[k4 * 4 + 1], a.y, acc[u]);
            acc2[u] = fmaf(wc[32 + k4 * 4 + 1], b.y, acc2[u]);
            acc3[u] = fmaf(wc[64 + k4 * 4 + 1], c.y, acc3[u]);
            acc[u] = fmaf(wc[k4 * 4 + 2], a.z, acc[u]);
            acc2[u] = fmaf(wc[32 + k4 * 4 + 2], b.z, acc2[u]);
            acc3[u] = fmaf(wc[64 + k4 * 4 + 2], c.z, acc3[u]);
            acc[u] = fmaf(wc[k4 * 4 + 3], a.w, acc[u]);
            acc2[u] = fmaf(wc[32 + k4 * 4 + 3], b.w, acc2[u]);
            acc3[u] = fmaf(wc[64 + k4 * 4 + 3], c.w, acc3[u]);
          }
        }
        dh[0] = acc[0] + acc2[0] + acc3[0];
        dh[1] = acc[1] + acc2[1] + acc3[1];
        off += step_stride;
      }
    }
  }
}

// ---------------------------------------------------------------------------- RPE (batch-axis GRU)
// X[w][h*C + c] = init_factor[(h*W + w)*C + c]
__global__ void rpe_gather_kernel(const float* __restrict__ emb, float* __restrict__ X, int H, int W, int C) {
  long long n = (long long)H * W * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    long long r = i / C;
    int h = (int)(r % H);
    int w = (int)(r / H);
    X[i] = emb[((long long)h * W + w) * C + c];
  }
}
__global__ void rpe_scatter_kernel(const float* __restrict__ dX, float* __restrict__ demb, int H, int W, int C) {
  long long n = (long long)H * W * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    long long r = i / C;
    int h = (int)(r % H);
    int w = (int)(r / H);
    demb[((long long)h * W + w) * C + c] = dX[i];
  }
}

// One recurrence step for both directions.  Thread per (dir, w, j).
//  GI   [2][Wd][3Hd]  (constant over steps: the GRU input is identical at every step)
//  GH   [2][Wd][3Hd]  = h_prev W_hh^T + b_hh   (this step)
//  HALL [2][N+1][Wd][Hd]  hidden states, HALL[:,0] == 0
//  GATES[N][2][Wd][4][Hd]  r,z,n,ghn (saved)       QPOS [N][Himg*Wd][C]
// Both gate kernels handle FOUR consecutive hidden units per thread (float4 loads / stores, 8-byte bf16 plane stores):
// one recurrence step is only 2*Wd*Hd = 262k elements, so the kernel is pure latency -- fewer, fatter threads with
// independent vector loads cut it roughly in half.  Hd % 4 == 0 and C % 4 == 0 (checked by the launchers).
__device__ __forceinline__ void split4_store(const float (&x)[4], __nv_bfloat16* hi, __nv_bfloat16* lo, long long i) {
  const __nv_bfloat162 h0 = __floats2bfloat162_rn(x[0], x[1]), h1 = __floats2bfloat162_rn(x[2], x[3]);
  const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
  const __nv_bfloat162 l0 = __floats2bfloat162_rn(x[0] - f0.x, x[1] - f0.y), l1 = __floats2bfloat162_rn(x[2] - f1.x, x[3] - f1.y);
  *reinterpret_cast<uint2*>(hi + i) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
  *reinterpret_cast<uint2*>(lo + i) = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
}
__device__ __forceinline__ void ld4(const float* p, float (&v)[4]) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void st4(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}

__global__ void rpe_gate_fwd_kernel(const float* __restrict__ GI, const float* __restrict__ GH,
                                    float* __restrict__ HALL, float* __restrict__ GATES,
                                    float* __restrict__ QPOS, __nv_bfloat16* __restrict__ HPL, long long plane_lo,
                                    int step, int N, int Wd, int Hd, int C, int Himg) {
  const int Hd4 = Hd >> 2;
  const int n4 = 2 * Wd * Hd4;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
    const int j = (i % Hd4) << 2;
    const int r = i / Hd4;
    const int w = r % Wd, dir = r / Wd;
    const long long g3 = ((long long)dir * Wd + w) * 3 * Hd + j;
    const long long hidx = (((long long)dir * (N + 1) + step) * Wd + w) * Hd + j;
    float gir[4], giz[4], gin[4], ghr[4], ghz[4], ghn[4], hp[4];
    ld4(GI + g3, gir); ld4(GI + g3 + Hd, giz); ld4(GI + g3 + 2 * Hd, gin);
    ld4(GH + g3, ghr); ld4(GH + g3 + Hd, ghz); ld4(GH + g3 + 2 * Hd, ghn);
    ld4(HALL + hidx, hp);
    float rr[4], zz[4], nn[4], hn[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      rr[k] = sigmoid_f(gir[k] + ghr[k]);
      zz[k] = sigmoid_f(giz[k] + ghz[k]);
      nn[k] = tanhf(gin[k] + rr[k] * ghn[k]);
      hn[k] = nn[k] + zz[k] * (hp[k] - nn[k]);
    }
    st4(HALL + hidx + (long long)Wd * Hd, hn);
    if (HPL) split4_store(hn, HPL, HPL + plane_lo, hidx + (long long)Wd * Hd);
    if (GATES) {
      const long long gi = ((((long long)step * 2 + dir) * Wd + w) * 4) * Hd + j;
      st4(GATES + gi, rr);
      st4(GATES + gi + Hd, zz);
      st4(GATES + gi + 2 * Hd, nn);
      st4(GATES + gi + 3 * Hd, ghn);
    }
    const int b = dir == 0 ? step : N - 1 - step;
    const int f = dir * Hd + j;
    const int hh = f / C, c = f % C;
    st4(QPOS + ((long long)b * Himg * Wd + (long long)hh * Wd + w) * C + c, hn);
  }
}

//  DH    [2][Wd][Hd]     carry: on entry grad wrt h_step coming from step+1; on exit the element-wise part
//                        (g*z) of grad wrt h_{step-1}; the GEMM adds DGH W_hh afterwards.
//  DGISUM[2][Wd][3Hd]    accumulated over steps
//  DGH   [2][N][Wd][3Hd] per-step grad wrt (W_hh h + b_hh)
__global__ void rpe_gate_bwd_kernel(const float* __restrict__ dQPOS, const float* __restrict__ HALL,
                                    const float* __restrict__ GATES, float* __restrict__ DH,
                                    float* __restrict__ DGISUM, float* __restrict__ DGH,
                                    __nv_bfloat16* __restrict__ DGHPL, long long plane_lo, int step, int N, int Wd,
                                    int Hd, int C, int Himg) {
  const int Hd4 = Hd >> 2;
  const int n4 = 2 * Wd * Hd4;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
    const int j = (i % Hd4) << 2;
    const int r = i / Hd4;
    const int w = r % Wd, dir = r / Wd;
    const int b = dir == 0 ? step : N - 1 - step;
    const int f = dir * Hd + j;
    const int hh = f / C, c = f % C;
    const long long e = (long long)r * Hd + j;          // element index in DH
    const long long gi = ((((long long)step * 2 + dir) * Wd + w) * 4) * Hd + j;
    const long long g3 = ((long long)dir * Wd + w) * 3 * Hd + j;
    const long long d3 = ((((long long)dir * N + step) * Wd + w) * 3) * Hd + j;
    float dq[4], dh[4], rr[4], zz[4], nn[4], ghn[4], hp[4], sr[4], sz[4], sn[4];
    ld4(dQPOS + ((long long)b * Himg * Wd + (long long)hh * Wd + w) * C + c, dq);
    ld4(DH + e, dh);
    ld4(GATES + gi, rr); ld4(GATES + gi + Hd, zz); ld4(GATES + gi + 2 * Hd, nn); ld4(GATES + gi + 3 * Hd, ghn);
    ld4(HALL + (((long long)dir * (N + 1) + step) * Wd + w) * Hd + j, hp);
    ld4(DGISUM + g3, sr); ld4(DGISUM + g3 + Hd, sz); ld4(DGISUM + g3 + 2 * Hd, sn);
    float dpr[4], dpz[4], dpn[4], dhn[4], carry[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float go = dq[k] + dh[k];
      const float dn = go * (1.f - zz[k]);
      const float dz = go * (hp[k] - nn[k]);
      dpn[k] = dn * (1.f - nn[k] * nn[k]);
      dpr[k] = dpn[k] * ghn[k] * rr[k] * (1.f - rr[k]);
      dpz[k] = dz * zz[k] * (1.f - zz[k]);
      dhn[k] = dpn[k] * rr[k];
      carry[k] = go * zz[k];
      sr[k] += dpr[k]; sz[k] += dpz[k]; sn[k] += dpn[k];
    }
    st4(DGISUM + g3, sr); st4(DGISUM + g3 + Hd, sz); st4(DGISUM + g3 + 2 * Hd, sn);
    st4(DGH + d3, dpr); st4(DGH + d3 + Hd, dpz); st4(DGH + d3 + 2 * Hd, dhn);
    if (DGHPL) {
      split4_store(dpr, DGHPL, DGHPL + plane_lo, d3);
      split4_store(dpz, DGHPL, DGHPL + plane_lo, d3 + Hd);
      split4_store(dhn, DGHPL, DGHPL + plane_lo, d3 + 2 * Hd);
    }
    st4(DH + e, carry);
  }
}

}  // namespace

// warp-level tensor-core scans (gru_mma.cu); TATT_GRU_MMA=0 selects the scalar kernels above
int tatt_gru32_scan_fwd_mma_launch(const float* GI, const float* Whh, const float* bhh, float* OUT, float* GATES,
                                   int nseq, int T, int s_inner, long long outer_stride, long long inner_stride,
                                   long long t_stride, cudaStream_t st);
int tatt_gru32_scan_bwd_mma_launch(const float* dOUT, const float* GATES, const float* Whh, float* dGI, float* dGH,
                                   int nseq, int T, int s_inner, long long outer_stride, long long inner_stride,
                                   long long t_stride, cudaStream_t st);
static bool gru_mma_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("TATT_GRU_MMA");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}

extern "C" {

// rows are addressed as base(seq) + t*t_stride, base(seq) = (seq / s_inner)*outer_stride + (seq % s_inner)*inner_stride
int tatt_gru32_scan_fwd(const float* GI, const float* Whh, const float* bhh, float* OUT, float* GATES, int nseq,
                        int T, int s_inner, long long outer_stride, long long inner_stride, long long t_stride,
                        void* stream) {
  if (nseq <= 0 || T <= 0) return 0;
  TATT_REQUIRE(s_inner >= 1, "gru32_scan_fwd: s_inner must be >= 1");
  if (gru_mma_enabled())
    return tatt_gru32_scan_fwd_mma_launch(GI, Whh, bhh, OUT, GATES, nseq, T, s_inner, outer_stride, inner_stride,
                                          t_stride, (cudaStream_t)stream);
  long long warps = 2LL * ((nseq + 1) / 2);
  int blocks = (int)((warps + 3) / 4);
  gru32_scan_fwd_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(GI, Whh, bhh, OUT, GATES, nseq, T, s_inner,
                                                                  outer_stride, inner_stride, t_stride);
  TATT_LAUNCH_CHECK("gru32_scan_fwd_kernel");
  return 0;
}

int tatt_gru32_scan_bwd(const float* dOUT, const float* GATES, const float* Whh, float* dGI, float* dGH, int nseq,
                        int T, int s_inner, long long outer_stride, long long inner_stride, long long t_stride,
                        void* stream) {
  if (nseq <= 0 || T <= 0) return 0;
  TATT_REQUIRE(s_inner >= 1, "gru32_scan_bwd: s_inner must be >= 1");
  if (gru_mma_enabled())
    return tatt_gru32_scan_bwd_mma_launch(dOUT, GATES, Whh, dGI, dGH, nseq, T, s_inner, outer_stride, inner_stride,
                                          t_stride, (cudaStream_t)stream);
  long long warps = 2LL * ((nseq + 1) / 2);
  int blocks = (int)((warps + 3) / 4);
  gru32_scan_bwd_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(dOUT, GATES, Whh, dGI, dGH, nseq, T, s_inner,
                                                                  outer_stride, inner_stride, t_stride);
  TATT_LAUNCH_CHECK("gru32_scan_bwd_kernel");
  return 0;
}

int tatt_rpe_gather(const float* emb, float* X, int H, int W, int C, void* stream) {
  long long n = (long long)H * W * C;
  rpe_gather_kernel<<<(int)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(emb, X, H, W, C);
  TATT_LAUNCH_CHECK("rpe_gather_kernel");
  return 0;
}
int tatt_rpe_scatter(const float* dX, float* demb, int H, int W, int C, void* stream) {
  long long n = (long long)H * W * C;
  rpe_scatter_kernel<<<(int)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dX, demb, H, W, C);
  TATT_LAUNCH_CHECK("rpe_scatter_kernel");
  return 0;
}

int tatt_rpe_gate_fwd(const float* GI, const float* GH, float* HALL, float* GATES, float* QPOS, void* HPL,
                      long long plane_lo, int step, int N, int Wd, int Hd, int C, int Himg, void* stream) {
  TATT_REQUIRE(2 * Hd == Himg * C, "rpe_gate_fwd: 2*Hd (%d) must equal Himg*C (%d)", 2 * Hd, Himg * C);
  TATT_REQUIRE(Hd % 4 == 0 && C % 4 == 0, "rpe_gate_fwd: Hd (%d) and C (%d) must be multiples of 4", Hd, C);
  long long n = 2LL * Wd * (Hd / 4);
  rpe_gate_fwd_kernel<<<(int)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(GI, GH, HALL, GATES, QPOS,
                                                                                (__nv_bfloat16*)HPL, plane_lo, step, N,
                                                                                Wd, Hd, C, Himg);
  TATT_LAUNCH_CHECK("rpe_gate_fwd_kernel");
  return 0;
}
int tatt_rpe_gate_bwd(const float* dQPOS, const float* HALL, const float* GATES, float* DH, float* DGISUM,
                      float* DGH, void* DGHPL, long long plane_lo, int step, int N, int Wd, int Hd, int C, int Himg,
                      void* stream) {
  TATT_REQUIRE(2 * Hd == Himg * C, "rpe_gate_bwd: 2*Hd (%d) must equal Himg*C (%d)", 2 * Hd, Himg * C);
  TATT_REQUIRE(Hd % 4 == 0 && C % 4 == 0, "rpe_gate_bwd: Hd (%d) and C (%d) must be multiples of 4", Hd, C);
  long long n = 2LL * Wd * (Hd / 4);
  rpe_gate_bwd_kernel<<<(int)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(dQPOS, HALL, GATES, DH, DGISUM, DGH,
                                                                                (__nv_bfloat16*)DGHPL, plane_lo,
                                                                                step, N, Wd, Hd, C, Himg);
  TATT_LAUNCH_CHECK("rpe_gate_bwd_kernel");
  return 0;
}

int tatt_memcpy_d2d(void* dst, const void* src, long long bytes, void* stream) {
  TATT_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}
int tatt_memset0(void* dst, long long bytes, void* stream) {
  TATT_CUDA(cudaMemsetAsync(dst, 0, (size_t)bytes, (cudaStream_t)stream));
  return 0;
}

}  // extern "C"
