// GRU recurrences.
//  (1) SRB BiGRU, hidden 32 (GruBlock, model/tsrn.py:1067-1084): one warp per (sequence, direction),
//      h in registers, W_hh rows in registers, strided row addressing so that vertical (T=H) and
//      horizontal (T=W) scans run directly on the NHWC feature map -- the reference's
//      permute/contiguous copies (tsrn.py:1076-1083, 906) do not exist here.
//  (2) RPE BiGRU of the TP Interpreter (model/transformer_v2.py:177, 215-221; quirk Q1: recurrence
//      over the BATCH axis): per-step gate kernels around the batched recurrent GEMM (gemm.cu).
// Gate order r,z,n; n = tanh(gi_n + r * (W_hn h + b_hn)); h' = (1-z) n + z h   (torch.nn.GRU).
#include <cuda_bf16.h>
#include "common.cuh"

namespace {

__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16* hi, __nv_bfloat16* lo, long long i) {
  __nv_bfloat16 h = __float2bfloat16_rn(x);
  hi[i] = h;
  lo[i] = __float2bfloat16_rn(x - __bfloat162float(h));
}

// fast activations for the recurrences: MUFU ex2/rcp based, abs error ~1e-6 (tolerance of the path: 1e-3)
__device__ __forceinline__ float fsigmoid(float x) { return __frcp_rn(1.f + __expf(-x)); }
__device__ __forceinline__ float ftanh(float x) { return 1.f - 2.f * __frcp_rn(1.f + __expf(2.f * x)); }

// GI    [rows][192]  = [dir][gate][32]   (x W_ih^T + b_ih, both directions)
// OUT   [rows][64]   = [dir][32]
// GATES [rows][320]  = [dir][r,z,n,ghn,hprev][32]   (saved for the backward; may be NULL)
__global__ void __launch_bounds__(128)
gru32_scan_fwd_kernel(const float* __restrict__ GI, const float* __restrict__ Whh,
                      const float* __restrict__ bhh, float* __restrict__ OUT, float* __restrict__ GATES,
                      int nseq, int T, int s_inner, long long outer_stride, long long inner_stride,
                      long long t_stride) {
  __shared__ __align__(16) float hs[4][32];     // per-warp broadcast buffer for the hidden state
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int dir = (int)(gw & 1);
  const long long seq = gw >> 1;
  if (seq >= nseq) return;
  const long long base = (seq / s_inner) * outer_stride + (seq % s_inner) * inner_stride;

  float wr[32], wz[32], wn[32];
  const float* W = Whh + dir * 96 * 32;
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    wr[k] = W[(lane)*32 + k];
    wz[k] = W[(32 + lane) * 32 + k];
    wn[k] = W[(64 + lane) * 32 + k];
  }
  const float br = bhh[dir * 96 + lane], bz = bhh[dir * 96 + 32 + lane], bn = bhh[dir * 96 + 64 + lane];

  float h = 0.f;
  long long row = base + (dir == 0 ? 0 : (long long)(T - 1)) * t_stride;
  const long long step_stride = dir == 0 ? t_stride : -t_stride;
  // The recurrence is a dependent chain of ~500 cycles per step while the GI row of a step comes from
  // L2/HBM (~1000+ cycles): keep a ring of PF steps of input projections in flight.
  constexpr int PF = 6;
  float pr[PF], pz[PF], pn[PF];
#pragma unroll
  for (int j = 0; j < PF; ++j) {
    pr[j] = pz[j] = pn[j] = 0.f;
    if (j < T) {
      const float* gp = GI + (row + j * step_stride) * 192 + dir * 96 + lane;
      pr[j] = gp[0];
      pz[j] = gp[32];
      pn[j] = gp[64];
    }
  }
  for (int s0 = 0; s0 < T; s0 += PF) {
#pragma unroll
    for (int j = 0; j < PF; ++j) {
      const int s = s0 + j;
      if (s < T) {
        const float gr = pr[j], gz = pz[j], gn = pn[j];
        if (s + PF < T) {
          const float* np = GI + (row + PF * step_stride) * 192 + dir * 96 + lane;
          pr[j] = np[0];
          pz[j] = np[32];
          pn[j] = np[64];
        }
        float ar = br, az = bz, an = bn;
        float ar2 = 0.f, az2 = 0.f, an2 = 0.f;          // two partial chains halve the FMA dependency depth
        // broadcast h through shared memory: 1 STS + 8 LDS.128 instead of 32 shuffles
        __syncwarp();
        hs[wib][lane] = h;
        __syncwarp();
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
          const float4 hv = *reinterpret_cast<const float4*>(&hs[wib][k4 * 4]);
          ar = fmaf(wr[k4 * 4 + 0], hv.x, ar);
          az = fmaf(wz[k4 * 4 + 0], hv.x, az);
          an = fmaf(wn[k4 * 4 + 0], hv.x, an);
          ar2 = fmaf(wr[k4 * 4 + 1], hv.y, ar2);
          az2 = fmaf(wz[k4 * 4 + 1], hv.y, az2);
          an2 = fmaf(wn[k4 * 4 + 1], hv.y, an2);
          ar = fmaf(wr[k4 * 4 + 2], hv.z, ar);
          az = fmaf(wz[k4 * 4 + 2], hv.z, az);
          an = fmaf(wn[k4 * 4 + 2], hv.z, an);
          ar2 = fmaf(wr[k4 * 4 + 3], hv.w, ar2);
          az2 = fmaf(wz[k4 * 4 + 3], hv.w, az2);
          an2 = fmaf(wn[k4 * 4 + 3], hv.w, an2);
        }
        ar += ar2;
        az += az2;
        an += an2;
        float r = fsigmoid(gr + ar);
        float z = fsigmoid(gz + az);
        float n = ftanh(gn + r * an);
        float hn = n + z * (h - n);
        OUT[row * 64 + dir * 32 + lane] = hn;
        if (GATES) {
          float* g = GATES + row * 320 + dir * 160 + lane;
          g[0] = r;
          g[32] = z;
          g[64] = n;
          g[96] = an;
          g[128] = h;
        }
        h = hn;
        row += step_stride;
      }
    }
  }
}

// dGI [rows][192] (grad wrt the input projections), dGH [rows][192] (grad wrt W_hh h + b_hh)
__global__ void __launch_bounds__(128)
gru32_scan_bwd_kernel(const float* __restrict__ dOUT, const float* __restrict__ GATES,
                      const float* __restrict__ Whh, float* __restrict__ dGI, float* __restrict__ dGH, int nseq,
                      int T, int s_inner, long long outer_stride, long long inner_stride, long long t_stride) {
  __shared__ __align__(16) float gs[4][96];     // per-warp broadcast buffer: dpr | dpz | dhn
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int dir = (int)(gw & 1);
  const long long seq = gw >> 1;
  if (seq >= nseq) return;
  const long long base = (seq / s_inner) * outer_stride + (seq % s_inner) * inner_stride;

  float wc[96];  // column `lane` of W_hh[dir]
  const float* W = Whh + dir * 96 * 32;
#pragma unroll
  for (int i = 0; i < 96; ++i) wc[i] = W[i * 32 + lane];

  float dh = 0.f;
  // walk the recurrence backwards: last processed step first
  long long row = base + (dir == 0 ? (long long)(T - 1) : 0) * t_stride;
  const long long step_stride = dir == 0 ? -t_stride : t_stride;
  constexpr int PF = 4;
  float qr[PF], qz[PF], qn[PF], qg[PF], qh[PF], qo[PF];
#pragma unroll
  for (int j = 0; j < PF; ++j) {
    qr[j] = qz[j] = qn[j] = qg[j] = qh[j] = qo[j] = 0.f;
    if (j < T) {
      const long long rj = row + j * step_stride;
      const float* g = GATES + rj * 320 + dir * 160 + lane;
      qr[j] = g[0]; qz[j] = g[32]; qn[j] = g[64]; qg[j] = g[96]; qh[j] = g[128];
      qo[j] = dOUT[rj * 64 + dir * 32 + lane];
    }
  }
  for (int s0 = 0; s0 < T; s0 += PF) {
#pragma unroll
    for (int j = 0; j < PF; ++j) {
      const int s = s0 + j;
      if (s < T) {
        const float r = qr[j], z = qz[j], n = qn[j], ghn = qg[j], hp = qh[j];
        const float go = qo[j] + dh;
        if (s + PF < T) {
          const long long rj = row + PF * step_stride;
          const float* g = GATES + rj * 320 + dir * 160 + lane;
          qr[j] = g[0]; qz[j] = g[32]; qn[j] = g[64]; qg[j] = g[96]; qh[j] = g[128];
          qo[j] = dOUT[rj * 64 + dir * 32 + lane];
        }
        float dn = go * (1.f - z);
        float dz = go * (hp - n);
        float dpn = dn * (1.f - n * n);
        float dpr = dpn * ghn * r * (1.f - r);
        float dpz = dz * z * (1.f - z);
        float dhn = dpn * r;
        float* o = dGI + row * 192 + dir * 96 + lane;
        o[0] = dpr;
        o[32] = dpz;
        o[64] = dpn;
        float* o2 = dGH + row * 192 + dir * 96 + lane;
        o2[0] = dpr;
        o2[32] = dpz;
        o2[64] = dhn;
        float acc = go * z, acc2 = 0.f, acc3 = 0.f;
        __syncwarp();
        gs[wib][lane] = dpr;
        gs[wib][32 + lane] = dpz;
        gs[wib][64 + lane] = dhn;
        __syncwarp();
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
          const float4 a = *reinterpret_cast<const float4*>(&gs[wib][k4 * 4]);
          const float4 b = *reinterpret_cast<const float4*>(&gs[wib][32 + k4 * 4]);
          const float4 c = *reinterpret_cast<const float4*>(&gs[wib][64 + k4 * 4]);
          acc = fmaf(wc[k4 * 4 + 0], a.x, acc);
          acc2 = fmaf(wc[32 + k4 * 4 + 0], b.x, acc2);
          acc3 = fmaf(wc[64 + k4 * 4 + 0], c.x, acc3);
          acc = fmaf(wc[k4 * 4 + 1], a.y, acc);
          acc2 = fmaf(wc[32 + k4 * 4 + 1], b.y, acc2);
          acc3 = fmaf(wc[64 + k4 * 4 + 1], c.y, acc3);
          acc = fmaf(wc[k4 * 4 + 2], a.z, acc);
          acc2 = fmaf(wc[32 + k4 * 4 + 2], b.z, acc2);
          acc3 = fmaf(wc[64 + k4 * 4 + 2], c.z, acc3);
          acc = fmaf(wc[k4 * 4 + 3], a.w, acc);
          acc2 = fmaf(wc[32 + k4 * 4 + 3], b.w, acc2);
          acc3 = fmaf(wc[64 + k4 * 4 + 3], c.w, acc3);
        }
        dh = acc + acc2 + acc3;
        row += step_stride;
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
__global__ void rpe_gate_fwd_kernel(const float* __restrict__ GI, const float* __restrict__ GH,
                                    float* __restrict__ HALL, float* __restrict__ GATES,
                                    float* __restrict__ QPOS, __nv_bfloat16* __restrict__ HPL, long long plane_lo,
                                    int step, int N, int Wd, int Hd, int C, int Himg) {
  long long n = 2LL * Wd * Hd;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    int j = (int)(i % Hd);
    long long r = i / Hd;
    int w = (int)(r % Wd);
    int dir = (int)(r / Wd);
    long long g3 = ((long long)dir * Wd + w) * 3 * Hd + j;
    float gr = GI[g3] + GH[g3];
    float gz = GI[g3 + Hd] + GH[g3 + Hd];
    float ghn = GH[g3 + 2 * Hd];
    float rr = sigmoid_f(gr), zz = sigmoid_f(gz);
    float nn = tanhf(GI[g3 + 2 * Hd] + rr * ghn);
    long long hidx = (((long long)dir * (N + 1) + step) * Wd + w) * Hd + j;
    float hp = HALL[hidx];
    float hn = nn + zz * (hp - nn);
    HALL[hidx + (long long)Wd * Hd] = hn;
    if (HPL) split_bf16(hn, HPL, HPL + plane_lo, hidx + (long long)Wd * Hd);
    if (GATES) {
      long long gi = ((((long long)step * 2 + dir) * Wd + w) * 4) * Hd + j;
      GATES[gi] = rr;
      GATES[gi + Hd] = zz;
      GATES[gi + 2 * Hd] = nn;
      GATES[gi + 3 * Hd] = ghn;
    }
    int b = dir == 0 ? step : N - 1 - step;
    int f = dir * Hd + j;
    int hh = f / C, c = f % C;
    QPOS[((long long)b * Himg * Wd + (long long)hh * Wd + w) * C + c] = hn;
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
  long long n = 2LL * Wd * Hd;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    int j = (int)(i % Hd);
    long long r = i / Hd;
    int w = (int)(r % Wd);
    int dir = (int)(r / Wd);
    int b = dir == 0 ? step : N - 1 - step;
    int f = dir * Hd + j;
    int hh = f / C, c = f % C;
    float go = dQPOS[((long long)b * Himg * Wd + (long long)hh * Wd + w) * C + c] + DH[i];
    long long gi = ((((long long)step * 2 + dir) * Wd + w) * 4) * Hd + j;
    float rr = GATES[gi], zz = GATES[gi + Hd], nn = GATES[gi + 2 * Hd], ghn = GATES[gi + 3 * Hd];
    float hp = HALL[(((long long)dir * (N + 1) + step) * Wd + w) * Hd + j];
    float dn = go * (1.f - zz);
    float dz = go * (hp - nn);
    float dpn = dn * (1.f - nn * nn);
    float dpr = dpn * ghn * rr * (1.f - rr);
    float dpz = dz * zz * (1.f - zz);
    long long g3 = ((long long)dir * Wd + w) * 3 * Hd + j;
    DGISUM[g3] += dpr;
    DGISUM[g3 + Hd] += dpz;
    DGISUM[g3 + 2 * Hd] += dpn;
    long long d3 = ((((long long)dir * N + step) * Wd + w) * 3) * Hd + j;
    DGH[d3] = dpr;
    DGH[d3 + Hd] = dpz;
    DGH[d3 + 2 * Hd] = dpn * rr;
    if (DGHPL) {
      split_bf16(dpr, DGHPL, DGHPL + plane_lo, d3);
      split_bf16(dpz, DGHPL, DGHPL + plane_lo, d3 + Hd);
      split_bf16(dpn * rr, DGHPL, DGHPL + plane_lo, d3 + 2 * Hd);
    }
    DH[i] = go * zz;
  }
}

}  // namespace

extern "C" {

// rows are addressed as base(seq) + t*t_stride, base(seq) = (seq / s_inner)*outer_stride + (seq % s_inner)*inner_stride
int tatt_gru32_scan_fwd(const float* GI, const float* Whh, const float* bhh, float* OUT, float* GATES, int nseq,
                        int T, int s_inner, long long outer_stride, long long inner_stride, long long t_stride,
                        void* stream) {
  if (nseq <= 0 || T <= 0) return 0;
  TATT_REQUIRE(s_inner >= 1, "gru32_scan_fwd: s_inner must be >= 1");
  long long warps = 2LL * nseq;
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
  long long warps = 2LL * nseq;
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
  long long n = 2LL * Wd * Hd;
  rpe_gate_fwd_kernel<<<(int)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(GI, GH, HALL, GATES, QPOS,
                                                                                (__nv_bfloat16*)HPL, plane_lo, step, N,
                                                                                Wd, Hd, C, Himg);
  TATT_LAUNCH_CHECK("rpe_gate_fwd_kernel");
  return 0;
}
int tatt_rpe_gate_bwd(const float* dQPOS, const float* HALL, const float* GATES, float* DH, float* DGISUM,
                      float* DGH, void* DGHPL, long long plane_lo, int step, int N, int Wd, int Hd, int C, int Himg,
                      void* stream) {
  TATT_REQUIRE(2 * Hd == Himg * C, "rpe_gate_bwd: 2*Hd (%d) must equal Himg*C (%d)", 2 * Hd, Himg * C);
  long long n = 2LL * Wd * Hd;
  rpe_gate_bwd_kernel<<<(int)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dQPOS, HALL, GATES, DH, DGISUM, DGH,
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
