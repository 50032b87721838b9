// Kernels of the CRNN text-prior generator that feeds the hot path: /root/reference/model/crnn/crnn.py:5-93
// (7 convs + 2 BidirectionalLSTM(256)), interfaces/base.py:797-815 (parse_crnn_data: bicubic resize to 32 x 100,
// RGB -> gray), interfaces/super_resolution.py:794-799 (softmax / permute to the [N, 37, 1, 26] text prior).
// The convolutions, BatchNorm(+ReLU), linear layers and the per-step recurrent GEMMs run on the engines the SR path
// already has (gemm.cu / tc2_gemm.cu / norm.cu); this file adds what only the CRNN needs:
//   * bicubic (A = -0.75, align_corners=False, like torch's upsample_bicubic2d) resize + gray conversion
//   * general max pooling (kernel / stride / padding; the CRNN's (2,2),(2,1),(0,1) windows overlap), NHWC
//   * crop / zero-pad of an NHWC map (conv6 is a 2x2 "valid" convolution, run as a same-size conv + crop)
//   * the LSTM cell: gate math forward / backward of one time step of both directions
//   * [A][B][C] -> [B][A][C] permute, softmax over the classes (+ its backward) with the permuted prior as a side output
#include "common.cuh"

namespace {

#define GRID_STRIDE(i, n) \
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (n); i += (long long)gridDim.x * blockDim.x)

inline int ew_grid(long long n) {
  long long b = (n + 255) / 256;
  return (int)(b < 1 ? 1 : (b > 148 * 16 ? 148 * 16 : b));
}

// ------------------------------------------------------------------------------------------------ bicubic + gray
__device__ __forceinline__ void cubic_coeffs(float t, float (&w)[4]) {
  const float A = -0.75f;
  const float x0 = t + 1.f, x1 = t, x2 = 1.f - t, x3 = 2.f - t;
  w[0] = ((A * x0 - 5.f * A) * x0 + 8.f * A) * x0 - 4.f * A;
  w[1] = ((A + 2.f) * x1 - (A + 3.f)) * x1 * x1 + 1.f;
  w[2] = ((A + 2.f) * x2 - (A + 3.f)) * x2 * x2 + 1.f;
  w[3] = ((A * x3 - 5.f * A) * x3 + 8.f * A) * x3 - 4.f * A;
}

// img [N][C>=3][H][W] (NCHW) -> out [N][OH][OW] (= NCHW [N,1,OH,OW]): 0.299 R' + 0.587 G' + 0.114 B' of the resized planes
__global__ void __launch_bounds__(256)
bicubic_gray_kernel(const float* __restrict__ img, float* __restrict__ out, int N, int C, int H, int W, int OH, int OW) {
  const float sh = (float)H / (float)OH, sw = (float)W / (float)OW;
  GRID_STRIDE(i, (long long)N * OH * OW) {
    const int ox = (int)(i % OW);
    const long long q = i / OW;
    const int oy = (int)(q % OH);
    const int n = (int)(q / OH);
    const float ry = sh * (oy + 0.5f) - 0.5f, rx = sw * (ox + 0.5f) - 0.5f;
    const float fy = floorf(ry), fx = floorf(rx);
    const int iy = (int)fy, ix = (int)fx;
    float wy[4], wx[4];
    cubic_coeffs(ry - fy, wy);
    cubic_coeffs(rx - fx, wx);
    float ch[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* p = img + ((long long)n * C + c) * H * W;
      float acc = 0.f;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int yy = min(max(iy - 1 + a, 0), H - 1);
        float row = 0.f;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int xx = min(max(ix - 1 + b, 0), W - 1);
          row = fmaf(p[yy * W + xx], wx[b], row);
        }
        acc = fmaf(row, wy[a], acc);
      }
      ch[c] = acc;
    }
    out[i] = 0.299f * ch[0] + 0.587f * ch[1] + 0.114f * ch[2];
  }
}

// ------------------------------------------------------------------------------------------------ max pooling (NHWC)
struct PoolP {
  int H, W, C, OH, OW, kh, kw, sh, sw, ph, pw;
};

__global__ void __launch_bounds__(256)
maxpool2d_fwd_kernel(const float* __restrict__ in, float* __restrict__ out, long long N, const PoolP p) {
  GRID_STRIDE(i, N * p.OH * p.OW * p.C) {
    const int c = (int)(i % p.C);
    long long q = i / p.C;
    const int ox = (int)(q % p.OW);
    q /= p.OW;
    const int oy = (int)(q % p.OH);
    const long long n = q / p.OH;
    float m = -INFINITY;
    for (int a = 0; a < p.kh; ++a) {
      const int y = oy * p.sh - p.ph + a;
      if (y < 0 || y >= p.H) continue;
      for (int b = 0; b < p.kw; ++b) {
        const int x = ox * p.sw - p.pw + b;
        if (x < 0 || x >= p.W) continue;
        const float v = in[((n * p.H + y) * p.W + x) * p.C + c];
        m = v > m ? v : m;
      }
    }
    out[i] = m;
  }
}

// gather form (no atomics): every input element sums the gradients of the windows whose arg-max it is; ties go to the
// first maximum in row-major window order, like torch's max_pool2d
__global__ void __launch_bounds__(256)
maxpool2d_bwd_kernel(const float* __restrict__ in, const float* __restrict__ dout, float* __restrict__ din, long long N,
                     const PoolP p) {
  GRID_STRIDE(i, N * p.H * p.W * p.C) {
    const int c = (int)(i % p.C);
    long long q = i / p.C;
    const int x = (int)(q % p.W);
    q /= p.W;
    const int y = (int)(q % p.H);
    const long long n = q / p.H;
    float g = 0.f;
    // windows (oy, ox) with oy*sh - ph <= y < oy*sh - ph + kh
    const int oy_hi = min((y + p.ph) / p.sh, p.OH - 1), ox_hi = min((x + p.pw) / p.sw, p.OW - 1);
    for (int oy = oy_hi; oy >= 0 && oy * p.sh - p.ph + p.kh > y; --oy)
      for (int ox = ox_hi; ox >= 0 && ox * p.sw - p.pw + p.kw > x; --ox) {
        float m = -INFINITY;
        int ay = -1, ax = -1;
        for (int a = 0; a < p.kh; ++a) {
          const int yy = oy * p.sh - p.ph + a;
          if (yy < 0 || yy >= p.H) continue;
          for (int b = 0; b < p.kw; ++b) {
            const int xx = ox * p.sw - p.pw + b;
            if (xx < 0 || xx >= p.W) continue;
            const float v = in[((n * p.H + yy) * p.W + xx) * p.C + c];
            if (v > m || ay < 0) {
              m = v;
              ay = yy;
              ax = xx;
            }
          }
        }
        if (ay == y && ax == x) g += dout[((n * p.OH + oy) * p.OW + ox) * p.C + c];
      }
    din[i] = g;
  }
}

// ------------------------------------------------------------------------------------------------ crop / zero-pad
// out[n][y < OH][x < OW][c] = in[n][y][x][c]           (fwd: crop of the top-left corner)
__global__ void __launch_bounds__(256)
crop_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out, long long N, int H, int W, int OH, int OW, int C) {
  GRID_STRIDE(i, N * OH * OW * C) {
    const int c = (int)(i % C);
    long long q = i / C;
    const int x = (int)(q % OW);
    q /= OW;
    const int y = (int)(q % OH);
    const long long n = q / OH;
    out[i] = in[((n * H + y) * W + x) * C + c];
  }
}
// out[n][y][x][c] = (y < IH && x < IW) ? in[n][y][x][c] : 0      (bwd: zero-pad back to H x W)
__global__ void __launch_bounds__(256)
pad_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out, long long N, int IH, int IW, int H, int W, int C) {
  GRID_STRIDE(i, N * H * W * C) {
    const int c = (int)(i % C);
    long long q = i / C;
    const int x = (int)(q % W);
    q /= W;
    const int y = (int)(q % H);
    const long long n = q / H;
    out[i] = (y < IH && x < IW) ? in[((n * IH + y) * IW + x) * C + c] : 0.f;
  }
}

// [A][B][C] -> [B][A][C]
__global__ void __launch_bounds__(256)
permute_102_kernel(const float* __restrict__ in, float* __restrict__ out, int A, int B, int C) {
  GRID_STRIDE(i, (long long)A * B * C) {
    const int c = (int)(i % C);
    const long long q = i / C;
    const int a = (int)(q % A);
    const int b = (int)(q / A);
    out[i] = in[((long long)a * B + b) * C + c];
  }
}

// ------------------------------------------------------------------------------------------------ LSTM cell
// Layouts (T time steps, Nb sequences, H hidden units, both directions d in {0, 1}):
//   G   [T*Nb][8H]  row t*Nb+n: [d][i f g o][H]; on entry W_ih x + b_ih, on exit the ACTIVATED gates (saved for backward)
//   GH  [2][Nb][4H] W_hh h_prev + b_hh of this step, or NULL at step 0 (h_prev = 0: b_hh is added from BHH)
//   CS  [2][T][Nb][H] cell state after processing time t
//   OUT [T*Nb][2H]  hidden state after processing time t: [h_fwd | h_bwd]  (= the nn.LSTM output)
// step s processes time t = s (d = 0) and t = T-1-s (d = 1).
__global__ void __launch_bounds__(256)
lstm_gate_fwd_kernel(float* __restrict__ G, const float* __restrict__ GH, const float* __restrict__ BHH,
                     float* __restrict__ CS, float* __restrict__ OUT, int s, int T, int Nb, int H) {
  GRID_STRIDE(idx, 2LL * Nb * H) {
    const int j = (int)(idx % H);
    const long long q = idx / H;
    const int n = (int)(q % Nb);
    const int d = (int)(q / Nb);
    const int t = d == 0 ? s : T - 1 - s;
    const int tp = d == 0 ? t - 1 : t + 1;
    float* g = G + ((long long)t * Nb + n) * 8 * H + (long long)d * 4 * H + j;
    float a[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      a[k] = g[k * H] + (GH ? GH[((long long)d * Nb + n) * 4 * H + k * H + j] : BHH[(long long)d * 4 * H + k * H + j]);
    const float ig = sigmoid_f(a[0]), fg = sigmoid_f(a[1]), gg = tanhf(a[2]), og = sigmoid_f(a[3]);
    const float cp = s == 0 ? 0.f : CS[(((long long)d * T + tp) * Nb + n) * H + j];
    const float c = fg * cp + ig * gg;
    g[0] = ig; g[H] = fg; g[2 * H] = gg; g[3 * H] = og;
    CS[(((long long)d * T + t) * Nb + n) * H + j] = c;
    OUT[((long long)t * Nb + n) * 2 * H + (long long)d * H + j] = og * tanhf(c);
  }
}

// dG [T*Nb][8H] receives the PRE-activation gate gradients of time t; DH [2][Nb][H] = dgates W_hh of the step processed
// before this one in backward order (ignored at s == T-1); DC [2][Nb][H] running cell-state gradient (in/out).
__global__ void __launch_bounds__(256)
lstm_gate_bwd_kernel(const float* __restrict__ G, const float* __restrict__ CS, const float* __restrict__ dOUT,
                     const float* __restrict__ DH, float* __restrict__ DC, float* __restrict__ dG, int s, int T, int Nb,
                     int H) {
  GRID_STRIDE(idx, 2LL * Nb * H) {
    const int j = (int)(idx % H);
    const long long q = idx / H;
    const int n = (int)(q % Nb);
    const int d = (int)(q / Nb);
    const int t = d == 0 ? s : T - 1 - s;
    const int tp = d == 0 ? t - 1 : t + 1;
    const long long go = ((long long)t * Nb + n) * 8 * H + (long long)d * 4 * H + j;
    const float ig = G[go], fg = G[go + H], gg = G[go + 2 * H], og = G[go + 3 * H];
    const float c = CS[(((long long)d * T + t) * Nb + n) * H + j];
    const float cp = s == 0 ? 0.f : CS[(((long long)d * T + tp) * Nb + n) * H + j];
    const long long so = ((long long)d * Nb + n) * H + j;
    float dh = dOUT[((long long)t * Nb + n) * 2 * H + (long long)d * H + j];
    float dc = 0.f;
    if (s != T - 1) {
      dh += DH[so];
      dc = DC[so];
    }
    const float tc = tanhf(c);
    const float dct = dc + dh * og * (1.f - tc * tc);
    dG[go] = dct * gg * ig * (1.f - ig);
    dG[go + H] = dct * cp * fg * (1.f - fg);
    dG[go + 2 * H] = dct * ig * (1.f - gg * gg);
    dG[go + 3 * H] = dh * tc * og * (1.f - og);
    DC[so] = dct * fg;
  }
}

// ------------------------------------------------------------------------------------------------ class softmax
// logits [R][C] (R = T*Nb rows, C <= 64 classes): probs [R][C]; prior [Nb][C][T] (= [N, C, 1, T]) optional.  One warp per row.
__global__ void __launch_bounds__(256)
softmax_prior_fwd_kernel(const float* __restrict__ logits, float* __restrict__ probs, float* __restrict__ prior, int T,
                         int Nb, int C) {
  const int lane = threadIdx.x & 31;
  const long long R = (long long)T * Nb;
  for (long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); r < R; r += (long long)gridDim.x * 8) {
    const float* x = logits + r * C;
    const float v0 = lane < C ? x[lane] : -INFINITY, v1 = lane + 32 < C ? x[lane + 32] : -INFINITY;
    const float m = warp_max(fmaxf(v0, v1));
    const float e0 = lane < C ? expf(v0 - m) : 0.f, e1 = lane + 32 < C ? expf(v1 - m) : 0.f;
    const float inv = 1.f / warp_sum(e0 + e1);
    const int t = (int)(r / Nb), n = (int)(r % Nb);
    if (lane < C) {
      probs[r * C + lane] = e0 * inv;
      if (prior) prior[((long long)n * C + lane) * T + t] = e0 * inv;
    }
    if (lane + 32 < C) {
      probs[r * C + lane + 32] = e1 * inv;
      if (prior) prior[((long long)n * C + lane + 32) * T + t] = e1 * inv;
    }
  }
}
// dlogits = p * (dp - sum_c dp_c p_c)
__global__ void __launch_bounds__(256)
softmax_bwd_kernel(const float* __restrict__ probs, const float* __restrict__ dprobs, float* __restrict__ dlogits,
                   long long R, int C) {
  const int lane = threadIdx.x & 31;
  for (long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); r < R; r += (long long)gridDim.x * 8) {
    const float p0 = lane < C ? probs[r * C + lane] : 0.f, p1 = lane + 32 < C ? probs[r * C + lane + 32] : 0.f;
    const float g0 = lane < C ? dprobs[r * C + lane] : 0.f, g1 = lane + 32 < C ? dprobs[r * C + lane + 32] : 0.f;
    const float dot = warp_sum(p0 * g0 + p1 * g1);
    if (lane < C) dlogits[r * C + lane] = p0 * (g0 - dot);
    if (lane + 32 < C) dlogits[r * C + lane + 32] = p1 * (g1 - dot);
  }
}

}  // namespace

extern "C" {

int tatt_bicubic_gray(const float* img, float* out, int N, int C, int H, int W, int OH, int OW, void* stream) {
  TATT_REQUIRE(N >= 1 && C >= 3 && H >= 1 && W >= 1 && OH >= 1 && OW >= 1, "bicubic_gray: bad shape [%d,%d,%d,%d] -> %dx%d", N,
               C, H, W, OH, OW);
  bicubic_gray_kernel<<<ew_grid((long long)N * OH * OW), 256, 0, (cudaStream_t)stream>>>(img, out, N, C, H, W, OH, OW);
  TATT_LAUNCH_CHECK("bicubic_gray_kernel");
  return 0;
}

static int pool_params(PoolP& p, int H, int W, int C, int kh, int kw, int sh, int sw, int ph, int pw) {
  TATT_REQUIRE(kh >= 1 && kw >= 1 && sh >= 1 && sw >= 1 && ph >= 0 && pw >= 0 && 2 * ph <= kh && 2 * pw <= kw,
               "maxpool2d: bad window k=(%d,%d) s=(%d,%d) p=(%d,%d)", kh, kw, sh, sw, ph, pw);
  p.H = H; p.W = W; p.C = C; p.kh = kh; p.kw = kw; p.sh = sh; p.sw = sw; p.ph = ph; p.pw = pw;
  p.OH = (H + 2 * ph - kh) / sh + 1;          // floor mode (nn.MaxPool2d default)
  p.OW = (W + 2 * pw - kw) / sw + 1;
  TATT_REQUIRE(p.OH >= 1 && p.OW >= 1, "maxpool2d: window larger than the padded input");
  return 0;
}

/* out [N][OH][OW][C], OH = (H + 2 ph - kh) / sh + 1 (floor), same for OW */
int tatt_maxpool2d_fwd(const float* in, float* out, long long N, int H, int W, int C, int kh, int kw, int sh, int sw, int ph,
                       int pw, void* stream) {
  PoolP p;
  if (int rc = pool_params(p, H, W, C, kh, kw, sh, sw, ph, pw)) return rc;
  if (N <= 0) return 0;
  maxpool2d_fwd_kernel<<<ew_grid(N * p.OH * p.OW * C), 256, 0, (cudaStream_t)stream>>>(in, out, N, p);
  TATT_LAUNCH_CHECK("maxpool2d_fwd_kernel");
  return 0;
}

int tatt_maxpool2d_bwd(const float* in, const float* dout, float* din, long long N, int H, int W, int C, int kh, int kw,
                       int sh, int sw, int ph, int pw, void* stream) {
  PoolP p;
  if (int rc = pool_params(p, H, W, C, kh, kw, sh, sw, ph, pw)) return rc;
  if (N <= 0) return 0;
  maxpool2d_bwd_kernel<<<ew_grid(N * H * W * C), 256, 0, (cudaStream_t)stream>>>(in, dout, din, N, p);
  TATT_LAUNCH_CHECK("maxpool2d_bwd_kernel");
  return 0;
}

/* pad == 0: out [N][OH][OW][C] = top-left crop of in [N][H][W][C]; pad == 1: out [N][H][W][C] = in [N][OH][OW][C] zero-padded */
int tatt_crop_nhwc(const float* in, float* out, long long N, int H, int W, int OH, int OW, int C, int pad, void* stream) {
  TATT_REQUIRE(OH >= 1 && OW >= 1 && OH <= H && OW <= W && C >= 1, "crop_nhwc: bad sizes %dx%d -> %dx%d", H, W, OH, OW);
  if (N <= 0) return 0;
  if (pad)
    pad_nhwc_kernel<<<ew_grid(N * H * W * C), 256, 0, (cudaStream_t)stream>>>(in, out, N, OH, OW, H, W, C);
  else
    crop_nhwc_kernel<<<ew_grid(N * OH * OW * C), 256, 0, (cudaStream_t)stream>>>(in, out, N, H, W, OH, OW, C);
  TATT_LAUNCH_CHECK("crop_nhwc_kernel");
  return 0;
}

int tatt_permute_102(const float* in, float* out, int A, int B, int C, void* stream) {
  TATT_REQUIRE(A >= 1 && B >= 1 && C >= 1, "permute_102: bad shape");
  permute_102_kernel<<<ew_grid((long long)A * B * C), 256, 0, (cudaStream_t)stream>>>(in, out, A, B, C);
  TATT_LAUNCH_CHECK("permute_102_kernel");
  return 0;
}

int tatt_lstm_gate_fwd(float* G, const float* GH, const float* BHH, float* CS, float* OUT, int s, int T, int Nb, int H,
                       void* stream) {
  TATT_REQUIRE(T >= 1 && s >= 0 && s < T && Nb >= 1 && H >= 1, "lstm_gate_fwd: bad step %d of %d", s, T);
  TATT_REQUIRE(GH != nullptr || s == 0, "lstm_gate_fwd: GH may only be NULL at step 0");
  lstm_gate_fwd_kernel<<<ew_grid(2LL * Nb * H), 256, 0, (cudaStream_t)stream>>>(G, GH, BHH, CS, OUT, s, T, Nb, H);
  TATT_LAUNCH_CHECK("lstm_gate_fwd_kernel");
  return 0;
}

int tatt_lstm_gate_bwd(const float* G, const float* CS, const float* dOUT, const float* DH, float* DC, float* dG, int s,
                       int T, int Nb, int H, void* stream) {
  TATT_REQUIRE(T >= 1 && s >= 0 && s < T && Nb >= 1 && H >= 1, "lstm_gate_bwd: bad step %d of %d", s, T);
  lstm_gate_bwd_kernel<<<ew_grid(2LL * Nb * H), 256, 0, (cudaStream_t)stream>>>(G, CS, dOUT, DH, DC, dG, s, T, Nb, H);
  TATT_LAUNCH_CHECK("lstm_gate_bwd_kernel");
  return 0;
}

int tatt_softmax_prior_fwd(const float* logits, float* probs, float* prior, int T, int Nb, int C, void* stream) {
  TATT_REQUIRE(T >= 1 && Nb >= 1 && C >= 1 && C <= 64, "softmax_prior_fwd: needs 1 <= classes <= 64 (got %d)", C);
  const long long R = (long long)T * Nb;
  softmax_prior_fwd_kernel<<<(int)((R + 7) / 8 < 1184 ? (R + 7) / 8 : 1184), 256, 0, (cudaStream_t)stream>>>(logits, probs,
                                                                                                           prior, T, Nb, C);
  TATT_LAUNCH_CHECK("softmax_prior_fwd_kernel");
  return 0;
}

int tatt_softmax_bwd(const float* probs, const float* dprobs, float* dlogits, long long R, int C, void* stream) {
  TATT_REQUIRE(R >= 1 && C >= 1 && C <= 64, "softmax_bwd: needs 1 <= classes <= 64 (got %d)", C);
  softmax_bwd_kernel<<<(int)((R + 7) / 8 < 1184 ? (R + 7) / 8 : 1184), 256, 0, (cudaStream_t)stream>>>(probs, dprobs, dlogits,
                                                                                                     R, C);
  TATT_LAUNCH_CHECK("softmax_bwd_kernel");
  return 0;
}

}  // extern "C"
