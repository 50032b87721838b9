// Generic fp32 tiled GEMM / implicit-GEMM convolution engine (CUDA-core FFMA, fp32 accumulate).
//
// One templated kernel, pluggable A/B tile loaders:
//   A_ROW      A[M][K] row-major                       (linear fwd / bwd-data)
//   A_COL      A given as [K][M] (A^T row-major)       (linear weight-grad: dW = dY^T X)
//   A_IM2COL   A = im2col(X[nimg][H][W][C]) (NHWC)     (conv fwd, conv bwd-data with flipped W)
//   A_IM2COL_T A^T = im2col(X), reduce over pixels     (conv weight-grad)
//   B_KN       B[K][N] row-major      B_NK   B given as [N][K] (weights [out][in])
// Batched via blockIdx.z (batch * splitk); split-K results are reduced with fp32 atomics.
//
// Replaces, on the hot path, the cuDNN conv / cuBLAS GEMM calls PyTorch dispatches for the
// reference's nn.Conv2d / nn.Linear / GRU projections (model/tsrn.py:596-623, 876-888, 1070-1071;
// model/transformer_v2.py:453-458, 785-790).
#include <stdlib.h>
#include "common.cuh"
#include "gemm_params.cuh"

namespace {

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

template <int BM, int BN, int BK, int TM, int TN, int AMODE, int BMODE>
__global__ void __launch_bounds__((BM / TM) * (BN / TN)) gemm_kernel(const GemmP p) {
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int LDAS = BM + 4;
  constexpr int LDBS = BN + 4;
  constexpr int KQ = BK / 4;
  constexpr int A_V = BM * BK / 4;
  constexpr int B_V = BK * BN / 4;
  constexpr int A_IT = (A_V + NT - 1) / NT;
  constexpr int B_IT = (B_V + NT - 1) / NT;
  static_assert(TN == 4, "TN must be 4");
  static_assert(TM % 4 == 0, "TM must be a multiple of 4");
  static_assert(NT % KQ == 0, "thread count must be a multiple of BK/4");

  __shared__ __align__(16) float As[2][BK][LDAS];
  __shared__ __align__(16) float Bs[2][BK][LDBS];
  __shared__ int s_y[(AMODE == A_IM2COL) ? BM : 1];
  __shared__ int s_x[(AMODE == A_IM2COL) ? BM : 1];
  __shared__ int s_n[(AMODE == A_IM2COL) ? BM : 1];

  const int tid = threadIdx.x;
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
  const bool vecA = (p.flags & F_VECA) != 0;
  const bool vecB = (p.flags & F_VECB) != 0;
  const int HW = p.cH * p.cW;

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
    __syncthreads();
  }

  float4 ra[A_IT];
  float4 rb[B_IT];

  auto load_tiles = [&](int kb) {
    // ---------------- A
    if (AMODE == A_ROW) {
#pragma unroll
      for (int it = 0; it < A_IT; ++it) {
        int idx = tid + it * NT;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx < A_V) {
          int row = idx / KQ, kq = idx % KQ;
          int gm = m0 + row, gk = kb + kq * 4;
          if (gm < p.M && gk < kend) {
            const float* src = A + (long long)gm * p.lda + gk;
            if (vecA && gk + 3 < kend) {
              v = ldg4(src);
            } else {
              v.x = __ldg(src);
              if (gk + 1 < kend) v.y = __ldg(src + 1);
              if (gk + 2 < kend) v.z = __ldg(src + 2);
              if (gk + 3 < kend) v.w = __ldg(src + 3);
            }
          }
        }
        ra[it] = v;
      }
    } else if (AMODE == A_IM2COL) {
      const int kq = tid % KQ;
      const int gk = kb + kq * 4;
      unsigned tap = fd_div((unsigned)gk, p.fdC);
      int ci = gk - (int)tap * p.cC;
      unsigned ky = fd_div(tap, p.fdKW);
      int kx = (int)tap - (int)ky * p.KW;
      const int dy = (int)ky - p.padH, dx = kx - p.padW;
      const bool kok = gk < kend;
#pragma unroll
      for (int it = 0; it < A_IT; ++it) {
        int idx = tid + it * NT;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx < A_V) {
          int row = idx / KQ;
          int iy = s_y[row] + dy, ix = s_x[row] + dx;
          if (kok && iy >= 0 && iy < p.cH && ix >= 0 && ix < p.cW) {
            long long off = ((long long)(s_n[row] + iy * p.cW + ix)) * p.cC + ci;
            v = ldg4(A + off);
          }
        }
        ra[it] = v;
      }
    } else if (AMODE == A_COL) {
#pragma unroll
      for (int it = 0; it < A_IT; ++it) {
        int idx = tid + it * NT;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx < A_V) {
          int kk = idx / (BM / 4), mq = idx % (BM / 4);
          int gk = kb + kk, gm = m0 + mq * 4;
          if (gk < kend && gm < p.M) {
            const float* src = A + (long long)gk * p.lda + gm;
            if (vecA && gm + 3 < p.M) {
              v = ldg4(src);
            } else {
              v.x = __ldg(src);
              if (gm + 1 < p.M) v.y = __ldg(src + 1);
              if (gm + 2 < p.M) v.z = __ldg(src + 2);
              if (gm + 3 < p.M) v.w = __ldg(src + 3);
            }
          }
        }
        ra[it] = v;
      }
    } else {  // A_IM2COL_T : M index = (tap, ci), K index = pixel
#pragma unroll
      for (int it = 0; it < A_IT; ++it) {
        int idx = tid + it * NT;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx < A_V) {
          int kk = idx / (BM / 4), mq = idx % (BM / 4);
          int gp = kb + kk, gi = m0 + mq * 4;
          if (gp < kend && gi < p.M) {
            unsigned tap = fd_div((unsigned)gi, p.fdC);
            int ci = gi - (int)tap * p.cC;
            unsigned ky = fd_div(tap, p.fdKW);
            int kx = (int)tap - (int)ky * p.KW;
            unsigned n = fd_div((unsigned)gp, p.fdHW);
            unsigned rem = (unsigned)gp - n * (unsigned)HW;
            unsigned y = fd_div(rem, p.fdW);
            int x = (int)(rem - y * (unsigned)p.cW);
            int iy = (int)y + (int)ky - p.padH, ix = x + kx - p.padW;
            if (iy >= 0 && iy < p.cH && ix >= 0 && ix < p.cW) {
              long long off = ((long long)((int)(n * (unsigned)HW) + iy * p.cW + ix)) * p.cC + ci;
              v = ldg4(A + off);
            }
          }
        }
        ra[it] = v;
      }
    }
    // ---------------- B
    if (BMODE == B_KN) {
#pragma unroll
      for (int it = 0; it < B_IT; ++it) {
        int idx = tid + it * NT;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx < B_V) {
          int kk = idx / (BN / 4), nq = idx % (BN / 4);
          int gk = kb + kk, gn = n0 + nq * 4;
          if (gk < kend && gn < p.N) {
            const float* src = B + (long long)gk * p.ldb + gn;
            if (vecB && gn + 3 < p.N) {
              v = ldg4(src);
            } else {
              v.x = __ldg(src);
              if (gn + 1 < p.N) v.y = __ldg(src + 1);
              if (gn + 2 < p.N) v.z = __ldg(src + 2);
              if (gn + 3 < p.N) v.w = __ldg(src + 3);
            }
          }
        }
        rb[it] = v;
      }
    } else {
#pragma unroll
      for (int it = 0; it < B_IT; ++it) {
        int idx = tid + it * NT;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx < B_V) {
          int n = idx / KQ, kq = idx % KQ;
          int gn = n0 + n, gk = kb + kq * 4;
          if (gn < p.N && gk < kend) {
            const float* src = B + (long long)gn * p.ldb + gk;
            if (vecB && gk + 3 < kend) {
              v = ldg4(src);
            } else {
              v.x = __ldg(src);
              if (gk + 1 < kend) v.y = __ldg(src + 1);
              if (gk + 2 < kend) v.z = __ldg(src + 2);
              if (gk + 3 < kend) v.w = __ldg(src + 3);
            }
          }
        }
        rb[it] = v;
      }
    }
  };

  auto store_tiles = [&](int buf) {
    if (AMODE == A_ROW || AMODE == A_IM2COL) {
#pragma unroll
      for (int it = 0; it < A_IT; ++it) {
        int idx = tid + it * NT;
        if (idx < A_V) {
          int row = idx / KQ, kq = idx % KQ;
          As[buf][kq * 4 + 0][row] = ra[it].x;
          As[buf][kq * 4 + 1][row] = ra[it].y;
          As[buf][kq * 4 + 2][row] = ra[it].z;
          As[buf][kq * 4 + 3][row] = ra[it].w;
        }
      }
    } else {
#pragma unroll
      for (int it = 0; it < A_IT; ++it) {
        int idx = tid + it * NT;
        if (idx < A_V) {
          int kk = idx / (BM / 4), mq = idx % (BM / 4);
          *reinterpret_cast<float4*>(&As[buf][kk][mq * 4]) = ra[it];
        }
      }
    }
    if (BMODE == B_KN) {
#pragma unroll
      for (int it = 0; it < B_IT; ++it) {
        int idx = tid + it * NT;
        if (idx < B_V) {
          int kk = idx / (BN / 4), nq = idx % (BN / 4);
          *reinterpret_cast<float4*>(&Bs[buf][kk][nq * 4]) = rb[it];
        }
      }
    } else {
#pragma unroll
      for (int it = 0; it < B_IT; ++it) {
        int idx = tid + it * NT;
        if (idx < B_V) {
          int n = idx / KQ, kq = idx % KQ;
          Bs[buf][kq * 4 + 0][n] = rb[it].x;
          Bs[buf][kq * 4 + 1][n] = rb[it].y;
          Bs[buf][kq * 4 + 2][n] = rb[it].z;
          Bs[buf][kq * 4 + 3][n] = rb[it].w;
        }
      }
    }
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int tx = tid % (BN / TN);
  const int ty = tid / (BN / TN);
  const int nk = (kend > kbeg) ? (kend - kbeg + BK - 1) / BK : 0;

  if (nk > 0) {
    load_tiles(kbeg);
    store_tiles(0);
  }
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles(kbeg + (kt + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        float4 t = *reinterpret_cast<const float4*>(&As[buf][kk][ty * TM + i]);
        a[i] = t.x;
        a[i + 1] = t.y;
        a[i + 2] = t.z;
        a[i + 3] = t.w;
      }
      {
        float4 t = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * TN]);
        b[0] = t.x;
        b[1] = t.y;
        b[2] = t.z;
        b[3] = t.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) store_tiles(buf ^ 1);
    __syncthreads();
  }

  // ---------------- epilogue
  const bool atomic = (p.flags & F_ATOMIC) != 0;
  const bool accum = (p.flags & F_ACCUM) != 0;
  const bool relu = (p.flags & F_RELU) != 0;
  const bool vecC = (p.flags & F_VECC) != 0;
  const int gn0 = n0 + tx * TN;
  float bv[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) bv[j] = (bias && zs == 0 && gn0 + j < p.N) ? __ldg(bias + gn0 + j) : 0.f;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int gm = m0 + ty * TM + i;
    if (gm >= p.M) continue;
    float* dst = C + (long long)gm * p.ldc + gn0;
    if (atomic) {
#pragma unroll
      for (int j = 0; j < TN; ++j)
        if (gn0 + j < p.N) atomicAdd(dst + j, acc[i][j] + bv[j]);
    } else if (vecC && gn0 + 3 < p.N) {
      float4 v = make_float4(acc[i][0] + bv[0], acc[i][1] + bv[1], acc[i][2] + bv[2], acc[i][3] + bv[3]);
      if (accum) {
        float4 o = *reinterpret_cast<const float4*>(dst);
        v.x += o.x;
        v.y += o.y;
        v.z += o.z;
        v.w += o.w;
      }
      if (relu) {
        v.x = fmaxf(v.x, 0.f);
        v.y = fmaxf(v.y, 0.f);
        v.z = fmaxf(v.z, 0.f);
        v.w = fmaxf(v.w, 0.f);
      }
      *reinterpret_cast<float4*>(dst) = v;
    } else {
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        if (gn0 + j < p.N) {
          float v = acc[i][j] + bv[j];
          if (accum) v += dst[j];
          if (relu) v = fmaxf(v, 0.f);
          dst[j] = v;
        }
      }
    }
  }
}

template <int BM, int BN, int BK, int TM, int TN>
static int launch_cfg(const GemmP& p, int amode, int bmode, cudaStream_t st) {
  dim3 grid(ceil_div(p.M, BM), ceil_div(p.N, BN), p.batch * p.splitk);
  dim3 block((BM / TM) * (BN / TN));
#define TATT_GEMM_CASE(AM, BMo)                                             \
  if (amode == AM && bmode == BMo) {                                        \
    gemm_kernel<BM, BN, BK, TM, TN, AM, BMo><<<grid, block, 0, st>>>(p);    \
    TATT_LAUNCH_CHECK("gemm_kernel");                                       \
    return 0;                                                               \
  }
  TATT_GEMM_CASE(A_ROW, B_NK)
  TATT_GEMM_CASE(A_ROW, B_KN)
  TATT_GEMM_CASE(A_COL, B_KN)
  TATT_GEMM_CASE(A_IM2COL, B_KN)
  TATT_GEMM_CASE(A_IM2COL_T, B_KN)
#undef TATT_GEMM_CASE
  return tatt_set_error("gemm: unsupported loader combination amode=%d bmode=%d", amode, bmode);
}

static inline bool aligned16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

static int run_gemm(GemmP p, int amode, int bmode, bool want_split, cudaStream_t st) {
  if (p.M <= 0 || p.N <= 0) return 0;
  // vector-access eligibility
  int fl = p.flags & (F_ACCUM | F_RELU | F_APLANES | F_BPLANES | F_BF16 | F_A_VALID);
  bool va, vb;
  if (amode == A_ROW || amode == A_COL)
    va = (p.lda % 4 == 0) && aligned16(p.A) && (p.sA % 4 == 0);
  else
    va = true;
  vb = (p.ldb % 4 == 0) && aligned16(p.B) && (p.sB % 4 == 0);
  bool vc = (p.ldc % 4 == 0) && aligned16(p.C) && (p.sC % 4 == 0);
  if (va) fl |= F_VECA;
  if (vb) fl |= F_VECB;
  if (vc) fl |= F_VECC;
  p.flags = fl;

  // tensor-core path first (tc_gemm.cu); -1 = shape not eligible -> FFMA kernels below
  static const bool tc_on = []() {
    const char* e = getenv("TATT_TC");
    return !(e && e[0] == '0');
  }();
  static const bool tc2_on = []() {
    const char* e = getenv("TATT_TC2");
    return !(e && e[0] == '0');
  }();
  if (tc_on && tc2_on && !p.no_tc && (p.ws || (p.flags & (F_APLANES | F_BPLANES)))) {
    int rc = tatt_tc2_gemm_launch(p, amode, bmode, want_split, p.ws, p.ws_bytes, st);
    if (rc >= 0) return rc;
  }
  if (tc_on && !p.no_tc) {
    int rc = tatt_tc_gemm_launch(p, amode, bmode, want_split, st);
    if (rc >= 0) return rc;
  }

  // config selection
  int cfg;  // 0: 128x64 (256 thr)  1: 64x64 (256 thr)  2: 256x4 (32 thr)
  if (p.N <= 4) {
    cfg = 2;
  } else if (want_split) {
    cfg = (p.M > 64) ? 0 : 1;
  } else {
    long long ctas_l = (long long)ceil_div(p.M, 128) * ceil_div(p.N, 64) * p.batch;
    cfg = (ctas_l >= 148) ? 0 : 1;
  }
  const int BKc = (cfg == 2) ? 8 : 16;
  const int BMc = (cfg == 0) ? 128 : (cfg == 1 ? 64 : 256);
  const int BNc = (cfg == 2) ? 4 : 64;
  p.splitk = 1;
  p.kper = ((p.K + BKc - 1) / BKc) * BKc;
  if (want_split) {
    long long tiles = (long long)ceil_div(p.M, BMc) * ceil_div(p.N, BNc) * p.batch;
    long long target = 148LL * 4;
    int sk = (int)((target + tiles - 1) / tiles);
    int maxsk = ceil_div(p.K, BKc * 8);
    if (sk > maxsk) sk = maxsk;
    if (sk < 1) sk = 1;
    int kper = ceil_div(p.K, sk);
    kper = ((kper + BKc - 1) / BKc) * BKc;
    sk = ceil_div(p.K, kper);
    p.splitk = sk;
    p.kper = kper;
    p.flags |= F_ATOMIC;
    p.flags &= ~(F_RELU | F_ACCUM);
  }
  if (cfg == 0) return launch_cfg<128, 64, 16, 8, 4>(p, amode, bmode, st);
  if (cfg == 1) return launch_cfg<64, 64, 16, 4, 4>(p, amode, bmode, st);
  return launch_cfg<256, 4, 8, 8, 4>(p, amode, bmode, st);
}

// ------------------------------------------------------------------ weight (un)packing kernels
__global__ void conv_pack_kernel(const float* __restrict__ W, float* __restrict__ Wt, int Cout, int Cin, int KH,
                                 int KW, int CinP, int CoutP, int flip) {
  // flip == 0: Wt[((ky*KW+kx)*CinP + ci)*CoutP + co] = W[co][ci][ky][kx]
  // flip == 1: Wt[((ky*KW+kx)*CoutP + co)*CinP + ci] = W[co][ci][KH-1-ky][KW-1-kx]
  long long total = (long long)KH * KW * CinP * CoutP;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int co, ci, tap;
    if (!flip) {
      co = (int)(i % CoutP);
      long long r = i / CoutP;
      ci = (int)(r % CinP);
      tap = (int)(r / CinP);
    } else {
      ci = (int)(i % CinP);
      long long r = i / CinP;
      co = (int)(r % CoutP);
      tap = (int)(r / CoutP);
    }
    int ky = tap / KW, kx = tap % KW;
    if (flip) {
      ky = KH - 1 - ky;
      kx = KW - 1 - kx;
    }
    float v = 0.f;
    if (co < Cout && ci < Cin) v = W[(((long long)co * Cin + ci) * KH + ky) * KW + kx];
    Wt[i] = v;
  }
}

__global__ void conv_unpack_grad_kernel(const float* __restrict__ dWt, float* __restrict__ dW, int Cout, int Cin,
                                        int KH, int KW, int CinP, int CoutP) {
  long long total = (long long)Cout * Cin * KH * KW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int kx = (int)(i % KW);
    long long r = i / KW;
    int ky = (int)(r % KH);
    r /= KH;
    int ci = (int)(r % Cin);
    int co = (int)(r / Cin);
    dW[i] = dWt[((long long)(ky * KW + kx) * CinP + ci) * CoutP + co];
  }
}

// "kx-expansion" of a KHxKW convolution with very few output channels (the 9x9 64->4 output conv):
// T[p][kx*CoP+co] = sum_{ky,ci} X[p + (ky-padH, 0)][ci] * W[co][ci][ky][kx]   (a K = KH*Cin, N = KW*CoP GEMM),
// out[y][x][co] = bias[co] + sum_kx T[y][x + kx - padW][kx*CoP+co]              (horizontal shift-sum).
__global__ void kxexp_pack_kernel(const float* __restrict__ W, float* __restrict__ Wt, int Cout, int Cin, int KH,
                                  int KW, int CinP, int CoP) {
  long long total = (long long)KH * CinP * KW * CoP;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int co = (int)(i % CoP);
    long long r = i / CoP;
    int kx = (int)(r % KW);
    r /= KW;
    int ci = (int)(r % CinP);
    int ky = (int)(r / CinP);
    Wt[i] = (co < Cout && ci < Cin) ? W[(((long long)co * Cin + ci) * KH + ky) * KW + kx] : 0.f;
  }
}
__global__ void kxexp_unpack_grad_kernel(const float* __restrict__ dWt, float* __restrict__ dW, int Cout, int Cin,
                                         int KH, int KW, int CinP, int CoP) {
  long long total = (long long)Cout * Cin * KH * KW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int kx = (int)(i % KW);
    long long r = i / KW;
    int ky = (int)(r % KH);
    r /= KH;
    int ci = (int)(r % Cin);
    int co = (int)(r / Cin);
    dW[i] = dWt[(((long long)ky * CinP + ci) * KW + kx) * CoP + co];
  }
}
// out[p][co] = bias[co] + sum_kx T[(y, x+kx-padW)][kx*CoP+co]
__global__ void kxexp_reduce_kernel(const float* __restrict__ T, const float* __restrict__ bias,
                                    float* __restrict__ out, long long P, int W, int KW, int CoP, int padW) {
  long long total = P * CoP;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int co = (int)(i % CoP);
    long long pix = i / CoP;
    int x = (int)(pix % W);
    float a = bias ? bias[co] : 0.f;
    for (int kx = 0; kx < KW; ++kx) {
      int xs = x + kx - padW;
      if (xs >= 0 && xs < W) a += T[(pix + (kx - padW)) * (long long)(KW * CoP) + kx * CoP + co];
    }
    out[i] = a;
  }
}
// dT[(y,x')][kx*CoP+co] = dOut[(y, x' - (kx-padW))][co]  (0 outside the row)
__global__ void kxexp_expand_kernel(const float* __restrict__ dOut, float* __restrict__ dT, long long P, int W,
                                    int KW, int CoP, int padW) {
  const int NC = KW * CoP;
  long long total = P * NC;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % NC);
    long long pix = i / NC;
    int kx = c / CoP, co = c - kx * CoP;
    int x = (int)(pix % W);
    int xo = x - (kx - padW);
    dT[i] = (xo >= 0 && xo < W) ? dOut[(pix - (kx - padW)) * CoP + co] : 0.f;
  }
}

__global__ void colsum_kernel(const float* __restrict__ X, long long ldx, float* __restrict__ out, long long P,
                              int C, int rows_per_cta) {
  __shared__ float red[4][64];
  int c = blockIdx.y * 64 + (threadIdx.x & 63);
  int ty = threadIdx.x >> 6;
  long long r0 = (long long)blockIdx.x * rows_per_cta;
  long long r1 = r0 + rows_per_cta;
  if (r1 > P) r1 = P;
  float s = 0.f;
  if (c < C)
    for (long long r = r0 + ty; r < r1; r += 4) s += X[r * ldx + c];
  red[ty][threadIdx.x & 63] = s;
  __syncthreads();
  if (ty == 0 && c < C) {
    float t = red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x];
    atomicAdd(out + c, t);
  }
}

}  // namespace

// ====================================================================== C-ABI
extern "C" {

int tatt_gemm(int amode, int bmode, const float* A, long long lda, const float* B, long long ldb, float* C,
              long long ldc, const float* bias, int M, int N, int K, int batch, long long sA, long long sB,
              long long sC, long long sBias, int flags, long long loA, long long loB, void* ws, long long ws_bytes,
              void* stream) {
  TATT_REQUIRE(amode == A_ROW || amode == A_COL, "tatt_gemm: amode must be 0 (row) or 1 (col)");
  TATT_REQUIRE(bmode == B_KN || bmode == B_NK, "tatt_gemm: bad bmode");
  TATT_REQUIRE(batch >= 1 && M >= 0 && N >= 0 && K >= 0, "tatt_gemm: bad sizes");
  GemmP p = {};
  p.A = A; p.B = B; p.C = C; p.bias = bias;
  p.M = M; p.N = N; p.K = K;
  p.lda = lda; p.ldb = ldb; p.ldc = ldc;
  p.sA = sA; p.sB = sB; p.sC = sC; p.sBias = sBias;
  p.batch = batch;
  p.flags = flags & (F_ACCUM | F_RELU | F_APLANES | F_BPLANES | F_BF16);
  p.loA = loA;
  p.loB = loB;
  p.no_tc = (flags & F_FP32) ? 1 : 0;
  p.ws = ws;
  p.ws_bytes = ws_bytes;
  if (flags & (F_APLANES | F_BPLANES)) {
    const bool want = (flags & F_ATOMIC) != 0;
    if (want && (flags & F_ZEROC)) {
      TATT_REQUIRE(ldc == N && (batch == 1 || sC == (long long)M * N), "tatt_gemm: F_ZEROC needs a dense C");
      TATT_CUDA(cudaMemsetAsync(C, 0, sizeof(float) * (size_t)batch * M * N, (cudaStream_t)stream));
    }
    if ((p.ldc % 4 == 0) && ((((uintptr_t)C) & 15) == 0) && (p.sC % 4 == 0)) p.flags |= F_VECC;
    int rc = tatt_tc2_gemm_launch(p, amode, bmode, want, ws, ws_bytes, (cudaStream_t)stream);
    if (rc < 0) return tatt_set_error("tatt_gemm: operand planes need the tcgen05 path");
    return rc;
  }
  bool split = (flags & F_ATOMIC) != 0;
  if (flags & F_ZEROC) {  // dense C only
    TATT_REQUIRE(ldc == N && (batch == 1 || sC == (long long)M * N), "tatt_gemm: F_ZEROC needs a dense C");
    TATT_CUDA(cudaMemsetAsync(C, 0, sizeof(float) * (size_t)batch * M * N, (cudaStream_t)stream));
  }
  return run_gemm(p, amode, bmode, split, (cudaStream_t)stream);
}

// bf16 hi/lo operand planes for tatt_gemm flags 256 (A) / 512 (B): planes[rows][round8(cols)] or, transposed,
// planes[cols][round8(rows)]
int tatt_split_bf16(const float* src, long long ld, long long rows, int cols, int transpose, void* hi, void* lo,
                    float* colsum, void* stream) {
  return tatt_tc2_split(src, ld, rows, cols, transpose, hi, lo, colsum, (cudaStream_t)stream);
}

int tatt_colsum(const float* X, long long ldx, float* out, long long P, int C, int zero_first, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (zero_first) TATT_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * C, st));
  if (P <= 0 || C <= 0) return 0;
  int rows = 256;
  if (P > 256LL * 2048) rows = (int)((P + 2047) / 2048);
  dim3 grid(ceil_div(P, rows), ceil_div(C, 64));
  colsum_kernel<<<grid, 256, 0, st>>>(X, ldx, out, P, C, rows);
  TATT_LAUNCH_CHECK("colsum_kernel");
  return 0;
}

int tatt_conv_weight_pack(const float* W, float* Wt, int Cout, int Cin, int KH, int KW, int CinP, int CoutP,
                          int flip, void* stream) {
  long long total = (long long)KH * KW * CinP * CoutP;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 4096) blocks = 4096;
  conv_pack_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(W, Wt, Cout, Cin, KH, KW, CinP, CoutP, flip);
  TATT_LAUNCH_CHECK("conv_pack_kernel");
  return 0;
}

int tatt_conv_weight_unpack_grad(const float* dWt, float* dW, int Cout, int Cin, int KH, int KW, int CinP,
                                 int CoutP, void* stream) {
  long long total = (long long)Cout * Cin * KH * KW;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 4096) blocks = 4096;
  conv_unpack_grad_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(dWt, dW, Cout, Cin, KH, KW, CinP, CoutP);
  TATT_LAUNCH_CHECK("conv_unpack_grad_kernel");
  return 0;
}

int tatt_conv_kxexp_pack(const float* W, float* Wt, int Cout, int Cin, int KH, int KW, int CinP, int CoP,
                         void* stream) {
  long long total = (long long)KH * CinP * KW * CoP;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 4096) blocks = 4096;
  kxexp_pack_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(W, Wt, Cout, Cin, KH, KW, CinP, CoP);
  TATT_LAUNCH_CHECK("kxexp_pack_kernel");
  return 0;
}
int tatt_conv_kxexp_unpack_grad(const float* dWt, float* dW, int Cout, int Cin, int KH, int KW, int CinP, int CoP,
                                void* stream) {
  long long total = (long long)Cout * Cin * KH * KW;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 4096) blocks = 4096;
  kxexp_unpack_grad_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(dWt, dW, Cout, Cin, KH, KW, CinP, CoP);
  TATT_LAUNCH_CHECK("kxexp_unpack_grad_kernel");
  return 0;
}
int tatt_conv_kxexp_reduce(const float* T, const float* bias, float* out, long long P, int W, int KW, int CoP,
                           int padW, void* stream) {
  if (P <= 0) return 0;
  long long total = P * CoP;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  kxexp_reduce_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(T, bias, out, P, W, KW, CoP, padW);
  TATT_LAUNCH_CHECK("kxexp_reduce_kernel");
  return 0;
}
int tatt_conv_kxexp_expand(const float* dOut, float* dT, long long P, int W, int KW, int CoP, int padW,
                           void* stream) {
  if (P <= 0) return 0;
  long long total = P * KW * CoP;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  kxexp_expand_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(dOut, dT, P, W, KW, CoP, padW);
  TATT_LAUNCH_CHECK("kxexp_expand_kernel");
  return 0;
}

// Y[nimg*H*W][Cout] (=|+=) im2col(X[nimg][H][W][Cin]) * Wt[KH*KW*Cin][Cout] + bias ; stride 1.
int tatt_conv2d_igemm(const float* X, const float* Wt, const float* bias, float* Y, int nimg, int H, int W, int Cin,
                      int Cout, int KH, int KW, int padH, int padW, int flags, void* ws, long long ws_bytes,
                      void* stream) {
  TATT_REQUIRE(Cin % 4 == 0, "conv2d_igemm: Cin (%d) must be a multiple of 4 (pad channels)", Cin);
  TATT_REQUIRE((long long)nimg * H * W < (1LL << 31), "conv2d_igemm: too many pixels");
  if (KH == 3 && KW == 3 && padH == 1 && padW == 1 && Cin % 64 == 0 && Cin <= 256 && Cout % 64 == 0 && Cout <= 256 && ws &&
      !(flags & (F_ACCUM | F_RELU | F_FP32))) {
    int rc = tatt_tc3_conv3x3_launch(X, Wt, bias, Y, nimg, H, W, Cin, Cout, (flags & F_BF16) ? 1 : 0, (flags & F_A_VALID) ? 1 : 0, ws,
                                     ws_bytes, nullptr, (cudaStream_t)stream);
    if (rc >= 0) return rc;
  }
  GemmP p = {};
  p.A = X; p.B = Wt; p.C = Y; p.bias = bias;
  p.M = nimg * H * W; p.N = Cout; p.K = KH * KW * Cin;
  p.lda = 0; p.ldb = Cout; p.ldc = Cout;
  p.batch = 1;
  p.flags = flags & (F_ACCUM | F_RELU | F_BF16 | F_A_VALID);
  p.no_tc = (flags & F_FP32) ? 1 : 0;
  p.ws = ws;
  p.ws_bytes = ws_bytes;
  p.cH = H; p.cW = W; p.cC = Cin; p.KH = KH; p.KW = KW; p.padH = padH; p.padW = padW;
  p.fdHW = make_fd(H * W); p.fdW = make_fd(W); p.fdC = make_fd(Cin); p.fdKW = make_fd(KW);
  return run_gemm(p, A_IM2COL, B_KN, false, (cudaStream_t)stream);
}

// 3x3 / 64 -> 64 convolution with the BatchNorm statistics of its output from the same pass: stats[c][0..63] = sum over
// the pixels CTA c stored, stats[c][64..127] = sum of squares (TATT_CONV_STATS_ROWS rows of 128 floats, zeroed here; rows
// of CTAs that do not exist stay zero).  Only the persistent TMA kernel does this: returns an error when the shape is not
// served by it (W % 128, H % 2, workspace) -- tatt_conv3x3_stats_supported() tells.
int tatt_conv3x3_stats_supported(int H, int W, int Cin, int Cout) {
  static const bool off = []() {
    const char* e = getenv("TATT_TMA");
    return e && atoi(e) != 3;
  }();
  return (!off && Cin == 64 && Cout == 64 && W % 128 == 0 && H % 2 == 0) ? 1 : 0;
}
int tatt_conv3x3_stats(const float* X, const float* Wt, const float* bias, float* Y, int nimg, int H, int W, int flags,
                       void* ws, long long ws_bytes, void* stats, void* stream) {
  TATT_REQUIRE(stats != nullptr && ws != nullptr && !(flags & (F_ACCUM | F_RELU | F_FP32)), "conv3x3_stats: bad arguments");
  TATT_REQUIRE((long long)nimg * H * W < (1LL << 31), "conv3x3_stats: too many pixels");
  cudaStream_t st = (cudaStream_t)stream;
  TATT_CUDA(cudaMemsetAsync(stats, 0, sizeof(float) * 128 * TATT_CONV_STATS_ROWS, st));
  int dev = 0, nsm = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  TATT_REQUIRE(nsm <= TATT_CONV_STATS_ROWS, "conv3x3_stats: device has more SMs (%d) than statistics rows", nsm);
  int rc = tatt_tc3_conv3x3_launch(X, Wt, bias, Y, nimg, H, W, 64, 64, (flags & F_BF16) ? 1 : 0, (flags & F_A_VALID) ? 1 : 0, ws,
                                   ws_bytes, reinterpret_cast<float*>(stats), st);
  if (rc < 0) return tatt_set_error("conv3x3_stats: shape [%d,%d,%d] is not served by the TMA kernel", nimg, H, W);
  return rc;
}

// dWt[KH*KW*Cin][Cout] = im2col(X)^T * dY[nimg*H*W][Cout]   (zeroed here, split-K atomics)
int tatt_conv2d_wgrad(const float* X, const float* dY, float* dWt, int nimg, int H, int W, int Cin, int Cout,
                      int KH, int KW, int padH, int padW, int flags, void* ws, long long ws_bytes, void* stream) {
  TATT_REQUIRE(Cin % 4 == 0, "conv2d_wgrad: Cin (%d) must be a multiple of 4", Cin);
  TATT_REQUIRE((long long)nimg * H * W < (1LL << 31), "conv2d_wgrad: too many pixels");
  cudaStream_t st = (cudaStream_t)stream;
  TATT_CUDA(cudaMemsetAsync(dWt, 0, sizeof(float) * (size_t)KH * KW * Cin * Cout, st));
  if (KH == 3 && KW == 3 && padH == 1 && padW == 1 && Cin == 64 && Cout % 64 == 0 && Cout <= 256 && ws &&
      !(flags & F_FP32)) {
    int rc = tatt_tc3_conv3x3_wgrad_launch(X, dY, dWt, nimg, H, W, Cout, (flags & F_BF16) ? 1 : 0,
                                           (flags & F_A_VALID) ? 1 : 0, (flags & F_B_VALID) ? 1 : 0, ws, ws_bytes, st);
    if (rc >= 0) return rc;
  }
  // (flag 4096 is a hint for the TMA kernel only: the generic engine below splits dY itself)
  GemmP p = {};
  p.A = X; p.B = dY; p.C = dWt; p.bias = nullptr;
  p.M = KH * KW * Cin; p.N = Cout; p.K = nimg * H * W;
  p.lda = 0; p.ldb = Cout; p.ldc = Cout;
  p.batch = 1;
  p.flags = flags & (F_BF16 | F_A_VALID);
  p.no_tc = (flags & F_FP32) ? 1 : 0;
  p.ws = ws;
  p.ws_bytes = ws_bytes;
  p.cH = H; p.cW = W; p.cC = Cin; p.KH = KH; p.KW = KW; p.padH = padH; p.padW = padW;
  p.fdHW = make_fd(H * W); p.fdW = make_fd(W); p.fdC = make_fd(Cin); p.fdKW = make_fd(KW);
  return run_gemm(p, A_IM2COL_T, B_KN, true, st);
}

}  // extern "C"
