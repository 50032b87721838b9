// Parameter block shared by the FFMA (gemm.cu) and tcgen05 (tc_gemm.cu) GEMM / implicit-GEMM engines.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct FastDiv {
  unsigned mul, shr, d;
};
static inline FastDiv make_fd(unsigned d) {
  FastDiv f;
  f.d = d;
  if (d <= 1) {
    f.mul = 0;
    f.shr = 0;
    f.d = 1;
    return f;
  }
  int lg = 31 - __builtin_clz(d);
  if (d & (d - 1)) lg += 1;  // ceil(log2(d))
  int p = 31 + lg;
  unsigned long long m = ((1ull << p) + d - 1) / d;
  f.mul = (unsigned)m;
  f.shr = (unsigned)(p - 32);
  return f;
}
__device__ __forceinline__ unsigned fd_div(unsigned n, const FastDiv& f) {
  return f.d == 1 ? n : (__umulhi(n, f.mul) >> f.shr);
}

enum { A_ROW = 0, A_COL = 1, A_IM2COL = 2, A_IM2COL_T = 3 };
enum { B_KN = 0, B_NK = 1 };
enum { F_ACCUM = 1, F_RELU = 2, F_ATOMIC = 4, F_VECA = 8, F_VECB = 16, F_VECC = 32, F_ZEROC = 64, F_FP32 = 128, F_APLANES = 256, F_BPLANES = 512, F_BF16 = 1024, F_A_VALID = 2048, F_B_VALID = 4096 };

struct GemmP {
  const float* A;
  const float* B;
  float* C;
  const float* bias;
  int M, N, K;
  long long lda, ldb, ldc;
  long long sA, sB, sC, sBias;
  int batch, splitk, kper;
  int flags;
  long long loA, loB;  // F_APLANES / F_BPLANES: element offset from the hi plane (A / B pointer) to the lo plane
  void* ws;            // optional workspace for the pre-split bf16 planes of the v2 tcgen05 engine
  long long ws_bytes;
  int no_tc;   // 1: force the fp32 FFMA kernels (ill-conditioned sub-graphs, e.g. the STN head)
  // conv geometry (IM2COL modes): X[nimg][cH][cW][cC]
  int cH, cW, cC, KH, KW, padH, padW;
  FastDiv fdHW, fdW, fdC, fdKW;
};


// tcgen05 path (tc_gemm.cu); returns 0 on success, -1 if the shape is not eligible (caller falls through to FFMA)
int tatt_tc_gemm_launch(GemmP p, int amode, int bmode, bool want_split, cudaStream_t st);
int tatt_tc2_split(const float* src, long long ld, long long rows, int cols, int transpose, void* hi, void* lo,
                   float* colsum, cudaStream_t st);
// v2 (tc2_gemm.cu): operands pre-split into bf16 planes in `ws`; same return convention
int tatt_tc2_gemm_launch(GemmP p, int amode, int bmode, bool want_split, void* ws, long long ws_bytes,
                         cudaStream_t st);

// TMA + halo-reuse kernel for the 3x3 convolutions with 64 k channels in and out (tc3_conv.cu); same return convention
int tatt_tc3_conv3x3_launch(const float* X, const float* Wt, const float* bias, float* Y, int nimg, int H, int W,
                            int Cin, int Cout, int single, int a_valid, void* ws, long long ws_bytes, float* stats,
                            cudaStream_t st);
int tatt_tc3_conv3x3_wgrad_launch(const float* X, const float* dY, float* dWt, int nimg, int H, int W, int Cout,
                                  int single, int a_valid, int b_valid, void* ws, long long ws_bytes, cudaStream_t st);
