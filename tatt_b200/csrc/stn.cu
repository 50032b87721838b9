// Thin-plate-spline grid generation fused with the bilinear grid_sample warp, forward and backward.
// Reference: TPSSpatialTransformer.forward + grid_sample (model/tps_spatial_transformer.py:97-112, 10-18):
//   Y = cat(ctrl[20x2], 0[3x2]); M = inverse_kernel[23x23] Y; src = target_coordinate_repr[HWx23] M;
//   grid = 2*clamp(src,0,1)-1; F.grid_sample(x, grid)  (bilinear, zeros padding, align_corners=False).
// Image layout here is NHWC with 4 channels (RGB+mask, or RGB + one zero pad channel).
#include "common.cuh"

namespace {

constexpr int NCP = 20;       // control points
constexpr int NK = NCP + 3;   // TPS basis size

__device__ __forceinline__ void tps_mapping(const float* __restrict__ invK, const float* __restrict__ ctrl,
                                            double (*Mm)[2]) {
  if (threadIdx.x < NK * 2) {
    int k = threadIdx.x >> 1, xy = threadIdx.x & 1;
    double a = 0.0;   // the TPS system is ill-conditioned: accumulate the mapping in double
    for (int i = 0; i < NCP; ++i) a += (double)invK[k * NK + i] * (double)ctrl[i * 2 + xy];
    Mm[k][xy] = a;
  }
}

struct Bilin {
  int x0, y0;
  float wx1, wy1;
};
__device__ __forceinline__ Bilin bilin_setup(float sx, float sy, int H, int W) {
  float gx = 2.f * fminf(fmaxf(sx, 0.f), 1.f) - 1.f;
  float gy = 2.f * fminf(fmaxf(sy, 0.f), 1.f) - 1.f;
  float ix = ((gx + 1.f) * W - 1.f) * 0.5f;
  float iy = ((gy + 1.f) * H - 1.f) * 0.5f;
  float fx = floorf(ix), fy = floorf(iy);
  Bilin b;
  b.x0 = (int)fx;
  b.y0 = (int)fy;
  b.wx1 = ix - fx;
  b.wy1 = iy - fy;
  return b;
}
__device__ __forceinline__ float4 fetch4(const float* __restrict__ img, int y, int x, int H, int W) {
  if (x < 0 || x >= W || y < 0 || y >= H) return make_float4(0.f, 0.f, 0.f, 0.f);
  return __ldg(reinterpret_cast<const float4*>(img) + (long long)y * W + x);
}

__global__ void tps_sample_fwd_kernel(const float* __restrict__ X, const float* __restrict__ ctrl,
                                      const float* __restrict__ invK, const float* __restrict__ repr,
                                      float* __restrict__ OUT, float* __restrict__ SRC, int H, int W) {
  __shared__ double Mm[NK][2];
  const int n = blockIdx.y;
  tps_mapping(invK, ctrl + (long long)n * NCP * 2, Mm);
  __syncthreads();
  const int HW = H * W;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  double sxd = 0.0, syd = 0.0;
  const float* rp = repr + (long long)p * NK;
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    double r = (double)__ldg(rp + k);
    sxd += r * Mm[k][0];
    syd += r * Mm[k][1];
  }
  const float sx = (float)sxd, sy = (float)syd;
  if (SRC) {
    SRC[((long long)n * HW + p) * 2 + 0] = sx;
    SRC[((long long)n * HW + p) * 2 + 1] = sy;
  }
  Bilin b = bilin_setup(sx, sy, H, W);
  const float* img = X + (long long)n * HW * 4;
  float4 i00 = fetch4(img, b.y0, b.x0, H, W), i01 = fetch4(img, b.y0, b.x0 + 1, H, W);
  float4 i10 = fetch4(img, b.y0 + 1, b.x0, H, W), i11 = fetch4(img, b.y0 + 1, b.x0 + 1, H, W);
  float wx0 = 1.f - b.wx1, wy0 = 1.f - b.wy1;
  float w00 = wx0 * wy0, w01 = b.wx1 * wy0, w10 = wx0 * b.wy1, w11 = b.wx1 * b.wy1;
  float4 o;
  o.x = i00.x * w00 + i01.x * w01 + i10.x * w10 + i11.x * w11;
  o.y = i00.y * w00 + i01.y * w01 + i10.y * w10 + i11.y * w11;
  o.z = i00.z * w00 + i01.z * w01 + i10.z * w10 + i11.z * w11;
  o.w = i00.w * w00 + i01.w * w01 + i10.w * w10 + i11.w * w11;
  reinterpret_cast<float4*>(OUT)[(long long)n * HW + p] = o;
}

// one CTA per sample; dynamic smem: dsrc[HW][2]
__global__ void tps_sample_bwd_kernel(const float* __restrict__ X, const float* __restrict__ ctrl,
                                      const float* __restrict__ invK, const float* __restrict__ repr,
                                      const float* __restrict__ dOUT, float* __restrict__ dctrl, int H, int W) {
  extern __shared__ float dsrc[];
  __shared__ double Mm[NK][2];
  __shared__ float dM[NK][2];
  const int n = blockIdx.x;
  const int HW = H * W;
  tps_mapping(invK, ctrl + (long long)n * NCP * 2, Mm);
  __syncthreads();
  const float* img = X + (long long)n * HW * 4;
  for (int p = threadIdx.x; p < HW; p += blockDim.x) {
    double sxd = 0.0, syd = 0.0;
    const float* rp = repr + (long long)p * NK;
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      double r = (double)__ldg(rp + k);
      sxd += r * Mm[k][0];
      syd += r * Mm[k][1];
    }
    const float sx = (float)sxd, sy = (float)syd;
    Bilin b = bilin_setup(sx, sy, H, W);
    float4 i00 = fetch4(img, b.y0, b.x0, H, W), i01 = fetch4(img, b.y0, b.x0 + 1, H, W);
    float4 i10 = fetch4(img, b.y0 + 1, b.x0, H, W), i11 = fetch4(img, b.y0 + 1, b.x0 + 1, H, W);
    float4 g = __ldg(reinterpret_cast<const float4*>(dOUT) + (long long)n * HW + p);
    float wx0 = 1.f - b.wx1, wy0 = 1.f - b.wy1;
    float dix = g.x * ((i01.x - i00.x) * wy0 + (i11.x - i10.x) * b.wy1) +
                g.y * ((i01.y - i00.y) * wy0 + (i11.y - i10.y) * b.wy1) +
                g.z * ((i01.z - i00.z) * wy0 + (i11.z - i10.z) * b.wy1) +
                g.w * ((i01.w - i00.w) * wy0 + (i11.w - i10.w) * b.wy1);
    float diy = g.x * ((i10.x - i00.x) * wx0 + (i11.x - i01.x) * b.wx1) +
                g.y * ((i10.y - i00.y) * wx0 + (i11.y - i01.y) * b.wx1) +
                g.z * ((i10.z - i00.z) * wx0 + (i11.z - i01.z) * b.wx1) +
                g.w * ((i10.w - i00.w) * wx0 + (i11.w - i01.w) * b.wx1);
    // ix = ((g+1)W-1)/2, g = 2*clamp(s)-1  => d ix / d s = W inside [0,1], 0 outside
    dsrc[p * 2 + 0] = (sx >= 0.f && sx <= 1.f) ? dix * (float)W : 0.f;
    dsrc[p * 2 + 1] = (sy >= 0.f && sy <= 1.f) ? diy * (float)H : 0.f;
  }
  __syncthreads();
  // dM[k][xy] = sum_p repr[p][k] * dsrc[p][xy]  : warp w handles (k,xy) pairs w, w+8, ...
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  for (int pair = warp; pair < NK * 2; pair += nwarp) {
    int k = pair >> 1, xy = pair & 1;
    float a = 0.f;
    for (int p = lane; p < HW; p += 32) a = fmaf(__ldg(repr + (long long)p * NK + k), dsrc[p * 2 + xy], a);
    a = warp_sum(a);
    if (lane == 0) dM[k][xy] = a;
  }
  __syncthreads();
  if (threadIdx.x < NCP * 2) {
    int i = threadIdx.x >> 1, xy = threadIdx.x & 1;
    float a = 0.f;
    for (int k = 0; k < NK; ++k) a = fmaf(invK[k * NK + i], dM[k][xy], a);
    dctrl[(long long)n * NCP * 2 + i * 2 + xy] = a;
  }
}

}  // namespace

extern "C" {

// X [N][H][W][4] -> OUT [N][H][W][4]; ctrl [N][20][2]; invK [23][23]; repr [H*W][23]; SRC [N][H*W][2] or NULL
int tatt_tps_sample_fwd(const float* X, const float* ctrl, const float* invK, const float* repr, float* OUT,
                        float* SRC, int N, int H, int W, void* stream) {
  if (N <= 0) return 0;
  dim3 grid((H * W + 255) / 256, N);
  tps_sample_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(X, ctrl, invK, repr, OUT, SRC, H, W);
  TATT_LAUNCH_CHECK("tps_sample_fwd_kernel");
  return 0;
}

// gradient wrt the control points only (the LR image itself never requires grad on this path)
int tatt_tps_sample_bwd(const float* X, const float* ctrl, const float* invK, const float* repr, const float* dOUT,
                        float* dctrl, int N, int H, int W, void* stream) {
  if (N <= 0) return 0;
  size_t smem = sizeof(float) * 2 * (size_t)H * W;
  TATT_REQUIRE(smem <= 160 * 1024, "tps_sample_bwd: H*W=%d too large for the per-sample SMEM tile", H * W);
  if (smem > 40 * 1024)
    TATT_CUDA(cudaFuncSetAttribute(tps_sample_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tps_sample_bwd_kernel<<<N, 256, smem, (cudaStream_t)stream>>>(X, ctrl, invK, repr, dOUT, dctrl, H, W);
  TATT_LAUNCH_CHECK("tps_sample_bwd_kernel");
  return 0;
}

}  // extern "C"
