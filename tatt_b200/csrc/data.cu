// Device half of the TextZoom collate: dataset/dataset.py:1266-1319 (resizeNormalize) --
// after the host-side PIL resize the reference does, per image, `ToTensor()` (uint8 HWC -> float CHW / 255) and, with
// mask=True, appends the binary mask `convert('L') -> 0 if L > mean(L) else 255 -> ToTensor`.  Here that is one kernel
// over the whole uint8 batch, bit-exact (integer luma, integer mean comparison, IEEE division by 255):
//   L = (19595 R + 38470 G + 7471 B + 32768) >> 16           (PIL's ImagingConvert rgb2l)
//   mask = (L * H * W > sum L) ? 0 : 1                        (== L > float64 mean, see tests)
#include "common.cuh"

namespace {

// one CTA per image: pass 1 block-reduces sum(L) in integers, pass 2 writes the CHW planes
__global__ void __launch_bounds__(256)
collate_u8_kernel(const unsigned char* __restrict__ img, float* __restrict__ out, int H, int W, int with_mask) {
  __shared__ unsigned long long red[8];
  __shared__ unsigned long long total_s;
  const int n = blockIdx.x, HW = H * W;
  const unsigned char* src = img + (long long)n * HW * 3;
  const int CO = with_mask ? 4 : 3;
  float* dst = out + (long long)n * CO * HW;
  if (with_mask) {
    unsigned long long s = 0;
    for (int i = threadIdx.x; i < HW; i += 256) {
      const unsigned int r = src[3 * i], g = src[3 * i + 1], b = src[3 * i + 2];
      s += (19595u * r + 38470u * g + 7471u * b + 0x8000u) >> 16;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long t = 0;
      for (int k = 0; k < 8; ++k) t += red[k];
      total_s = t;
    }
    __syncthreads();
  }
  const unsigned long long total = with_mask ? total_s : 0ull;
  for (int i = threadIdx.x; i < HW; i += 256) {
    const unsigned int r = src[3 * i], g = src[3 * i + 1], b = src[3 * i + 2];
    dst[i] = (float)r / 255.f;
    dst[HW + i] = (float)g / 255.f;
    dst[2 * HW + i] = (float)b / 255.f;
    if (with_mask) {
      const unsigned long long L = (19595u * r + 38470u * g + 7471u * b + 0x8000u) >> 16;
      dst[3 * HW + i] = (L * (unsigned long long)HW > total) ? 0.f : 1.f;
    }
  }
}

}  // namespace

extern "C" {

int tatt_collate_u8(const void* img, float* out, int N, int H, int W, int with_mask, void* stream) {
  TATT_REQUIRE(N >= 1 && H >= 1 && W >= 1, "collate_u8: bad shape [%d,%d,%d,3]", N, H, W);
  collate_u8_kernel<<<N, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const unsigned char*>(img), out, H, W, with_mask);
  TATT_LAUNCH_CHECK("collate_u8_kernel");
  return 0;
}

}  // extern "C"
