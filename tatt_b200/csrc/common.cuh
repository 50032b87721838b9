// Shared helpers for the tatt_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/tatt_b200.h"

int tatt_set_error(const char* fmt, ...);

#define TATT_REQUIRE(cond, ...)                                   \
  do {                                                            \
    if (!(cond)) return tatt_set_error(__VA_ARGS__);              \
  } while (0)

#define TATT_LAUNCH_CHECK(name)                                                    \
  do {                                                                             \
    cudaError_t e__ = cudaGetLastError();                                          \
    if (e__ != cudaSuccess)                                                        \
      return tatt_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
  } while (0)

#define TATT_CUDA(call)                                                               \
  do {                                                                                \
    cudaError_t e__ = (call);                                                         \
    if (e__ != cudaSuccess)                                                           \
      return tatt_set_error("%s failed: %s", #call, cudaGetErrorString(e__));         \
  } while (0)

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// activation codes shared by several kernels
enum { ACT_NONE = 0, ACT_RELU = 1, ACT_MISH = 2 };

// mish(x) = x * tanh(softplus(x)) (model/tsrn.py:1061-1064, torch softplus threshold 20).  With e = e^x:
// tanh(ln(1+e)) = n / (n + 2), n = e (e + 2)  (no cancellation for very negative x), sigmoid(x) = e / (1 + e):
// one ex2 and one reciprocal replace tanhf(log1pf(expf(x))) (error ~1e-6; the path's tolerance is 1e-3).
__device__ __forceinline__ float mish_tanh_sp(float x, float& e_out) {
  const float e = __expf(fminf(x, 20.f));
  e_out = e;
  const float n = e * (e + 2.f);
  return __fdividef(n, n + 2.f);
}
__device__ __forceinline__ float mish_f(float x) {
  float e;
  const float th = mish_tanh_sp(x, e);
  return x > 20.f ? x : x * th;
}
__device__ __forceinline__ float mish_grad(float x) {
  if (x > 20.f) return 1.f;
  float e;
  const float th = mish_tanh_sp(x, e);
  const float sg = __fdividef(e, 1.f + e);
  return th + x * (1.f - th * th) * sg;
}
__device__ __forceinline__ float act_fwd(float z, int act) {
  if (act == ACT_RELU) return fmaxf(z, 0.f);
  if (act == ACT_MISH) return mish_f(z);
  return z;
}
__device__ __forceinline__ float act_grad(float z, int act) {
  if (act == ACT_RELU) return z > 0.f ? 1.f : 0.f;
  if (act == ACT_MISH) return mish_grad(z);
  return 1.f;
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

// Philox4x32-10 counter RNG (own stream; dropout parity with torch's Philox consumption is not
// attainable -- SURVEY 7 "Dropout parity" -- so masks are ours, reproducible from (seed, offset)).
__device__ __forceinline__ uint4 philox4x32_10(uint2 key, uint4 ctr) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
// uniform in [0,1) for element `idx` of dropout site (seed, offset)
__device__ __forceinline__ float4 philox_uniform4(unsigned long long seed, unsigned long long offset,
                                                 unsigned long long idx4) {
  uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  uint4 ctr = make_uint4((uint32_t)idx4, (uint32_t)(idx4 >> 32), (uint32_t)offset, (uint32_t)(offset >> 32));
  uint4 r = philox4x32_10(key, ctr);
  const float s = 2.3283064365386963e-10f;  // 2^-32
  return make_float4(r.x * s, r.y * s, r.z * s, r.w * s);
}
__device__ __forceinline__ float philox_uniform1(unsigned long long seed, unsigned long long offset,
                                                unsigned long long idx) {
  float4 u = philox_uniform4(seed, offset, idx >> 2);
  int k = (int)(idx & 3);
  return k == 0 ? u.x : (k == 1 ? u.y : (k == 2 ? u.z : u.w));
}
