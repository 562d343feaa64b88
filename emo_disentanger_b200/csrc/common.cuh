// Shared device/host helpers for libemo_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/emo_b200.h"

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------------------------------------
// error reporting (thread-local message, never throws / exits across the ABI)
// ---------------------------------------------------------------------------------------------
void emo_set_error(const char* fmt, ...);

#define EMO_CHECK_CUDA(expr)                                                                    \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      emo_set_error("%s:%d CUDA error %d (%s) in %s", __FILE__, __LINE__, (int)_e,              \
                    cudaGetErrorString(_e), #expr);                                             \
      return EMO_ERR_CUDA;                                                                      \
    }                                                                                           \
  } while (0)

#define EMO_REQUIRE(cond, ...)                                                                  \
  do {                                                                                          \
    if (!(cond)) {                                                                              \
      emo_set_error(__VA_ARGS__);                                                               \
      return EMO_ERR_ARG;                                                                       \
    }                                                                                           \
  } while (0)

#define EMO_LAUNCH_CHECK() EMO_CHECK_CUDA(cudaGetLastError())

// ---------------------------------------------------------------------------------------------
// dtype helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(bf16 x) { return __bfloat162float(x); }
template <typename T> __device__ __forceinline__ T from_f(float x);
template <> __device__ __forceinline__ float from_f<float>(float x) { return x; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float x) { return __float2bfloat16_rn(x); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void unpack_bf16x2(uint32_t p, float& lo, float& hi) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&p);
  lo = __low2float(v);
  hi = __high2float(v);
}

// 8-element (bf16: 16 B) / 4-element (fp32: 16 B) vector access used by the memory-bound kernels.
template <typename T> struct Vec;  // VEC elements per 16 bytes
template <> struct Vec<float> {
  static constexpr int N = 4;
  float v[4];
  __device__ __forceinline__ void load(const float* p) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Vec<bf16> {
  static constexpr int N = 8;
  float v[8];
  __device__ __forceinline__ void load(const bf16* p) {
    uint4 t = *reinterpret_cast<const uint4*>(p);
    unpack_bf16x2(t.x, v[0], v[1]); unpack_bf16x2(t.y, v[2], v[3]);
    unpack_bf16x2(t.z, v[4], v[5]); unpack_bf16x2(t.w, v[6], v[7]);
  }
  __device__ __forceinline__ void store(bf16* p) const {
    uint4 t;
    t.x = pack_bf16x2(v[0], v[1]); t.y = pack_bf16x2(v[2], v[3]);
    t.z = pack_bf16x2(v[4], v[5]); t.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = t;
  }
};

// ---------------------------------------------------------------------------------------------
// counter-based dropout RNG: one 32-bit hash per PAIR of elements, 16 bits per decision.
// keep(element e) <=> 16-bit lane of hash(seed, e/2) >= thr,  thr = round(p * 65536).
// The same (seed, element index) is re-evaluated in backward; nothing is stored.
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t emo_mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
  return x;
}
__host__ __device__ __forceinline__ uint32_t emo_drop_thr(float p) {
  return (uint32_t)(p * 65536.0f + 0.5f);
}
// hash for the pair containing element index e (e even -> low 16 bits, e odd -> high 16 bits).
// Two-multiply finaliser on (pair * golden + key): the mask costs issue slots in the GEMM epilogues and in the
// issue-bound LN backward (ncu: the previous three-multiply mix was ~30 % of the FFN1 GEMM's instructions); drop
// rate, adjacent / row / column correlations are at the noise level of the stronger mix (checked on 8 M elements).
__host__ __device__ __forceinline__ uint32_t emo_drop_key(uint64_t seed, uint32_t pair_hi) {
  return (uint32_t)seed ^ emo_mix32(pair_hi + (uint32_t)(seed >> 32) + 0x7f4a7c15u);
}
__host__ __device__ __forceinline__ uint32_t emo_drop_mix(uint32_t pair_lo, uint32_t key) {
  uint32_t h = pair_lo * 0x9E3779B9u + key;
  h ^= h >> 15;
  h *= 0x2C1B3C6Du;
  h ^= h >> 13;
  return h;
}
__device__ __forceinline__ uint32_t emo_drop_hash(uint64_t seed, uint64_t e) {
  uint64_t pair = e >> 1;
  return emo_drop_mix((uint32_t)pair, emo_drop_key(seed, (uint32_t)(pair >> 32)));
}
__device__ __forceinline__ bool emo_drop_keep(uint64_t seed, uint64_t e, uint32_t thr) {
  uint32_t h = emo_drop_hash(seed, e);
  uint32_t bits = (e & 1) ? (h >> 16) : (h & 0xffffu);
  return bits >= thr;
}

// ---------------------------------------------------------------------------------------------
// warp / block reductions
// ---------------------------------------------------------------------------------------------
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

__device__ __forceinline__ float gelu_new_f(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  return 0.5f * x * (1.0f + tanhf(k0 * (x + k1 * x * x * x)));
}
__device__ __forceinline__ float gelu_new_grad_f(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float t = tanhf(k0 * (x + k1 * x * x * x));
  float dt = (1.0f - t * t) * k0 * (1.0f + 3.0f * k1 * x * x);
  return 0.5f * (1.0f + t) + 0.5f * x * dt;
}

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (decode step): a kernel launched with the attribute may start while its
// predecessor drains; it must not touch the predecessor's outputs before pdl_wait().  Both instructions are
// no-ops for a normally launched kernel.  emo_set_pdl(1) switches the decode-step kernels to these launches.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
int emo_pdl_enabled();
template <typename... KArgs, typename... Args>
static inline cudaError_t emo_launch_dep(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = emo_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

static inline int emo_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}
