// Shared device/host helpers for the mvptr_b200 kernels (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mvptr_b200.h"

namespace mvptr {

// ---- error plumbing: C-ABI entry points return 0 / negative code, never throw ----
void set_last_error(const char* fmt, ...);
#define MVPTR_FAIL(code, ...)            \
  do {                                   \
    ::mvptr::set_last_error(__VA_ARGS__); \
    return (code);                       \
  } while (0)
// every kernel launch of the library passes through here (or through gemm's launch()): the exact
// count backs bench.py's `gpu_launches` (mvptr_launch_count)
extern unsigned long long g_launch_count;
#define MVPTR_CHECK_LAUNCH(name)                                                          \
  do {                                                                                    \
    ++::mvptr::g_launch_count;                                                            \
    cudaError_t e__ = cudaGetLastError();                                                 \
    if (e__ != cudaSuccess) MVPTR_FAIL(MVPTR_ERR_CUDA, "%s: %s", name, cudaGetErrorString(e__)); \
  } while (0)

struct ProfScope {
  ProfScope(const char* name, double work, cudaStream_t s);
  ~ProfScope();
  bool active_;
  cudaStream_t stream_;
  int index_ = -1;
};
#define MVPTR_PROF(name, work, stream) ::mvptr::ProfScope prof_scope__(name, work, (cudaStream_t)(stream))

typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf162;

constexpr int kNumSMs = 148;

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

// 8 bf16 <-> 8 float through ONE 16-byte access.  The carrier is a trivially copyable POD of four 32-bit
// words: a struct of __nv_bfloat162 members has user-provided copy operations, which made every
// `*reinterpret_cast<const bf16x8*>(p)` a member-wise copy -- four LDG.32 / STG.32 per lane instead of one
// LDG.128 / STG.128 (seen in the SASS of every row kernel).
struct alignas(16) bf16x8 {
  uint32_t w[4];
};
__device__ __forceinline__ void unpack8(const uint4& p, float* f) {
  const uint4 v = p;  // ONE 16-byte load when p refers to memory (member-wise reads were emitted as four LDG.32)
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);             // low half = element 2i
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);  // high half = element 2i + 1
  }
}
__device__ __forceinline__ void unpack8(const bf16x8& p, float* f) {
  unpack8(*reinterpret_cast<const uint4*>(&p), f);
}
__device__ __forceinline__ bf16x8 pack8(const float* f) {
  bf16x8 p;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bf162 t = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    p.w[i] = *reinterpret_cast<const uint32_t*>(&t);
  }
  return p;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  bf162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

// erf-GELU (reference modeling_bert.py:142-148) and its derivative.  erf through Abramowitz-Stegun
// 7.1.26 (|err| <= 1.5e-7, far below bf16 output rounding): 2 MUFU + ~10 FMA instead of erff's
// branchy ~25-instruction path -- the FFN epilogues run this per element next to the tensor pipe.
__device__ __forceinline__ float rcp_approx(float x) {  // one MUFU.RCP, no range fix-ups (x in [1, 1e19) here)
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_approx(float x) {  // one MUFU.EX2, no denormal fix-ups
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void erf_gauss(float x, float& erf_v, float& gauss) {
  const float t = rcp_approx(fmaf(0.3275911f * 0.70710678118654752f, fabsf(x), 1.0f));
  gauss = ex2_approx(x * x * -0.72134752044448170f);  // exp(-x^2/2) = 2^(-x^2 log2(e) / 2)
  float p = fmaf(t, 1.061405429f, -1.453152027f);
  p = fmaf(t, p, 1.421413741f);
  p = fmaf(t, p, -0.284496736f);
  p = fmaf(t, p, 0.254829592f);
  erf_v = copysignf(fmaf(-p * t, gauss, 1.0f), x);
}
__device__ __forceinline__ float gelu_erf(float x) {
  float e, g;
  erf_gauss(x, e, g);
  const float hx = 0.5f * x;
  return fmaf(hx, e, hx);
}
// gelu(x) and gelu'(x) from ONE erf / gaussian evaluation (the forward epilogue that saves gelu' for backward)
__device__ __forceinline__ void gelu_erf_both(float x, float& act, float& grad) {
  float e, g;
  erf_gauss(x, e, g);
  const float hx = 0.5f * x;
  act = fmaf(hx, e, hx);
  grad = fmaf(x * 0.39894228040143268f, g, fmaf(0.5f, e, 0.5f));
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  float e, g;
  erf_gauss(x, e, g);
  return fmaf(x * 0.39894228040143268f, g, 0.5f * (1.0f + e));
}

// Stateless per-element dropout RNG: a 32-bit integer hash of (site seed, element
// index).  Layout independent, so forward epilogues and backward kernels regenerate
// the same keep mask without storing it.
__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}
// Dropout epoch: a per-translation-unit device word mixed into every dropout hash.  A CUDA-graph
// replay cannot change kernel arguments (the per-site seeds are baked in at capture), so the graph
// starts with a copy of a pinned host counter into this word (mvptr_set_dropout_epoch) and every
// replay draws fresh masks.  Outside graphs it stays 0 and the host varies the seeds instead.
static __device__ uint32_t g_dropout_epoch = 0;
#define MVPTR_DEFINE_EPOCH_SETTER(fn)                                                                  \
  extern "C" int fn(const uint32_t* src, void* stream) {                                               \
    cudaError_t e = cudaMemcpyToSymbolAsync(::mvptr::g_dropout_epoch, src, sizeof(uint32_t), 0, cudaMemcpyDefault, \
                                            (cudaStream_t)stream);                                     \
    if (e != cudaSuccess) MVPTR_FAIL(MVPTR_ERR_CUDA, #fn ": %s", cudaGetErrorString(e));                \
    return 0;                                                                                          \
  }                                                                                                    \
  /* device address of this translation unit's epoch word (mvptr_step_params writes all of them) */   \
  extern "C" int fn##_addr(uint32_t** out) {                                                           \
    void* ptr = nullptr;                                                                               \
    cudaError_t e = cudaGetSymbolAddress(&ptr, ::mvptr::g_dropout_epoch);                              \
    if (e != cudaSuccess) MVPTR_FAIL(MVPTR_ERR_CUDA, #fn "_addr: %s", cudaGetErrorString(e));          \
    *out = static_cast<uint32_t*>(ptr);                                                                \
    return 0;                                                                                          \
  }

// One 32-bit hash serves TWO neighbouring elements (idx, idx^1): 16 random bits each, compared
// against keep_thr >> 16 (p_keep resolution 1.5e-5).  Halves the integer work of the dropout sites.
// Kernels mix the epoch into their site seed ONCE (site_seed) and pass the result to the helpers below:
// reading the epoch word inside dropout_bits put a dependent global load in front of every hash.
__device__ __forceinline__ uint32_t site_seed(uint32_t seed) { return seed + g_dropout_epoch * 0x85EBCA6Bu; }
__device__ __forceinline__ uint32_t dropout_bits(uint32_t seed, uint32_t idx) {
  return hash32((idx >> 1) * 0x9E3779B9u + seed);
}
__device__ __forceinline__ bool dropout_keep(uint32_t seed, uint32_t idx, uint32_t keep_thr) {
  // keep_thr = floor(keep_prob * 2^32)
  const uint32_t h = dropout_bits(seed, idx);
  return ((idx & 1u) ? (h >> 16) : (h & 0xffffu)) < (keep_thr >> 16);
}
// idx_even must be even: keep flags of elements idx_even and idx_even + 1 from one hash
__device__ __forceinline__ void dropout_pair(uint32_t seed, uint32_t idx_even, uint32_t keep_thr, bool& k0, bool& k1) {
  const uint32_t h = dropout_bits(seed, idx_even);
  const uint32_t t = keep_thr >> 16;
  k0 = (h & 0xffffu) < t;
  k1 = (h >> 16) < t;
}
// v[0..8) *= keep ? inv_keep : 0 for 8 consecutive elements starting at an index that is a multiple of 8
__device__ __forceinline__ void dropout8(float* v, uint32_t seed, uint32_t base8, uint32_t keep_thr, float inv_keep) {
#pragma unroll
  for (int j = 0; j < 8; j += 2) {
    bool k0, k1;
    dropout_pair(seed, base8 + j, keep_thr, k0, k1);
    v[j] = k0 ? v[j] * inv_keep : 0.f;
    v[j + 1] = k1 ? v[j + 1] * inv_keep : 0.f;
  }
}
static inline uint32_t keep_threshold(float p_drop) {
  double kp = 1.0 - (double)p_drop;
  if (kp >= 1.0) return 0xffffffffu;
  return ((uint32_t)(kp * 65536.0 + 0.5)) << 16;  // 16-bit resolution (see dropout_keep)
}

}  // namespace mvptr
