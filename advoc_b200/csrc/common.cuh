// Shared host/device helpers for the advoc_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/advoc_b200.h"

namespace advoc {

// ---------------------------------------------------------------------------
// error plumbing: no exception or abort ever crosses the C ABI
// ---------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int fail(advoc_status st, const char* fmt, ...);

#define ADVOC_CHECK_CUDA(expr)                                                              \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess)                                                                  \
      return ::advoc::fail(ADVOC_CUDA_ERROR, "%s failed: %s (%s:%d)", #expr,               \
                           cudaGetErrorString(_e), __FILE__, __LINE__);                     \
  } while (0)

#define ADVOC_REQUIRE(cond, st, ...)                                                        \
  do {                                                                                      \
    if (!(cond)) return ::advoc::fail(st, __VA_ARGS__);                                     \
  } while (0)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int sm_count();
int device_arch();  // major*10+minor of the current device, 0 without a device

// every kernel launch of this library bumps this counter (read with advoc_launch_count);
// bench.py reports it as `gpu_launches`
void count_launch(int n = 1);

// ---------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ float round_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// none / lrelu / relu as straight-line code: y = max(max(v, slope*v), floor).  The generic
// apply_act() switch below gets if-converted by the compiler (tanhf and expf evaluated for every
// element and then selected), which made the conv epilogues instruction-bound.
struct ActLin {
  float slope, floor;
};
__device__ __forceinline__ bool act_is_linear(int act) { return act < ADVOC_ACT_SIGMOID; }
__device__ __forceinline__ ActLin act_linear(int act, float alpha) {
  ActLin a;
  a.slope = act == ADVOC_ACT_LRELU ? alpha : 1.f;
  a.floor = act == ADVOC_ACT_RELU ? 0.f : __int_as_float(0xff800000);   // -inf
  return a;
}
__device__ __forceinline__ float apply_lin(float v, const ActLin& a) { return fmaxf(fmaxf(v, a.slope * v), a.floor); }

__device__ __forceinline__ float apply_act(float v, int act, float alpha) {
  switch (act) {
    case ADVOC_ACT_LRELU: return fmaxf(alpha * v, v);
    case ADVOC_ACT_RELU: return fmaxf(v, 0.f);
    case ADVOC_ACT_SIGMOID: return 1.f / (1.f + __expf(-v));
    case ADVOC_ACT_TANH: return tanhf(v);
    default: return v;
  }
}

// Counter-based keep/drop decision shared by every kernel that applies dropout
// (splitmix64 finaliser of seed ^ element index).  Deterministic, stateless.
__device__ __forceinline__ bool dropout_keep(uint64_t seed, uint64_t idx, float keep_prob) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (idx + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (float)(uint32_t)(z >> 40) * (1.0f / 16777216.0f) < keep_prob;
}

}  // namespace advoc
