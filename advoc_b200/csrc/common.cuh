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

// Counter-based keep/drop decision shared by every kernel that applies dropout: a 32-bit multiply-xorshift
// hash ("lowbias32" constants) of the element index, keyed by both halves of the seed.  Deterministic, stateless.
// (Round 2: the splitmix64 finaliser used before cost ~30 instructions per element -- three 64-bit multiplies --
// and made the epilogues of the dropout decoders the slowest stage of their kernels: 2000 warp instructions per
// 32-column chunk in profiles/r02K_conv_tc_small_stall_samples.txt.)
__device__ __forceinline__ bool dropout_keep(uint64_t seed, uint64_t idx, float keep_prob) {
  uint32_t x = (uint32_t)idx * 0x9E3779B1u + (uint32_t)seed;
  x ^= ((uint32_t)(idx >> 32) + (uint32_t)(seed >> 32)) * 0x85EBCA6Bu;
  x ^= x >> 16;
  x *= 0x21F0AAADu;
  x ^= x >> 15;
  x *= 0x735A2D97u;
  x ^= x >> 15;
  return (x >> 8) < (uint32_t)(keep_prob * 16777216.0f);
}

}  // namespace advoc
