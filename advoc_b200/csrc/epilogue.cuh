// Output-side contract shared by every convolution kernel (see advoc_epilogue in the header).
#pragma once
#include "common.cuh"

#include <cuda_fp16.h>

namespace advoc {

struct EpiDev {
  const float* bias;
  float* out0;
  float* out1;
  const uint8_t* mask;
  uint64_t seed;
  int act0, act1;
  int ld0, coff0, ld1, coff1;
  int Hs, Ws;       // stored spatial extent (Ws = store_w <= produced width)
  int Cout;
  float alpha;
  float keep_prob;  // 1 = no dropout
  int round;        // round stored values to tf32
  // backward-pass extensions (out0 only)
  const float* gate;
  int ldg, coffg, gate_act, gate_split;
  float gscale0, gscale1;
  int accumulate;
  // dropout counter in device memory (CUDA-graph replays draw fresh masks): see epi_seed()
  const unsigned long long* seed_ptr;
  int h0, h1;       // destination holds fp16 elements (ld / coff stay in elements); forward pass only
  int row_pad0;     // out0 rows are Ws + row_pad0 pixels apart (thin-input convolution only)
};

// Seed of the counter-based dropout generator: the by-value seed, or -- when the caller keeps a
// step counter in device memory -- counter * 0x9E3779B1 + seed (seed = the layer's salt), which is
// what the host computes for the eager path (nets.Generator.forward), so both paths draw the same
// masks for the same step.
__device__ __forceinline__ uint64_t epi_seed(const EpiDev& e) {
  return e.seed_ptr ? __ldg(e.seed_ptr) * 0x9E3779B1ull + e.seed : e.seed;
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
// four consecutive channels to an fp32 or fp16 destination (elem = element index of the first one)
__device__ __forceinline__ void store4(float* base, int is_half, size_t elem, const float (&y)[4]) {
  if (is_half) {
    *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(base) + elem) =
        make_uint2(pack_half2(y[0], y[1]), pack_half2(y[2], y[3]));
  } else {
    *reinterpret_cast<float4*>(base + elem) = make_float4(y[0], y[1], y[2], y[3]);
  }
}

__device__ __forceinline__ float gate_factor(const EpiDev& e, size_t pix, int n) {
  const float g = __ldg(e.gate + pix * e.ldg + e.coffg + n);
  const float d = g > 0.f ? 1.f : (e.gate_act == ADVOC_ACT_LRELU ? e.alpha : 0.f);
  return d * (n < e.gate_split ? e.gscale0 : e.gscale1);
}

// Validates the user epilogue against the produced tensor [N, Hs, Wfull, Cout] and lowers it.
int lower_epilogue(const advoc_epilogue* ep, int Hs, int Wfull, int Cout, EpiDev* out);

// One element.  pix = linear index of the stored pixel ((b*Hs + h)*Ws + w), n = channel.
__device__ __forceinline__ void epi_store(const EpiDev& e, size_t pix, int n, float acc) {
  float v = acc + (e.bias ? __ldg(e.bias + n) : 0.f);
  float scale = 1.f;
  if (e.keep_prob < 1.f) {
    const size_t idx = pix * e.Cout + n;
    const bool keep = e.mask ? (__ldg(e.mask + idx) != 0) : dropout_keep(epi_seed(e), idx, e.keep_prob);
    scale = keep ? 1.f / e.keep_prob : 0.f;
  }
  const bool lin = act_is_linear(e.act0) && act_is_linear(e.act1);   // real branch: see ActLin
  float y0;
  if (lin) y0 = apply_lin(v, act_linear(e.act0, e.alpha)) * scale;
  else y0 = apply_act(v, e.act0, e.alpha) * scale;
  if (e.gate) y0 *= gate_factor(e, pix, n);
  if (e.h0) {
    reinterpret_cast<__half*>(e.out0)[pix * e.ld0 + e.coff0 + n] = __float2half_rn(y0);
  } else {
    float* dst = e.out0 + pix * e.ld0 + e.coff0 + n;
    if (e.accumulate) y0 += *dst;
    if (e.round) y0 = round_tf32(y0);
    *dst = y0;
  }
  if (e.out1) {
    float y1;
    if (lin) y1 = apply_lin(v, act_linear(e.act1, e.alpha)) * scale;
    else y1 = apply_act(v, e.act1, e.alpha) * scale;
    if (e.h1) {
      reinterpret_cast<__half*>(e.out1)[pix * e.ld1 + e.coff1 + n] = __float2half_rn(y1);
    } else {
      if (e.round) y1 = round_tf32(y1);
      e.out1[pix * e.ld1 + e.coff1 + n] = y1;
    }
  }
}

// Four consecutive channels of one stored pixel (the vector form of epi_store); `acc` already holds the bias,
// `seed` = epi_seed(e) and `inv` = 1 / keep_prob come from the caller (loop-invariant: a kernel that calls this
// 32 times per pixel row should not reload the seed word and re-divide every time).
__device__ __forceinline__ void epi_store_vec4_core(const EpiDev& e, size_t pix, int n, float4 acc, uint64_t seed,
                                                    float inv);

__device__ __forceinline__ void epi_store_vec4(const EpiDev& e, size_t pix, int n, float4 acc) {
  if (e.bias) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(e.bias + n));
    acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
  }
  epi_store_vec4_core(e, pix, n, acc, e.keep_prob < 1.f ? epi_seed(e) : 0ull, 1.f / e.keep_prob);
}

__device__ __forceinline__ void epi_store_vec4_core(const EpiDev& e, size_t pix, int n, float4 acc, uint64_t seed,
                                                    float inv) {
  float v[4] = {acc.x, acc.y, acc.z, acc.w};
  float sc[4] = {1.f, 1.f, 1.f, 1.f};
  if (e.keep_prob < 1.f) {
    const size_t idx = pix * e.Cout + n;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool keep = e.mask ? (__ldg(e.mask + idx + j) != 0) : dropout_keep(seed, idx + j, e.keep_prob);
      sc[j] = keep ? inv : 0.f;
    }
  }
  const bool lin = act_is_linear(e.act0) && act_is_linear(e.act1);
  float y[4];
  if (lin) {
    const ActLin a0 = act_linear(e.act0, e.alpha);
#pragma unroll
    for (int j = 0; j < 4; ++j) y[j] = apply_lin(v[j], a0) * sc[j];
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) y[j] = apply_act(v[j], e.act0, e.alpha) * sc[j];
  }
  if (e.gate) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(e.gate + pix * e.ldg + e.coffg + n));
    const float gv[4] = {g.x, g.y, g.z, g.w};
    const float neg = e.gate_act == ADVOC_ACT_LRELU ? e.alpha : 0.f;
    const float s = n < e.gate_split ? e.gscale0 : e.gscale1;  // split is a multiple of 4
#pragma unroll
    for (int j = 0; j < 4; ++j) y[j] *= (gv[j] > 0.f ? 1.f : neg) * s;
  }
  if (e.accumulate) {   // fp32 destinations only (lower_epilogue)
    const float4 o = *reinterpret_cast<const float4*>(e.out0 + pix * e.ld0 + e.coff0 + n);
    y[0] += o.x; y[1] += o.y; y[2] += o.z; y[3] += o.w;
  }
  if (e.round && !e.h0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) y[j] = round_tf32(y[j]);
  }
  store4(e.out0, e.h0, pix * e.ld0 + e.coff0 + n, y);
  if (e.out1) {
    if (lin) {
      const ActLin a1 = act_linear(e.act1, e.alpha);
#pragma unroll
      for (int j = 0; j < 4; ++j) y[j] = apply_lin(v[j], a1) * sc[j];
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) y[j] = apply_act(v[j], e.act1, e.alpha) * sc[j];
    }
    if (e.round && !e.h1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) y[j] = round_tf32(y[j]);
    }
    store4(e.out1, e.h1, pix * e.ld1 + e.coff1 + n, y);
  }
}

}  // namespace advoc
