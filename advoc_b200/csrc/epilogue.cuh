// Output-side contract shared by every convolution kernel (see advoc_epilogue in the header).
#pragma once
#include "common.cuh"

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
};

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
    const bool keep = e.mask ? (__ldg(e.mask + idx) != 0) : dropout_keep(e.seed, idx, e.keep_prob);
    scale = keep ? 1.f / e.keep_prob : 0.f;
  }
  float y0 = apply_act(v, e.act0, e.alpha) * scale;
  if (e.gate) y0 *= gate_factor(e, pix, n);
  float* dst = e.out0 + pix * e.ld0 + e.coff0 + n;
  if (e.accumulate) y0 += *dst;
  if (e.round) y0 = round_tf32(y0);
  *dst = y0;
  if (e.out1) {
    float y1 = apply_act(v, e.act1, e.alpha) * scale;
    if (e.round) y1 = round_tf32(y1);
    e.out1[pix * e.ld1 + e.coff1 + n] = y1;
  }
}

}  // namespace advoc
