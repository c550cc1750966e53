"""CPU checks of the arithmetic identities the round-2 tensor-core kernels are built on (no GPU, no library
call: the schedules are restated in numpy / torch next to the oracle's layer).

* conv_tc_merged_kernel (csrc/conv_tc.cu): a k4 s2 pad-1 transposed convolution = for every input-grid position
  (a, b) the four outputs (2a + ph, 2b + pw), each the sum over the taps kh, kw of its parity of
  x[a + dh, b + dw] . w[kh, kw] with dh = (ph + 1 - kh) / 2, dw = (pw + 1 - kw) / 2: nine shifted views, 16 (class,
  tap) pairs, the centre view reaching all four classes.
* conv_one_in_tc_kernel / wgrad_thin_tc_kernel: x = hi + lo with hi = x truncated to tf32 and lo = x - hi (truncated
  by the tensor core), w = w_hi + w_lo rounded: x_hi w_hi + x_lo w_hi + x_hi w_lo reproduces the fp32 product to ~1e-6.
* common.cuh dropout_keep: the 32-bit counter hash keeps a fraction keep_prob of the elements and decorrelates
  consecutive seeds / indices.
* conv_tc.cu conv_to_one_tc: a conv to one channel = per input pixel the tap sums T[q][tap] = <x[q], w[tap]>, then
  y[p] = sum_tap T[p * s - pad + tap][tap].
"""
import numpy as np
import torch

from oracle import nets_torch as O


def _merged_tables():
  """The issue-order tables conv_transposed_tc builds for the merged kernel (host code restated)."""
  order = [4, 0, 1, 2, 3, 5, 6, 7, 8]
  entries = []
  for s9 in order:
    dh, dw = s9 // 3 - 1, s9 % 3 - 1
    pairs = []
    for ph in range(2):
      for pw in range(2):
        for kh in range(4):
          for kw in range(4):
            if (ph + 1 - kh) % 2 or (pw + 1 - kw) % 2:
              continue
            if (ph + 1 - kh) // 2 != dh or (pw + 1 - kw) // 2 != dw:
              continue
            pairs.append((2 * ph + pw, kh * 4 + kw))
    entries.append((dh, dw, pairs))
  return entries


def test_merged_parity_class_schedule_equals_transposed_conv():
  g = torch.Generator().manual_seed(0)
  B, H, W, Cin, Cout = 2, 5, 7, 6, 4
  x = torch.randn(B, H, W, Cin, generator=g, dtype=torch.float64)
  k = torch.randn(4, 4, Cout, Cin, generator=g, dtype=torch.float64)     # HWOI
  ref = O.deconv_same(x, k, torch.zeros(Cout, dtype=torch.float64), (2, 2))
  assert tuple(ref.shape) == (B, 2 * H, 2 * W, Cout)
  entries = _merged_tables()
  assert sum(len(p) for _, _, p in entries) == 16
  assert [len(p) for _, _, p in entries] == [4, 1, 2, 1, 2, 2, 1, 2, 1]
  assert [c for c, _ in entries[0][2]] == [0, 1, 2, 3]                   # centre view: one MMA over all classes
  xp = torch.zeros(B, H + 2, W + 2, Cin, dtype=torch.float64)
  xp[:, 1:-1, 1:-1] = x
  acc = torch.zeros(4, B, H, W, Cout, dtype=torch.float64)               # four accumulators per grid position
  for dh, dw, pairs in entries:
    view = xp[:, 1 + dh:1 + dh + H, 1 + dw:1 + dw + W]                   # x[a + dh, b + dw], zero outside
    for cls, tap in pairs:
      w = k[tap // 4, tap % 4]                                          # [Cout, Cin]
      acc[cls] += torch.einsum('bhwc,oc->bhwo', view, w)
  out = torch.zeros_like(ref)
  for cls in range(4):
    out[:, (cls >> 1)::2, (cls & 1)::2] = acc[cls]
  assert float((out - ref).abs().max()) < 1e-12
  # adjacent-class grouping used by the MMA issuer: (dh = +-1, dw = 0) pairs are classes (c, c + 1)
  for dh, dw, pairs in entries:
    if len(pairs) == 2:
      adjacent = pairs[1][0] == pairs[0][0] + 1
      assert adjacent == (dw == 0)


def _trunc_tf32(x):
  return (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def _round_tf32(x):
  return ((x.view(np.uint32) + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def test_three_pass_tf32_split_reproduces_the_fp32_product():
  rng = np.random.RandomState(1)
  x = (rng.randn(4096, 16) * 3).astype(np.float32)
  w = (rng.randn(16, 64) * 0.05).astype(np.float32)
  ref = x.astype(np.float64) @ w.astype(np.float64)
  x_hi = _trunc_tf32(x)
  x_lo = _trunc_tf32(x - x_hi)                       # x - hi is exact; the tensor core truncates it
  w_hi = _round_tf32(w)
  w_lo = _round_tf32(w - w_hi)
  got = (x_hi.astype(np.float64) @ w_hi + x_lo.astype(np.float64) @ w_hi + x_hi.astype(np.float64) @ w_lo)
  one_pass = _trunc_tf32(x).astype(np.float64) @ _round_tf32(w)
  rel = np.linalg.norm(got - ref) / np.linalg.norm(ref)
  rel1 = np.linalg.norm(one_pass - ref) / np.linalg.norm(ref)
  assert rel < 2e-6, rel
  assert rel1 > 1e-4                                 # what a single tf32 pass would cost the first layer


def _dropout_keep(seed, idx, keep_prob):
  """common.cuh dropout_keep restated (uint32 arithmetic)."""
  seed = np.uint64(seed)
  idx = idx.astype(np.uint64)
  m = np.uint64(0xFFFFFFFF)
  x = ((idx & m) * np.uint64(0x9E3779B1) + (seed & m)) & m
  x ^= ((((idx >> np.uint64(32)) + (seed >> np.uint64(32))) & m) * np.uint64(0x85EBCA6B)) & m
  x ^= x >> np.uint64(16)
  x = (x * np.uint64(0x21F0AAAD)) & m
  x ^= x >> np.uint64(15)
  x = (x * np.uint64(0x735A2D97)) & m
  x ^= x >> np.uint64(15)
  return (x >> np.uint64(8)) < np.uint64(int(keep_prob * 16777216.0))


def test_dropout_hash_statistics():
  idx = np.arange(1 << 20)
  a = _dropout_keep(1, idx, 0.5)
  b = _dropout_keep(2, idx, 0.5)
  assert abs(a.mean() - 0.5) < 3e-3 and abs(b.mean() - 0.5) < 3e-3
  assert abs((a == b).mean() - 0.5) < 3e-3                       # consecutive seeds: independent masks
  assert abs((a[1:] == a[:-1]).mean() - 0.5) < 3e-3              # neighbouring elements: independent
  assert abs(_dropout_keep(7, idx, 0.8).mean() - 0.8) < 3e-3
  # seeds as the engines derive them: counter * 0x9E3779B1 + layer salt
  c = _dropout_keep((5 * 0x9E3779B1 + 3) & 0xFFFFFFFFFFFFFFFF, idx, 0.5)
  d = _dropout_keep((6 * 0x9E3779B1 + 3) & 0xFFFFFFFFFFFFFFFF, idx, 0.5)
  assert abs((c == d).mean() - 0.5) < 3e-3


def test_conv_to_one_channel_as_tap_sums_plus_gather():
  g = torch.Generator().manual_seed(2)
  B, H, W, C = 2, 9, 11, 8
  x = torch.randn(B, H, W, C, generator=g, dtype=torch.float64)
  k = torch.randn(4, 4, C, 1, generator=g, dtype=torch.float64)
  ref = O.discrim_conv(x, k, torch.zeros(1, dtype=torch.float64), 1)      # pad 1 + VALID, stride 1
  T = torch.einsum('bhwc,tc->bhwt', x, k.reshape(16, C))                  # 1x1 "convolution" to the 16 tap sums
  ho, wo = H + 2 - 4 + 1, W + 2 - 4 + 1
  y = torch.zeros(B, ho, wo, 1, dtype=torch.float64)
  for oh in range(ho):
    for ow in range(wo):
      for kh in range(4):
        for kw in range(4):
          ih, iw = oh - 1 + kh, ow - 1 + kw
          if 0 <= ih < H and 0 <= iw < W:
            y[:, oh, ow, 0] += T[:, ih, iw, kh * 4 + kw]
  assert float((y - ref).abs().max()) < 1e-12
