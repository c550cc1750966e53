"""Developer check of the tcgen05 conv path: prints rel-L2 vs the CPU oracle per case, never
asserts (so one call shows every failure).  Run on the GPU box."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from advoc_b200 import _native as N  # noqa: E402
from advoc_b200 import nets  # noqa: E402
from oracle import nets_torch as O  # noqa: E402


def tf32(x):
  i = x.contiguous().view(torch.int32)
  return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def rel(a, b):
  a, b = a.detach().double().cpu(), b.detach().double().cpu()
  return float((a - b).norm() / b.norm().clamp_min(1e-300))


def run_case(kind, B, H, W, Cin, Cout, sh, sw, mode='same', math=N.MATH_AUTO):
  g = torch.Generator().manual_seed(H * 131 + W * 7 + Cin + Cout)
  if kind == 'conv':
    x = tf32(torch.randn(B, H, W, Cin, generator=g))
    k = torch.randn(4, 4, Cin, Cout, generator=g) * 0.05
    b = torch.randn(Cout, generator=g) * 0.1
    if mode == 'same':
      ref = O.conv_same(x, k, b, (sh, sw))
      ho, pt, _ = nets.same_pads(H, 4, sh)
      wo, pl, _ = nets.same_pads(W, 4, sw)
    else:
      ref = O.discrim_conv(x, k, b, sh)
      ho, wo, pt, pl = (H + 2 - 4) // sh + 1, (W + 2 - 4) // sw + 1, 1, 1
    y = torch.full((B, ho, wo, Cout), float('nan'), device='cuda')
    L = nets._Conv('t', 'conv', nets._desc(B, H, W, Cin, Cout, sh, sw, pt, pl, ho, wo, math))
    bd, xd, kd = b.cuda(), x.cuda(), k.cuda()
    wp = nets._pack_for_tc(L, kd, Cin)
    ep = nets._epilogue(bd, y, Cout, 0, N.ACT_NONE)
    L.run(xd, Cin, kd if wp is None else wp, ep)
  else:
    x = tf32(torch.relu(torch.randn(B, H, W, Cin, generator=g)))
    k = torch.randn(4, 4, Cout, Cin, generator=g) * 0.05
    b = torch.randn(Cout, generator=g) * 0.1
    ref = O.deconv_same(x, k, b, (sh, sw))[:, :, :-1, :]
    y = torch.full((B, H * sh, sw * W - 1, Cout), float('nan'), device='cuda')
    L = nets._Conv('t', 'deconv', nets._desc(B, H * sh, sw * W, Cout, Cin, sh, sw, 1, 1, H, W, math))
    bd, xd, kd = b.cuda(), x.cuda(), k.cuda()
    wp = nets._pack_for_tc(L, kd, Cin)
    ep = nets._epilogue(bd, y, Cout, 0, N.ACT_NONE, store_w=sw * W - 1)
    L.run(xd, Cin, kd if wp is None else wp, ep)
  torch.cuda.synchronize()
  flags = N.debug_flags()
  nan = int(torch.isnan(y).sum())
  print('%-6s B%d %dx%d Cin%d Cout%d s%d%d %s tc=%s rel=%.3e nan=%d dbg=%d' %
        (kind, B, H, W, Cin, Cout, sh, sw, mode, wp is not None, rel(torch.nan_to_num(y), ref), nan,
         flags), flush=True)
  return flags


if __name__ == '__main__':
  torch.cuda.set_device(0)
  print('arch', N.device_arch(), flush=True)
  cases = [
      ('conv', 1, 8, 16, 32, 32, 1, 1, 'valid'),     # smallest: 1 k-block per tap, BN=32, s1
      ('conv', 1, 16, 16, 32, 64, 2, 2, 'same'),     # BN=64, stride 2 even sizes
      ('conv', 2, 16, 33, 64, 128, 2, 2, 'same'),    # BN=128, odd width, 2 k-blocks
      ('conv', 2, 1, 5, 64, 32, 1, 2, 'same'),       # H stride 1
      ('conv', 2, 32, 65, 32, 64, 2, 2, 'valid'),    # PatchGAN pad-1 VALID
      ('conv', 2, 16, 32, 64, 96, 1, 1, 'valid'),
      ('conv', 3, 64, 129, 64, 128, 2, 2, 'same'),   # several tiles crossing rows and images
      ('deconv', 1, 4, 8, 32, 32, 2, 2, 'same'),
      ('deconv', 2, 8, 17, 64, 32, 2, 2, 'same'),
      ('deconv', 2, 1, 3, 32, 64, 1, 2, 'same'),
      ('deconv', 2, 16, 33, 512, 128, 2, 2, 'same'),
      ('deconv', 3, 32, 65, 256, 64, 2, 2, 'same'),
  ]
  for c in cases:
    if run_case(*c) != 0:
      print('pipeline timeout: stopping', flush=True)
      break
