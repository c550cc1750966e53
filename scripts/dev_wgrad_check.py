"""Developer check of the tcgen05 filter gradient against torch autograd (CPU).  GPU box only."""
import ctypes as C
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


def case(B, H, W, Cin, Cout, sh, sw, mode, math):
  g = torch.Generator().manual_seed(H * 31 + W + Cin * 7 + Cout)
  x = tf32(torch.randn(B, H, W, Cin, generator=g))
  k = (torch.randn(4, 4, Cin, Cout, generator=g) * 0.05).requires_grad_(True)
  if mode == 'same':
    y = O.conv_same(x, k, None, (sh, sw))
    ho, pt, _ = nets.same_pads(H, 4, sh)
    wo, pl, _ = nets.same_pads(W, 4, sw)
  else:
    y = O.discrim_conv(x, k, None, sh)
    ho, wo, pt, pl = (H + 2 - 4) // sh + 1, (W + 2 - 4) // sw + 1, 1, 1
  dy = tf32(torch.randn(y.shape, generator=g))
  (ref,) = torch.autograd.grad((y * dy).sum(), [k])
  d = nets._desc(B, H, W, Cin, Cout, sh, sw, pt, pl, ho, wo, math)
  dw = torch.zeros(4, 4, Cin, Cout, device='cuda')
  xd, dyd = x.cuda(), dy.cuda()
  N.call('advoc_conv2d_wgrad', C.byref(d), C.c_void_p(xd.data_ptr()), Cin, C.c_void_p(dyd.data_ptr()), Cout,
         C.c_void_p(dw.data_ptr()), None)
  torch.cuda.synchronize()
  if rel(dw, ref) > 1e-2:
    print('   dw.flat[:8]', dw.reshape(-1)[:8].tolist())
    print('   |dw|=%.4e |ref|=%.4e dw[0,0,0,:4]=%s ref=%s' % (float(dw.norm()), float(ref.norm()), dw[0,0,0,:4].tolist(), ref[0,0,0,:4].tolist()))
    a = dw.cpu().reshape(16, -1); b = ref.reshape(16, -1)
    print('   per-tap rel', ['%.2f' % (float((a[t]-b[t]).norm()/b[t].norm())) for t in range(16)])
  print('wgrad B%d %dx%d Cin%d Cout%d s%d%d %s math%d rel=%.3e dbg=%d' %
        (B, H, W, Cin, Cout, sh, sw, mode, math, rel(dw, ref), N.debug_flags()), flush=True)


if __name__ == '__main__':
  torch.cuda.set_device(0)
  for math in (N.MATH_AUTO,):
    case(1, 8, 16, 32, 32, 1, 1, 'valid', math)
    case(2, 16, 33, 64, 128, 2, 2, 'same', math)
    case(2, 16, 33, 128, 64, 2, 2, 'same', math)
    case(2, 32, 65, 32, 64, 2, 2, 'valid', math)
    case(3, 16, 32, 256, 96, 1, 1, 'valid', math)
    case(2, 64, 129, 160, 32, 2, 2, 'same', math)
