"""Developer tool: precision of the discriminator's input-gradient chain (the GAN part of d loss / d generated)
on the exact-fp32 path, layer by layer, against float64 autograd on the oracle."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from advoc_b200 import _native as N
from advoc_b200 import nets
from advoc_b200.train import TrainEngine
from oracle import nets_torch as O


def rel(a, b):
  a, b = a.detach().double().cpu(), b.detach().double().cpu()
  return float((a - b).norm() / b.norm().clamp_min(1e-300))


small = len(sys.argv) > 1 and sys.argv[1] == 'small'
ospec = O.SMALL if small else O.REGULAR
P = O.init_params(ospec, seed=0)
g = torch.Generator().manual_seed(12)
for k in P:
  if k.endswith('/bias'):
    P[k] = torch.randn(P[k].shape, generator=g) * 0.05
spec = nets.GenSpec(32, 5, (5, 4)) if small else nets.GenSpec(64, 8, (8, 7, 6))
eng = TrainEngine(spec, ospec.ndf, {k: v.cuda() for k, v in P.items()}, 1, math=N.MATH_FP32, l1_weight=0.0,
                  use_graphs=False)
target = torch.randn(1, 256, 513, 1, generator=g).abs() * 0.1
x = target + torch.randn(1, 256, 513, 1, generator=g) * 0.02
eng.g_step(x.cuda(), target.cuda(), dropout=None, apply=False)
torch.cuda.synchronize()
for dt in (torch.float64, torch.float32):
  Pd = {n: t.clone().to(dt) for n, t in P.items()}
  gen = O.generator(Pd, x.to(dt), ospec).detach().requires_grad_(True)
  p_fake, layers = O.discriminator(Pd, x.to(dt), gen, return_layers=True)
  for l in layers:
    l.retain_grad()
  loss = torch.mean(-torch.log(p_fake + O.EPS))
  loss.backward()
  if dt == torch.float64:
    ref_gen = gen.grad
    ref_pre = []
    for i, l in enumerate(layers):
      d = l.grad * (l * (1 - l) if i == 4 else torch.where(l > 0, torch.ones_like(l), torch.full_like(l, 0.2)))
      ref_pre.append(d)
    ref_fwd = [l.detach() for l in layers]
  else:
    print('cpu32: d loss_GAN / d generated vs f64: %.2e' % rel(gen.grad, ref_gen))
print('gpu  : d loss_GAN / d generated vs f64: %.2e   (|.| %.3e)' % (rel(eng.g_out, ref_gen), float(ref_gen.norm())))
for i in range(5):
  print('layer_%d  forward act %.2e   pre-activation gradient dz %.2e' % (i + 1, rel(eng.Df.act[i], ref_fwd[i]), rel(eng.dz[i], ref_pre[i])))
