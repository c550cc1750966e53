"""Developer tool: where does the regular generator's backward pass leave the oracle?  Compares the
activation gradients the engine stores (gCat[k]) with autograd on the oracle, layer by layer."""
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


math = N.MATH_FP32 if (len(sys.argv) < 2 or sys.argv[1] == 'fp32') else N.MATH_AUTO
P = O.init_params(O.REGULAR, seed=0)
g = torch.Generator().manual_seed(12)
for k in P:
  if k.endswith('/bias'):
    P[k] = torch.randn(P[k].shape, generator=g) * 0.05
spec = nets.GenSpec(64, 8, (8, 7, 6))
eng = TrainEngine(spec, 64, {k: v.cuda() for k, v in P.items()}, 1, math=math)
target = torch.randn(1, 256, 513, 1, generator=g).abs() * 0.1
x = target + torch.randn(1, 256, 513, 1, generator=g) * 0.02
if len(sys.argv) > 2 and sys.argv[2] == 'dfirst':
  eng.d_step(x.cuda(), target.cuda(), dropout=None, apply=False)
eng.g_step(x.cuda(), target.cuda(), dropout=None, apply=False)
torch.cuda.synchronize()


def oracle(dtype):
  Pg = {n: t.clone().to(dtype).requires_grad_(n.startswith('generator')) for n, t in P.items()}
  gen, layers = O.generator(Pg, x.to(dtype), O.REGULAR, return_layers=True)
  for l in layers:
    l.retain_grad()
  p_fake = O.discriminator(Pg, x.to(dtype), gen)
  loss = torch.mean(-torch.log(p_fake + O.EPS)) + 10.0 * torch.mean(torch.abs(target.to(dtype) - gen))
  loss.backward()
  return Pg, layers


Pg64, layers64 = oracle(torch.float64)
Pg, layers = oracle(torch.float32)
print('CPU f32 autograd vs CPU f64 autograd, GPU vs CPU f64:')
for nme in O.g_names(P):
  print('%-50s cpu32 %.2e  gpu %.2e' % (nme, rel(Pg[nme].grad, Pg64[nme].grad), rel(eng.flat.G[nme], Pg64[nme].grad)))
print('d loss / d generated  cpu32 %.2e gpu %.2e' % (rel(layers[-1].grad, layers64[-1].grad), rel(eng.g_out, layers64[-1].grad)))
n = 8
print('d loss / d generated:', rel(eng.g_out, layers[-1].grad))
for k in range(1, n + 1):
  Dk = eng.G.Dk[k]
  gc = eng.gCat[k]
  enc_total = layers[k - 1].grad            # d loss / d encoder_k output (skip + deeper path)
  if k < n:
    dec = layers[n + (n - 1 - k)].grad[:, :, :-1, :]      # d loss / d decoder_{k+1} output (cropped part)
    print('k=%d  gCat[:Dk] vs d/d decoder_%d out: %.2e   gCat[Dk:] (after accumulation) vs d/d encoder_%d out: %.2e'
          % (k, k + 1, rel(gc[..., :Dk], dec), k, rel(gc[..., Dk:], enc_total)))
  else:
    print('k=%d  gCat vs d/d encoder_%d out: %.2e' % (k, k, rel(gc, enc_total)))
for nme in O.g_names(P):
  print('%-50s %.2e  |g| %.3e' % (nme, rel(eng.flat.G[nme], Pg[nme].grad), float(Pg[nme].grad.norm())))
