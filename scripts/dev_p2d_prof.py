"""Developer tool: per-role cycle breakdown of every patch-kernel launch of one generator forward.
Run with ADVOC_P2D_PROFILE=1 ADVOC_P2D_VERBOSE=1 on a GPU box."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from advoc_b200 import _native as N
from advoc_b200 import nets

model = sys.argv[1] if len(sys.argv) > 1 else 'small'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
small = model == 'small'
spec = nets.GenSpec(32 if small else 64, 5 if small else 8, (5, 4) if small else (8, 7, 6))
P = nets.init_params(spec.ngf, spec.ngf, spec.n_enc, seed=0)
MATH = {'f16': N.MATH_F16, 'tf32': N.MATH_AUTO}[sys.argv[3] if len(sys.argv) > 3 else 'f16']
G = nets.Generator(spec, P, B, MATH)
G.prepare()
x = torch.rand(B, 256, 513, 1, device='cuda')
lib = N.lib()
fn = lib.advoc_p2d_profile_read
fn.restype = C.c_int
fn.argtypes = [C.c_void_p]
buf = np.zeros((512, 16), dtype=np.uint64)
orig = G._run_layer


def run(L, *a):
  torch.cuda.synchronize()
  fn(buf.ctypes.data)
  s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  s.record()
  orig(L, *a)
  e.record()
  torch.cuda.synchronize()
  fn(buf.ctypes.data)
  act = buf[buf[:, 4] > 0]
  if len(act) == 0:
    print('%-40s %8.1f us (not a patch kernel)' % (L.name, s.elapsed_time(e) * 1e3))
    return
  m = act.astype(np.float64).mean(0)
  tiles = m[8]
  print('%-40s %8.1f us ctas %d tiles/cta %.1f | MMA total %.0f: wait acc %.0f a %.0f b %.0f | '
        'A prod total %.0f wait %.0f | B prod total %.0f wait %.0f | EPI total %.0f wait_full %.0f bar %.0f '
        'ld %.0f math %.0f fence %.0f issue %.0f | per tile: mma %.0f' % (L.name, s.elapsed_time(e) * 1e3, len(act), tiles, m[4], m[5], m[6], m[7],
                                 m[0], m[1], m[2], m[3], m[9], m[10], m[11], m[12], m[13], m[14], m[15], m[4] / max(tiles, 1)))


G._run_layer = run
for _ in range(2):
  G.forward(x)
  print('----')
