"""Developer probe: cycles per TMA tiled load as a function of box size and loads in flight (csrc/selftest.cu)."""
import ctypes as C
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from advoc_b200 import _native as N
torch.zeros(1, device='cuda')
fn = N.lib().advoc_selftest_tma_rate
fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
for ctas, lanes in ((148, 1), (148, 2), (148, 4), (148, 8), (296, 1)):
  for depth in (2, 4):
    for rows in (16, 32, 64, 128, 256):
      if lanes * depth * rows * 128 > (200 if ctas == 148 else 100) * 1024:
        continue
      reps = 1024
      out = np.zeros(ctas, dtype=np.uint64)
      st = fn(out.ctypes.data, ctas, rows, reps, depth, lanes)
      assert st == 0, (st, N.last_error())
      cyc = out.astype(np.float64).mean() / (reps * lanes)      # cycles per load, all lanes of the CTA together
      per_sm = cyc / (ctas / 148.0)
      print('ctas %3d lanes %d depth %d box %5d B: %7.1f cycles/load per CTA = %6.1f B/clk per SM'
            % (ctas, lanes, depth, rows * 128, cyc, rows * 128 / per_sm))

print('pure issue cost: up to 64 loads into distinct slots, one barrier, no waits in between')
for rows in (8, 16, 32, 64, 128):
  out = np.zeros(148, dtype=np.uint64)
  st = fn(out.ctypes.data, 148, rows, 1, 1, 1)
  assert st == 0, (st, N.last_error())
  n = min(64, 64 * 1024 // (rows * 128))
  issue = (out >> np.uint64(32)).astype(np.float64).mean() / n
  total = (out & np.uint64(0xffffffff)).astype(np.float64).mean()
  print('box %5d B x %2d loads: %6.1f cycles to ISSUE each, all complete after %7.0f cycles (%5.1f B/clk per SM)'
        % (rows * 128, n, issue, total, n * rows * 128 / total))
