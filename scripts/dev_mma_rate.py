"""Developer probe: tcgen05.mma issue / completion rate per instruction (see csrc/selftest.cu)."""
import ctypes as C
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from advoc_b200 import _native as N
torch.zeros(1, device='cuda')
lib = N.lib()
fn = lib.advoc_selftest_mma_rate
fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
for mode, name in [(16, '128 threads, 193 KB smem'), (16 + 8192, '224 threads'), (16 + 16384, '216 KB smem'),
                   (16 + 8192 + 16384, '224 threads + 216 KB')]:
  for n in (32, 64, 128):
    ctas, reps = 148, 2048
    out = np.zeros((ctas, 2), dtype=np.uint64)
    st = fn(out.ctypes.data, ctas, n, reps, mode)
    assert st == 0, st
    print('%-26s N=%3d  issue %.1f cyc/MMA  complete %.1f cyc/MMA (ideal %d)'
          % (name, n, out[:, 0].mean() / reps, out[:, 1].mean() / reps, 128 * n // 256))
