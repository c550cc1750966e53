"""Developer tool: per-role cycle counters of deconv_one_tc_kernel (decoder_1).  ADVOC_ONE_PROFILE=1."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from advoc_b200 import _native as N, nets
model = sys.argv[1] if len(sys.argv) > 1 else 'small'
small = model == 'small'
spec = nets.GenSpec(32 if small else 64, 5 if small else 8, (5, 4) if small else (8, 7, 6))
P = nets.init_params(spec.ngf, spec.ngf, spec.n_enc, seed=0)
G = nets.Generator(spec, P, 32); G.prepare()
x = torch.rand(32, 256, 513, 1, device='cuda')
fn = N.lib().advoc_one_profile_read; fn.restype = C.c_int; fn.argtypes = [C.c_void_p]
buf = np.zeros((256, 8), dtype=np.uint64)
for _ in range(3):
  G.forward(x); torch.cuda.synchronize()
fn(buf.ctypes.data)
G.forward(x); torch.cuda.synchronize()
fn(buf.ctypes.data)
a = buf[buf[:, 2] > 0].astype(np.float64).mean(0)
print('tiles/cta %.1f | producer total %.0f wait_empty %.0f | MMA total %.0f wait_acc %.0f wait_a %.0f | epilogue total %.0f wait_full %.0f'
      % (a[7], a[0], a[1], a[2], a[3], a[4], a[5], a[6]))
print('per tile: total %.0f  epilogue busy %.0f  mma busy %.0f  producer busy %.0f' % (a[2] / a[7], (a[5] - a[6]) / a[7], (a[2] - a[3] - a[4]) / a[7], (a[0] - a[1]) / a[7]))
