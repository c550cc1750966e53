import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from advoc_b200 import _native as N
lib = N.lib()
lib.advoc_selftest_desc_shift.restype = C.c_int
lib.advoc_selftest_desc_shift.argtypes = [C.c_void_p, C.c_int, C.c_int]
out = torch.zeros(128, 32, device='cuda')
for mode in (0, 1):
  for shift in (0, 1, 2, 3, 5, 7, 8, 9, 13, 66, 131):
    st = lib.advoc_selftest_desc_shift(out.data_ptr(), shift, mode)
    o = out.cpu()
    rows = torch.arange(128) + shift
    exp = ((rows % 64) * 32).float()[:, None] + torch.arange(32).float()[None, :]
    ok = bool(torch.equal(o, exp))
    # which row did each output row come from (using column 0), if it is a clean row copy
    src = (o[:, 0] / 32).round().long() % 64
    clean = bool(torch.equal(o, (src * 32).float()[:, None] + torch.arange(32).float()[None, :]))
    print('mode', mode, 'shift', shift, 'status', st, 'exact', ok, 'clean-rows', clean,
          'src rows (mod 64) of m=0..11:', src[:12].tolist())
