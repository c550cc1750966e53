"""Developer probe: what bounds encoder_1 (one input channel, k4 s2, 32 x [256, 513] -> [128, 257, C])?
Times the layer with one / two outputs, fp32 / fp16, on the tcgen05 kernel and on the CUDA-core kernel."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from advoc_b200 import _native as N
from advoc_b200 import nets

B, C = 32, int(sys.argv[1]) if len(sys.argv) > 1 else 32
x = torch.rand(B, 256, 513, 1, device='cuda')
k = (torch.randn(4, 4, 1, C, device='cuda') * 0.02)
b = torch.zeros(C, device='cuda')
L = nets._Conv('t', 'conv', nets._desc(B, 256, 513, 1, C, 2, 2, 1, 1, 128, 257, N.MATH_AUTO))
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')


def run(tag, o0, o1, pad=0, flush_kind='write'):
  ld1 = o1.shape[3] if o1 is not None else 0
  ep = nets._epilogue(b, o0, C, 0, N.ACT_LRELU, o1, ld1, ld1 - C if o1 is not None else 0, N.ACT_RELU, row_pad0=pad)
  ts = []
  for i in range(8):
    if flush_kind == 'write':
      flush.zero_()
    elif flush_kind == 'read':
      flush.sum()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    L.run(x, 1, k, ep)
    e.record()
    torch.cuda.synchronize()
    if i >= 3:
      ts.append(s.elapsed_time(e) * 1e3)
  print('%-44s %-22s %6.1f us' % (tag, L.kernel_family(), sum(ts) / len(ts)))


for dt, name in ((torch.float16, 'fp16'), (torch.float32, 'fp32')):
  o0 = torch.zeros((B, 128, 257, C), device='cuda', dtype=dt)
  o0p = torch.zeros((B, 128, 258, C), device='cuda', dtype=dt)
  o1 = torch.zeros((B, 128, 257, 2 * C), device='cuda', dtype=dt)
  run(name + ' two outputs, L2 flushed by writes', o0, o1)
  run(name + ' two outputs, out0 rows padded', o0p, o1, pad=1)
  run(name + ' two outputs, L2 flushed by reads', o0, o1, flush_kind='read')
  run(name + ' two outputs, no flush', o0, o1, flush_kind='none')
  run(name + ' out0 only', o0, None)
