"""Worker of tests/test_gpu_train.py::test_two_rank_nccl_train_loop_matches_accumulated_gradients:
one rank of a 2-GPU data-parallel `TrainEngine.train_loop` (NCCL, overlapped all-reduces).  Launched
with `python -m torch.distributed.run --nproc-per-node 2 tests/dp_worker.py <out_dir>`; every rank
saves its parameters after the step and the all-reduced gradients."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def data(rank):
  g = torch.Generator().manual_seed(100 + rank)
  tgt = [torch.randn(1, 256, 513, 1, generator=g).abs() * 0.1 for _ in range(2)]
  x = [t + torch.randn(1, 256, 513, 1, generator=g) * 0.02 for t in tgt]
  return x, tgt


def main():
  out_dir = sys.argv[1]
  from advoc_b200 import _native as N
  from advoc_b200 import dist as D
  from advoc_b200 import nets
  from advoc_b200.train import TrainEngine
  rank, local, world = D.init('nccl')
  torch.cuda.set_device(local)
  spec = nets.GenSpec(32, 5, (5, 4))
  P = nets.init_params(32, 32, 5, seed=0)
  eng = TrainEngine(spec, 32, P, 1, math=N.MATH_FP32 if os.environ.get('DP_MATH') == 'fp32' else N.MATH_AUTO,
                    world_size=world, rank=rank, overlap=os.environ.get('DP_OVERLAP', '1') == '1')
  x, tgt = data(rank)
  step = eng.train_loop((x[0].cuda(), tgt[0].cuda()), (x[1].cuda(), tgt[1].cuda()), dropout=None)
  losses = eng.loss_values()
  torch.cuda.synchronize()
  assert step == 1 and N.debug_flags() == 0
  torch.save({'p': eng.flat.p.cpu(), 'g': eng.flat.g.cpu(), 'losses': losses}, os.path.join(out_dir, 'rank%d.pt' % rank))
  torch.distributed.barrier()
  torch.distributed.destroy_process_group()


if __name__ == '__main__':
  main()
