"""world_size-2 gloo test of the data-parallel host logic (advoc_b200/dist.py): sharding, the
flat-gradient all-reduce and the 1/world mean match a single-process gradient of the whole
batch.  Uses the CPU oracle nets as the differentiable function; no GPU."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _toy_grads(P, x, t):
  """Mean-loss gradient of a tiny conv 'discriminator' on batch slice (x, t)."""
  import torch.nn.functional as F
  w = P['w'].clone().requires_grad_(True)
  y = F.conv2d(x, w, padding=1)
  loss = ((y - t) ** 2).mean()
  (g,) = torch.autograd.grad(loss, [w])
  return g


def _worker(rank, world, port, out_dir):
  sys.path.insert(0, ROOT)
  os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                    MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
  from advoc_b200 import dist as D
  r, _, w = D.init('gloo')
  assert (r, w) == (rank, world)
  g = torch.Generator().manual_seed(0)
  P = {'w': torch.randn(4, 3, 3, 3, generator=g)}
  x = torch.randn(8, 3, 16, 16, generator=g)
  t = torch.randn(8, 4, 16, 16, generator=g)
  lo, hi = D.shard_range(8, world, rank)
  flat = torch.zeros(4 * 3 * 3 * 3 + 5)
  flat[:108] = _toy_grads(P, x[lo:hi], t[lo:hi]).reshape(-1)
  flat[108:] = float(rank + 1)                      # outside the reduced range: must stay local
  D.allreduce_sum_(flat, 0, 108)
  mean = flat[:108] / world                         # what adam's grad_scale = 1/world applies
  ref = _toy_grads(P, x, t).reshape(-1)
  ok = torch.allclose(mean, ref, atol=1e-6) and bool((flat[108:] == rank + 1).all())
  # the overlapped form TrainEngine uses: three buckets in flight at once, then wait
  flat2 = torch.zeros(113)
  flat2[:108] = _toy_grads(P, x[lo:hi], t[lo:hi]).reshape(-1)
  hs = [D.allreduce_sum_async(flat2, a, b) for a, b in ((40, 108), (10, 40), (0, 10), (5, 5))]
  assert hs[3] is None                               # empty range: nothing to do
  for h in hs[:3]:
    h.wait()
  ok = ok and torch.equal(flat2[:108], flat[:108]) and bool((flat2[108:] == 0).all())
  mx = D.max_over_ranks([float(rank), 1.0], 'cpu')
  ok = ok and mx == [float(world - 1), 1.0]
  with open(os.path.join(out_dir, 'ok%d' % rank), 'w') as f:
    f.write('1' if ok else '0')
  torch.distributed.destroy_process_group()


def test_two_rank_gloo_allreduce_matches_full_batch(tmp_path):
  port = _free_port()
  mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
  for r in range(2):
    assert open(os.path.join(str(tmp_path), 'ok%d' % r)).read() == '1'


def test_shard_range_contract():
  sys.path.insert(0, ROOT)
  from advoc_b200 import dist as D
  assert D.shard_range(256, 8, 3) == (96, 128)
  assert D.shard_range(32, 1, 0) == (0, 32)
  with pytest.raises(ValueError):
    D.shard_range(10, 4, 0)


def test_gradient_buckets_tile_the_generator_slice():
  """TrainEngine's all-reduce buckets (decoders | encoder_n..5 | encoder_4..1) are contiguous ranges of
  the name-sorted flat buffer and tile the generator slice exactly (host logic, CPU tensors)."""
  sys.path.insert(0, ROOT)
  from advoc_b200.train import FlatParams
  from oracle import nets_torch as O
  for spec, n in ((O.SMALL, 5), (O.REGULAR, 8)):
    f = FlatParams(O.init_params(spec, seed=0))
    idx = lambda name: int(name.split('/')[1].split('_')[1])
    dec = f.range_of(lambda m: m.startswith('generator/decoder_'))
    hi = f.range_of(lambda m: m.startswith('generator/encoder_') and idx(m) >= 5)
    lo = f.range_of(lambda m: m.startswith('generator/encoder_') and idx(m) < 5)
    assert dec[0] == 0 and dec[1] == lo[0] and lo[1] == hi[0] and hi[1] == f.n_gen
    assert f.dis_range() == (f.n_gen, f.total)
    big = (dec[1] - dec[0]) + (hi[1] - hi[0])
    if n == 8:
      assert big > 0.9 * f.n_gen      # the buckets that leave early carry most of the 54 M parameters
