"""Data-parallel plumbing for the train step: one process per GPU (torchrun-style), NCCL over
NVLink 5 / NVSwitch on the GPU box, gloo in the CPU tests.

The reference is single-GPU (README.md:57,67; advoc/loader.py:208-211); data parallelism is the
addition BASELINE.json asks for: the global minibatch is split evenly over ranks, weights and
Adam state are replicated, and each optimiser step is preceded by ONE sum-all-reduce of that
network's slice of the flat fp32 gradient buffer (11 MB for D, 218 MB for G on the regular
model); `advoc_adam_tf_step(grad_scale=1/world)` turns the sum into the mean.
"""
import os

import torch
import torch.distributed as dist


def env_world():
  """(rank, local_rank, world_size) from the torchrun environment."""
  return (int(os.environ.get('RANK', '0')), int(os.environ.get('LOCAL_RANK', '0')),
          int(os.environ.get('WORLD_SIZE', '1')))


def init(backend=None):
  """Joins the default process group when launched under torchrun; returns (rank, local, world)."""
  rank, local, world = env_world()
  if world > 1 and not dist.is_initialized():
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    if backend is None:
      backend = 'nccl' if torch.cuda.is_available() else 'gloo'
    kw = {}
    if backend == 'nccl':
      torch.cuda.set_device(local)
      kw['device_id'] = torch.device('cuda', local)
    dist.init_process_group(backend, **kw)
  return rank, local, world


def bind_to_gpu_numa(local_rank):
  """Pins the calling process to the CPU cores NVML reports as closest to its GPU, so that the pinned
  host buffers it allocates afterwards (first touch) and its copy threads sit on that GPU's NUMA node.
  With eight ranks streaming results to the host at once (infer.MelToMag.run_stream: 16.8 MB per 0.63 ms
  each) unbound processes share one node's memory controllers.  Returns the number of cores bound to, or
  0 when NVML / the affinity call is unavailable (nothing changes then)."""
  try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(int(local_rank))
    pynvml.nvmlDeviceSetCpuAffinity(h)
    return len(os.sched_getaffinity(0))
  except Exception:
    return 0


def shard_range(global_batch, world, rank):
  """Even split of the global minibatch; the reference's batch is a free parameter (no BN by
  default, advoc_model.py:21), so ranks only need equal shares for the mean to be exact."""
  if global_batch % world != 0:
    raise ValueError('global batch %d is not divisible by world size %d' % (global_batch, world))
  per = global_batch // world
  return rank * per, (rank + 1) * per


def allreduce_sum_(flat, lo, hi, group=None, world=None):
  """In-place sum-all-reduce of flat[lo:hi] (a contiguous view of the flat gradient buffer)."""
  world = dist.get_world_size(group) if (world is None and dist.is_initialized()) else (world or 1)
  if world > 1:
    dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM, group=group)
  return flat


def allreduce_sum_async(flat, lo, hi, group=None, world=None):
  """Starts the in-place sum-all-reduce of flat[lo:hi] and returns the work handle (`.wait()` makes
  the current stream wait for it), or None when there are no peers or the range is empty.  NCCL
  orders the collective behind the work already queued on the current stream and runs it on its own
  stream, so later kernels on the current stream overlap it."""
  world = dist.get_world_size(group) if (world is None and dist.is_initialized()) else (world or 1)
  if world <= 1 or hi <= lo:
    return None
  return dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM, group=group, async_op=True)


def max_over_ranks(values, device):
  """Element-wise max of a list of floats over all ranks (timing: the slowest rank counts)."""
  t = torch.tensor(values, dtype=torch.float64, device=device)
  if dist.is_initialized() and dist.get_world_size() > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
  return [float(v) for v in t]
