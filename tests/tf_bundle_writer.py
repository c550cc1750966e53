"""Test-side alias of the bundle writer (now product code: advoc_b200.tf_bundle.write_bundle); kept so the
tests that need small data blocks keep one import."""
from advoc_b200.tf_bundle import write_bundle as _write


def write_bundle(prefix, tensors, block_size=256, restart_interval=16, num_shards=1):
  return _write(prefix, tensors, block_size=block_size, restart_interval=restart_interval, num_shards=num_shards)
