"""TensorFlow-1 checkpoint reader (advoc_b200/tf_bundle.py) against a test-side
writer of the published formats, CRC-32C known answers, and the checkpoint loader on top of it.
CPU only.  Parity unpinned against TF-produced files: none ship with the reference."""
import os
import struct
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import tf_bundle_writer as W  # noqa: E402
from advoc_b200 import checkpoint, tf_bundle  # noqa: E402


def test_crc32c_known_answers():
  # RFC 3720 B.4 vectors (also LevelDB's crc32c_test)
  assert tf_bundle.crc32c(b'123456789') == 0xe3069283
  assert tf_bundle.crc32c(bytes(32)) == 0x8a9136aa
  assert tf_bundle.crc32c(b'\xff' * 32) == 0x62a8ab43
  assert tf_bundle.crc32c(bytes(range(32))) == 0x46dd794e
  assert tf_bundle.crc32c(bytes(range(31, -1, -1))) == 0x113fdb5c
  # incremental == one shot; the mask is a rotation + constant, so it is not the identity
  c = tf_bundle.crc32c(b'hello ')
  assert tf_bundle.crc32c(b'world', c) == tf_bundle.crc32c(b'hello world')
  assert tf_bundle.mask_crc(c) != c
  assert tf_bundle.mask_crc(0) == 0xa282ead8
  # the chunk-parallel path (buffers >= 64 KB) against the byte loop, with an odd tail and a running value
  rng = np.random.default_rng(3)
  for n in (65536, 300001, 1 << 20):
    buf = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
    for start in (0, 0x1234abcd):
      assert tf_bundle.crc32c(buf, start) == tf_bundle._crc32c_bytes(buf, start ^ 0xffffffff) ^ 0xffffffff
  assert tf_bundle.crc32c(b'123456789' * 10000) == tf_bundle.crc32c(b'123456789' * 5000, tf_bundle.crc32c(b'123456789' * 5000))


def _advoc_like(rng, with_slots=True):
  """Variables of an AdVoc-small training checkpoint (names: advoc_model_small.py scopes), tiny shapes."""
  T = {}
  chans = [1, 4, 8]
  for i in range(2):
    k = 'generator/encoder_%d/conv2d/' % (i + 1)
    T[k + 'kernel'] = rng.standard_normal((4, 4, chans[i], chans[i + 1])).astype(np.float32)
    T[k + 'bias'] = rng.standard_normal(chans[i + 1]).astype(np.float32)
  T['generator/decoder_1/conv2d_transpose/kernel'] = rng.standard_normal((4, 4, 1, 8)).astype(np.float32)
  T['generator/decoder_1/conv2d_transpose/bias'] = np.zeros(1, np.float32)
  T['discriminator/layer_1/conv2d/kernel'] = rng.standard_normal((4, 4, 2, 4)).astype(np.float32)
  T['discriminator/layer_1/conv2d/bias'] = rng.standard_normal(4).astype(np.float32)
  if with_slots:
    for k in list(T):
      T[k + '/Adam'] = rng.standard_normal(T[k].shape).astype(np.float32)
      T[k + '/Adam_1'] = np.abs(rng.standard_normal(T[k].shape)).astype(np.float32)
    T['beta1_power'] = np.float32(0.5 ** 7)
    T['beta2_power'] = np.float32(0.999 ** 7)
    T['beta1_power_1'] = np.float32(0.5 ** 7)
    T['beta2_power_1'] = np.float32(0.999 ** 7)
  T['global_step'] = np.int64(7)
  return T


@pytest.mark.parametrize('block_size,restart,shards', [(256, 16, 1), (64, 2, 1), (1 << 20, 16, 1), (300, 3, 3)])
def test_roundtrip(tmp_path, block_size, restart, shards):
  rng = np.random.default_rng(0)
  T = _advoc_like(rng)
  T['misc/f64'] = rng.standard_normal((3, 5))
  T['misc/i32'] = np.arange(-4, 8, dtype=np.int32).reshape(2, 6)
  T['misc/empty'] = np.zeros((0, 3), np.float32)
  T['misc/f16'] = rng.standard_normal(9).astype(np.float16)
  T['misc/bool'] = np.array([True, False, True])
  prefix = str(tmp_path / 'model.ckpt-7')
  W.write_bundle(prefix, T, block_size=block_size, restart_interval=restart, num_shards=shards)
  got = tf_bundle.read_bundle(prefix, verify_data=True)
  assert sorted(got) == sorted(T)
  for k, v in T.items():
    v = np.asarray(v)
    assert got[k].dtype == v.dtype and got[k].shape == v.shape, k
    np.testing.assert_array_equal(got[k], v)
  lv = dict(tf_bundle.list_variables(prefix))
  assert lv['generator/encoder_1/conv2d/kernel'] == [4, 4, 1, 4] and lv['global_step'] == []
  sub = tf_bundle.read_bundle(prefix, names=['global_step', 'misc/i32'])
  assert sorted(sub) == ['global_step', 'misc/i32']
  with pytest.raises(KeyError):
    tf_bundle.read_bundle(prefix, names=['nope'])


def test_hand_assembled_index(tmp_path):
  """A one-variable bundle written out byte by byte here (no writer helper): header + one float32[2]."""
  prefix = str(tmp_path / 'hand')
  payload = struct.pack('<2f', 1.5, -2.0)
  with open(prefix + '.data-00000-of-00001', 'wb') as f:
    f.write(payload)
  crc = tf_bundle.mask_crc(tf_bundle.crc32c(payload))
  header = bytes([0x08, 0x01])                                   # num_shards = 1
  entry = bytes([0x08, 0x01,                                      # dtype DT_FLOAT
                 0x12, 0x04, 0x12, 0x02, 0x08, 0x02,              # shape {dim {size: 2}}
                 0x28, 0x08,                                      # size = 8
                 0x35]) + struct.pack('<I', crc)                  # crc32c (fixed32)
  # entries: (shared, non_shared, value_len, key suffix, value); one restart at 0
  block = bytes([0, 0, len(header)]) + header + bytes([0, 1, len(entry)]) + b'w' + entry
  block += struct.pack('<II', 0, 1)
  def trailer(b):
    return b'\x00' + struct.pack('<I', tf_bundle.mask_crc(tf_bundle.crc32c(b + b'\x00')))
  meta = struct.pack('<II', 0, 1)
  out = block + trailer(block)
  meta_off = len(out)
  out += meta + trailer(meta)
  idx = bytes([0, 1, 2]) + b'x' + bytes([0, len(block)]) + struct.pack('<II', 0, 1)
  idx_off = len(out)
  out += idx + trailer(idx)
  footer = bytes([meta_off, len(meta), idx_off, len(idx)])
  footer += b'\x00' * (40 - len(footer)) + struct.pack('<Q', 0xdb4775248b80fb57)
  with open(prefix + '.index', 'wb') as f:
    f.write(out + footer)
  got = tf_bundle.read_bundle(prefix, verify_data=True)
  assert list(got) == ['w']
  np.testing.assert_array_equal(got['w'], np.array([1.5, -2.0], np.float32))


def test_corruption_is_detected(tmp_path):
  rng = np.random.default_rng(1)
  T = _advoc_like(rng, with_slots=False)
  prefix = str(tmp_path / 'm')
  W.write_bundle(prefix, T)
  raw = bytearray(open(prefix + '.index', 'rb').read())
  bad = bytearray(raw)
  bad[10] ^= 0x40
  open(prefix + '.index', 'wb').write(bad)
  with pytest.raises(tf_bundle.BundleError):
    tf_bundle.read_bundle(prefix)
  bad = bytearray(raw)
  bad[-1] ^= 0xff                                   # table magic
  open(prefix + '.index', 'wb').write(bad)
  with pytest.raises(tf_bundle.BundleError):
    tf_bundle.read_bundle(prefix)
  open(prefix + '.index', 'wb').write(raw)
  data_fp = prefix + '.data-00000-of-00001'
  d = bytearray(open(data_fp, 'rb').read())
  d[5] ^= 1
  open(data_fp, 'wb').write(d)
  tf_bundle.read_bundle(prefix)                      # tensor checksums are opt-in
  with pytest.raises(tf_bundle.BundleError):
    tf_bundle.read_bundle(prefix, verify_data=True)
  open(data_fp, 'wb').write(d[:-3])
  with pytest.raises(tf_bundle.BundleError):
    tf_bundle.read_bundle(prefix)


def test_checkpoint_loader_accepts_tf_prefix(tmp_path):
  rng = np.random.default_rng(2)
  T = _advoc_like(rng)
  d = tmp_path / 'train'
  d.mkdir()
  prefix = str(d / 'model.ckpt-7')
  W.write_bundle(prefix, T)
  (d / 'checkpoint').write_text('model_checkpoint_path: "model.ckpt-7"\nall_model_checkpoint_paths: "model.ckpt-7"\n')
  for path in (prefix, prefix + '.index', prefix + '.data-00000-of-00001', str(d)):
    P, step = checkpoint.load_params(path, device='cpu')
    assert step == 7
    assert sorted(P) == sorted(k for k in T if k.startswith(('generator/', 'discriminator/'))
                               and not k.endswith(('/Adam', '/Adam_1')))
    np.testing.assert_array_equal(P['generator/encoder_2/conv2d/kernel'].numpy(), T['generator/encoder_2/conv2d/kernel'])
  m, v, powers = checkpoint.load_adam_slots(prefix, device='cpu')
  assert sorted(m) == sorted(P) and sorted(v) == sorted(P)
  np.testing.assert_array_equal(v['discriminator/layer_1/conv2d/bias'].numpy(), T['discriminator/layer_1/conv2d/bias/Adam_1'])
  assert abs(powers['beta1_power'] - 0.5 ** 7) < 1e-9 and 'beta2_power_1' in powers
  with pytest.raises(FileNotFoundError):
    checkpoint.load_params(str(tmp_path / 'missing'), device='cpu')
  # save_params with a non-.npz path writes a TF prefix + the state file; extras keep their own names
  import torch
  d2 = tmp_path / 'out'
  d2.mkdir()
  slots = {k + '/Adam': torch.from_numpy(np.asarray(T[k + '/Adam'])) for k in P}
  checkpoint.save_params(str(d2 / 'model.ckpt-9'), P, step=9, extra=slots)
  assert sorted(os.listdir(str(d2))) == ['checkpoint', 'model.ckpt-9.data-00000-of-00001', 'model.ckpt-9.index']
  Q, step = checkpoint.load_params(str(d2), device='cpu')
  assert step == 9 and sorted(Q) == sorted(P)
  assert all(torch.equal(Q[k], P[k]) for k in P)
  back = tf_bundle.read_bundle(str(d2 / 'model.ckpt-9'), verify_data=True)
  np.testing.assert_array_equal(back['generator/encoder_1/conv2d/bias/Adam'], T['generator/encoder_1/conv2d/bias/Adam'])
  # the .npz container still loads through the same entry point
  npz = str(tmp_path / 'p.npz')
  import torch
  checkpoint.save_params(npz, {k: torch.from_numpy(np.asarray(T[k])) for k in P}, step=3)
  Q, step = checkpoint.load_params(npz, device='cpu')
  assert step == 3 and sorted(Q) == sorted(P)
