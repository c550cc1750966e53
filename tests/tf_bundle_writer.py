"""Test-side WRITER of TensorFlow-1 checkpoint bundles (test infrastructure, not product code).

Builds the files byte by byte from the published formats, independently of the reader's parsing
code: LevelDB table blocks (prefix-compressed keys, restart points every `restart_interval`
entries, 5-byte trailers with a masked CRC-32C), meta-index + index blocks, the 48-byte footer, and
hand-serialised BundleHeaderProto / BundleEntryProto messages (tensor_bundle.proto).  TensorFlow's
BundleWriter emits exactly this layout (no block compression, 4 KB-ish data blocks, entries sorted
by key, tensors back to back in `.data-00000-of-00001`).
"""
import struct

import numpy as np

from advoc_b200.tf_bundle import DT_OF, crc32c, mask_crc, TABLE_MAGIC


def _vi(n):
  out = bytearray()
  n &= (1 << 64) - 1
  while True:
    b = n & 0x7f
    n >>= 7
    if n:
      out.append(b | 0x80)
    else:
      out.append(b)
      return bytes(out)


def _field(num, wt, payload):
  return _vi((num << 3) | wt) + payload


def _shape_proto(shape):
  out = b''
  for d in shape:
    dim = _field(1, 0, _vi(d)) if d else b''
    out += _field(2, 2, _vi(len(dim)) + dim)
  return out


def entry_proto(dtype_enum, shape, shard, offset, size, crc):
  out = _field(1, 0, _vi(dtype_enum))
  sp = _shape_proto(shape)
  out += _field(2, 2, _vi(len(sp)) + sp)
  if shard:
    out += _field(3, 0, _vi(shard))
  if offset:
    out += _field(4, 0, _vi(offset))
  out += _field(5, 0, _vi(size))
  out += _field(6, 5, struct.pack('<I', crc))
  return out


def header_proto(num_shards):
  version = _field(1, 0, _vi(1))                      # VersionDef.producer = 1
  return _field(1, 0, _vi(num_shards)) + _field(3, 2, _vi(len(version)) + version)   # endianness LITTLE = default


class _BlockBuilder(object):
  def __init__(self, restart_interval):
    self.ri, self.buf, self.restarts, self.n, self.last = restart_interval, bytearray(), [0], 0, b''

  def add(self, key, value):
    shared = 0
    if self.n % self.ri == 0:
      if self.n:
        self.restarts.append(len(self.buf))
    else:
      while shared < min(len(key), len(self.last)) and key[shared] == self.last[shared]:
        shared += 1
    self.buf += _vi(shared) + _vi(len(key) - shared) + _vi(len(value)) + key[shared:] + value
    self.last, self.n = key, self.n + 1

  def finish(self):
    out = bytes(self.buf) + b''.join(struct.pack('<I', r) for r in self.restarts)
    return out + struct.pack('<I', len(self.restarts))


def _emit(fileobj, block):
  """Append block + trailer; return its (offset, size) handle."""
  off = fileobj.tell()
  fileobj.write(block)
  fileobj.write(b'\x00' + struct.pack('<I', mask_crc(crc32c(block + b'\x00'))))
  return off, len(block)


def write_bundle(prefix, tensors, block_size=256, restart_interval=16, num_shards=1):
  """tensors: {name: numpy array}.  Small `block_size` forces several data blocks; with
  num_shards > 1 the tensors are dealt round-robin over the data shards."""
  names = sorted(tensors, key=lambda s: s.encode('utf-8'))
  shard_files = [open('%s.data-%05d-of-%05d' % (prefix, s, num_shards), 'wb') for s in range(num_shards)]
  kv = [(b'', header_proto(num_shards))]
  for i, name in enumerate(names):
    a = np.asarray(tensors[name])
    a = a.astype(a.dtype.newbyteorder('<'), copy=False)
    raw = a.tobytes()
    sid = i % num_shards
    off = shard_files[sid].tell()
    shard_files[sid].write(raw)
    kv.append((name.encode('utf-8'),
               entry_proto(DT_OF[a.dtype], a.shape, sid, off, len(raw), mask_crc(crc32c(raw)))))
  for f in shard_files:
    f.close()
  with open(prefix + '.index', 'wb') as f:
    index = _BlockBuilder(1)
    blk = _BlockBuilder(restart_interval)
    for key, val in kv:
      blk.add(key, val)
      if len(blk.buf) >= block_size:
        h = _emit(f, blk.finish())
        index.add(blk.last + b'\x00', _vi(h[0]) + _vi(h[1]))   # any separator >= the block's last key
        blk = _BlockBuilder(restart_interval)
    if blk.n:
      h = _emit(f, blk.finish())
      index.add(blk.last + b'\x00', _vi(h[0]) + _vi(h[1]))
    mh = _emit(f, _BlockBuilder(1).finish())
    ih = _emit(f, index.finish())
    footer = _vi(mh[0]) + _vi(mh[1]) + _vi(ih[0]) + _vi(ih[1])
    footer += b'\x00' * (40 - len(footer)) + struct.pack('<Q', TABLE_MAGIC)
    f.write(footer)
