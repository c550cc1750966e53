"""Reader for TensorFlow-1 checkpoint bundles (`<prefix>.index` + `<prefix>.data-NNNNN-of-MMMMM`), so the
reference's published AdVoc / MelspecGAN checkpoints load without TensorFlow.

The reference restores through `tf.train.Saver` / `MonitoredTrainingSession`
(scripts/spectrogram_advoc.py:57-64, models/advoc/train_evaluate.py:61-64, README.md:188-195); the
files it reads are TF "tensor bundles":

  * `.index` is a LevelDB-format sorted table: data blocks of prefix-compressed (key, value) entries
    followed by a restart array, each block closed by a 5-byte trailer (compression type, masked
    CRC-32C), then a meta-index block, an index block and a 48-byte footer ending in the magic
    0xdb4775248b80fb57.  The empty key holds a `BundleHeaderProto`, every other key is a variable
    name holding a `BundleEntryProto` {dtype, shape, shard_id, offset, size, crc32c}.
  * `.data-*` hold the raw little-endian tensor bytes at (offset, size) of shard `shard_id`.

`write_bundle` emits the same layout TensorFlow's BundleWriter does (no block compression, entries sorted
by key, tensors back to back in the data shard), so checkpoints trained here restore in the reference's
`tf.train.Saver`.

Pure host-side Python: file parsing is not on the device hot path.  **Parity unpinned**: no checkpoint
ships with the reference and TensorFlow is not installable here, so reader and writer are built to the
published formats (LevelDB table format, tensor_bundle.proto) and exercised against each other, against a
bundle assembled byte by byte in tests/test_tf_bundle.py and against CRC-32C known answers, not against a
TF-produced file.
"""
import os
import struct

import numpy as np

TABLE_MAGIC = 0xdb4775248b80fb57
_FOOTER_LEN = 48
_MASK_DELTA = 0xa282ead8

# tensorflow/core/framework/types.proto
_DTYPES = {
    1: np.dtype('<f4'), 2: np.dtype('<f8'), 3: np.dtype('<i4'), 4: np.dtype('u1'), 5: np.dtype('<i2'),
    6: np.dtype('i1'), 9: np.dtype('<i8'), 10: np.dtype('?'), 17: np.dtype('<u2'), 19: np.dtype('<f2'),
    22: np.dtype('<u4'), 23: np.dtype('<u8'),
}
DT_OF = {v: k for k, v in _DTYPES.items()}


class BundleError(ValueError):
  pass


# ---------------------------------------------------------------------------------------------
# CRC-32C (Castagnoli), masked the LevelDB way
# ---------------------------------------------------------------------------------------------
def _make_table():
  tab = []
  for n in range(256):
    c = n
    for _ in range(8):
      c = (c >> 1) ^ 0x82f63b78 if c & 1 else c >> 1
    tab.append(c)
  return tab


_CRC_TABLE = _make_table()


def _crc32c_bytes(data, c):
  """Byte-at-a-time register update (no init / final xor)."""
  tab = _CRC_TABLE
  for b in data:
    c = tab[(c ^ b) & 0xff] ^ (c >> 8)
  return c


_CRC_LANES = 4096


def crc32c(data, crc=0):
  """CRC-32C of `data`, continuing from `crc` (0 for a fresh checksum).  Large buffers (checkpoint tensors
  run to hundreds of MB) are cut into 4096 equal chunks whose registers advance together as one numpy
  vector; the register update is linear over GF(2), so the chunk results are folded in order with the
  32x32 bit matrix of "advance by one chunk of zeros" (computed alongside from the 32 basis vectors)."""
  mv = memoryview(data).cast('B')
  n = len(mv)
  state = crc ^ 0xffffffff
  L = n // _CRC_LANES
  if L < 16:
    return _crc32c_bytes(bytes(mv), state) ^ 0xffffffff
  K = _CRC_LANES
  tab = np.array(_CRC_TABLE, dtype=np.uint32)
  cols = np.ascontiguousarray(np.frombuffer(mv[:K * L], dtype=np.uint8).reshape(K, L).T)   # [L, K]
  st = np.zeros(K, dtype=np.uint32)
  zb = np.left_shift(np.uint32(1), np.arange(32, dtype=np.uint32))
  for j in range(L):
    st = tab[(st ^ cols[j]) & 0xff] ^ (st >> 8)
    zb = tab[zb & 0xff] ^ (zb >> 8)
  zcols = [int(v) for v in zb]
  for r in st.tolist():
    adv, bit = 0, 0
    while state:
      if state & 1:
        adv ^= zcols[bit]
      state >>= 1
      bit += 1
    state = adv ^ r
  return _crc32c_bytes(bytes(mv[K * L:]), state) ^ 0xffffffff


def mask_crc(crc):
  return ((((crc >> 15) | (crc << 17)) & 0xffffffff) + _MASK_DELTA) & 0xffffffff


# ---------------------------------------------------------------------------------------------
# varints / minimal protobuf wire decoding
# ---------------------------------------------------------------------------------------------
def _varint(buf, pos):
  out, shift = 0, 0
  while True:
    if pos >= len(buf):
      raise BundleError('truncated varint')
    b = buf[pos]
    pos += 1
    out |= (b & 0x7f) << shift
    if not b & 0x80:
      return out, pos
    shift += 7
    if shift > 63:
      raise BundleError('varint too long')


def _proto_fields(buf):
  """Yield (field number, wire type, value) of one serialised message."""
  pos = 0
  while pos < len(buf):
    tag, pos = _varint(buf, pos)
    field, wt = tag >> 3, tag & 7
    if wt == 0:
      val, pos = _varint(buf, pos)
    elif wt == 1:
      val = struct.unpack_from('<Q', buf, pos)[0]
      pos += 8
    elif wt == 2:
      n, pos = _varint(buf, pos)
      val = bytes(buf[pos:pos + n])
      if len(val) != n:
        raise BundleError('truncated length-delimited field')
      pos += n
    elif wt == 5:
      val = struct.unpack_from('<I', buf, pos)[0]
      pos += 4
    else:
      raise BundleError('unsupported protobuf wire type %d' % wt)
    yield field, wt, val


def _signed64(v):
  return v - (1 << 64) if v >= (1 << 63) else v


def _parse_shape(buf):
  """TensorShapeProto: repeated Dim dim = 2 {int64 size = 1}; bool unknown_rank = 3."""
  dims = []
  for field, _, val in _proto_fields(buf):
    if field == 2:
      size = 0
      for f2, _, v2 in _proto_fields(val):
        if f2 == 1:
          size = _signed64(v2)
      dims.append(size)
    elif field == 3 and val:
      raise BundleError('tensor of unknown rank in a checkpoint')
  return tuple(dims)


def _parse_entry(buf):
  """BundleEntryProto (tensorflow/core/protobuf/tensor_bundle.proto)."""
  e = dict(dtype=0, shape=(), shard_id=0, offset=0, size=0, crc32c=None, sliced=False)
  for field, _, val in _proto_fields(buf):
    if field == 1:
      e['dtype'] = val
    elif field == 2:
      e['shape'] = _parse_shape(val)
    elif field == 3:
      e['shard_id'] = val
    elif field == 4:
      e['offset'] = _signed64(val)
    elif field == 5:
      e['size'] = _signed64(val)
    elif field == 6:
      e['crc32c'] = val
    elif field == 7:
      e['sliced'] = True
  return e


def _parse_header(buf):
  """BundleHeaderProto: int32 num_shards = 1; Endianness endianness = 2 (LITTLE = 0); VersionDef version = 3."""
  h = dict(num_shards=1, endianness=0)
  for field, _, val in _proto_fields(buf):
    if field == 1:
      h['num_shards'] = val
    elif field == 2:
      h['endianness'] = val
  return h


# ---------------------------------------------------------------------------------------------
# the sorted table
# ---------------------------------------------------------------------------------------------
def _read_block(data, offset, size, verify):
  end = offset + size
  if end + 5 > len(data):
    raise BundleError('block handle past the end of the index file')
  ctype = data[end]
  if verify:
    stored = struct.unpack_from('<I', data, end + 1)[0]
    if mask_crc(crc32c(data[offset:end + 1])) != stored:
      raise BundleError('index block checksum mismatch at offset %d' % offset)
  if ctype != 0:
    raise NotImplementedError('compressed index block (type %d); TF writes checkpoints uncompressed' % ctype)
  return data[offset:end]


def _block_entries(block):
  """Entries of one block in order: prefix-compressed keys, restart array ignored (sequential scan)."""
  if len(block) < 4:
    raise BundleError('block too small')
  n_restarts = struct.unpack_from('<I', block, len(block) - 4)[0]
  limit = len(block) - 4 - 4 * n_restarts
  if limit < 0:
    raise BundleError('bad restart count')
  pos, key = 0, b''
  while pos < limit:
    shared, pos = _varint(block, pos)
    non_shared, pos = _varint(block, pos)
    vlen, pos = _varint(block, pos)
    if shared > len(key) or pos + non_shared + vlen > limit:
      raise BundleError('corrupt block entry')
    key = key[:shared] + bytes(block[pos:pos + non_shared])
    pos += non_shared
    yield key, bytes(block[pos:pos + vlen])
    pos += vlen


def read_index(prefix, verify=True):
  """-> (header dict, {variable name: entry dict}) of `<prefix>.index`."""
  path = prefix + '.index'
  with open(path, 'rb') as f:
    data = f.read()
  if len(data) < _FOOTER_LEN:
    raise BundleError('%s: too short for a table footer' % path)
  footer = data[-_FOOTER_LEN:]
  if struct.unpack_from('<Q', footer, 40)[0] != TABLE_MAGIC:
    raise BundleError('%s: not a TensorFlow checkpoint index (bad table magic)' % path)
  pos = 0
  _, pos = _varint(footer, pos)        # meta-index handle (unused)
  _, pos = _varint(footer, pos)
  idx_off, pos = _varint(footer, pos)
  idx_size, pos = _varint(footer, pos)
  header, entries = None, {}
  for _, handle in _block_entries(_read_block(data, idx_off, idx_size, verify)):
    off, p = _varint(handle, 0)
    size, _ = _varint(handle, p)
    for key, val in _block_entries(_read_block(data, off, size, verify)):
      if key == b'':
        header = _parse_header(val)
      else:
        entries[key.decode('utf-8')] = _parse_entry(val)
  if header is None:
    raise BundleError('%s: no bundle header entry' % path)
  if header['endianness'] != 0:
    raise NotImplementedError('big-endian checkpoint')
  return header, entries


def list_variables(prefix):
  """[(name, shape)] sorted by name, like tf.train.list_variables."""
  _, entries = read_index(prefix)
  return [(k, list(e['shape'])) for k, e in sorted(entries.items())]


def read_bundle(prefix, names=None, verify_data=False):
  """Read the checkpoint at `prefix` -> {variable name: numpy array}.  `names`: optional iterable or
  predicate selecting variables.  `verify_data` also checks every tensor's CRC-32C (~50 MB/s on one
  core; the index blocks are always verified)."""
  header, entries = read_index(prefix)
  if names is not None and not callable(names):
    wanted = set(names)
    missing = wanted - set(entries)
    if missing:
      raise KeyError('not in checkpoint %s: %s' % (prefix, sorted(missing)))
    names = wanted.__contains__
  shards, out = {}, {}
  try:
    for name, e in sorted(entries.items()):
      if names is not None and not names(name):
        continue
      if e['sliced']:
        raise NotImplementedError('%s: partitioned (sliced) variable' % name)
      if e['dtype'] not in _DTYPES:
        raise NotImplementedError('%s: unsupported dtype enum %d' % (name, e['dtype']))
      dt = _DTYPES[e['dtype']]
      count = int(np.prod(e['shape'], dtype=np.int64)) if e['shape'] else 1
      if count * dt.itemsize != e['size']:
        raise BundleError('%s: %d bytes stored for shape %s of %s' % (name, e['size'], e['shape'], dt))
      sid = e['shard_id']
      if sid not in shards:
        shards[sid] = open('%s.data-%05d-of-%05d' % (prefix, sid, header['num_shards']), 'rb')
      f = shards[sid]
      f.seek(e['offset'])
      raw = f.read(e['size'])
      if len(raw) != e['size']:
        raise BundleError('%s: data shard %d truncated' % (name, sid))
      if verify_data and e['crc32c'] is not None and mask_crc(crc32c(raw)) != e['crc32c']:
        raise BundleError('%s: tensor checksum mismatch' % name)
      out[name] = np.frombuffer(raw, dtype=dt).reshape(e['shape']).copy()
  finally:
    for f in shards.values():
      f.close()
  return out


# ---------------------------------------------------------------------------------------------
# writer
# ---------------------------------------------------------------------------------------------
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


def _entry_proto(dtype_enum, shape, shard, offset, size, crc):
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


def _header_proto(num_shards):
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


def write_bundle(prefix, tensors, block_size=4096, restart_interval=16, num_shards=1):
  """Write {name: numpy array} as the TF-1 checkpoint `<prefix>.index` + `<prefix>.data-*`.  `block_size` /
  `restart_interval` are the table's data-block size and restart spacing (LevelDB defaults 4096 / 16); with
  num_shards > 1 the tensors are dealt round-robin over the data shards."""
  names = sorted(tensors, key=lambda s: s.encode('utf-8'))
  shard_files = [open('%s.data-%05d-of-%05d' % (prefix, s, num_shards), 'wb') for s in range(num_shards)]
  kv = [(b'', _header_proto(num_shards))]
  for i, name in enumerate(names):
    a = np.asarray(tensors[name])
    a = a.astype(a.dtype.newbyteorder('<'), copy=False)
    raw = a.tobytes()
    sid = i % num_shards
    off = shard_files[sid].tell()
    shard_files[sid].write(raw)
    kv.append((name.encode('utf-8'),
               _entry_proto(DT_OF[a.dtype], a.shape, sid, off, len(raw), mask_crc(crc32c(raw)))))
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


def is_bundle(prefix):
  return os.path.isfile(prefix + '.index')


def latest_checkpoint(train_dir):
  """The prefix named by `<train_dir>/checkpoint` (`model_checkpoint_path: "..."`), like
  tf.train.latest_checkpoint (scripts/spectrogram_advoc.py takes the prefix itself: --ckpt_fp)."""
  state = os.path.join(train_dir, 'checkpoint')
  if not os.path.isfile(state):
    return None
  with open(state) as f:
    for line in f:
      line = line.strip()
      if line.startswith('model_checkpoint_path:'):
        p = line.split(':', 1)[1].strip().strip('"')
        if not os.path.isabs(p):
          p = os.path.join(train_dir, p)
        return p if is_bundle(p) else None
  return None
