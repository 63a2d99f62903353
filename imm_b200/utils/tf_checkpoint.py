"""TensorFlow "TensorBundle" (Saver V2) checkpoint reader / writer, without TensorFlow.

The reference saves with `tf.train.Saver(tf.global_variables()).save(session, '<logdir>/model.ckpt', global_step)`
(cnn_train_multi.py:436-439,511-513) and restores with `tf.train.Saver(var_list).restore(session, fname)` /
`tf.train.NewCheckpointReader(fname).get_variable_to_shape_map()` (cnn_train_multi.py:404-433, eval_imm.py:81-94).
Those calls touch exactly two files per checkpoint, which this module writes and reads bit-compatibly:

  <prefix>.index                 an SSTable (TensorFlow lib/io/table == the LevelDB table format): key "" ->
                                 BundleHeaderProto, key <variable name> -> BundleEntryProto, keys sorted bytewise;
  <prefix>.data-00000-of-00001   the tensors' raw little-endian bytes, back to back in key order;

plus the directory's `checkpoint` state file (text proto with model_checkpoint_path / all_model_checkpoint_paths)
that `tf.train.latest_checkpoint` reads.  The `.meta` MetaGraphDef is not written: nothing on the reference's
restore paths reads it.

Format facts restated here (published formats; TensorFlow 1.10 sources are not part of /root/reference):
  * table block  = entries | restart offsets (uint32 LE each) | num_restarts (uint32 LE); an entry is
    varint32 shared, varint32 non_shared, varint32 value_len, key suffix, value; restart interval 16 for data
    blocks and 1 for the index block; every block is followed by a 5-byte trailer: compression type (0 = none)
    and the masked CRC-32C of block + type byte;  mask(c) = rotr15(c) + 0xa282ead8;
  * footer (48 bytes) = metaindex handle | index handle (varint64 offset, varint64 size each) zero-padded to 40
    bytes | magic 0xdb4775248b80fb57 little-endian;
  * BundleHeaderProto  {1: num_shards (int32), 2: endianness (enum, LITTLE = 0), 3: VersionDef {1: producer}};
  * BundleEntryProto   {1: dtype (enum DataType), 2: TensorShapeProto {2: Dim {1: size}}, 3: shard_id,
                        4: offset (int64), 5: size (int64), 6: crc32c (fixed32, masked CRC-32C of the bytes)}.

CRC-32C comes from the C ABI library (immb_crc32c, host code); the module fails loudly when the library is missing.
"""
import ctypes
import os
import struct
from collections import OrderedDict

import numpy as np

TABLE_MAGIC = 0xdb4775248b80fb57
BLOCK_SIZE = 256 * 1024          # TensorFlow table::Options default
RESTART_INTERVAL = 16
MASK_DELTA = 0xa282ead8

# tensorflow/core/framework/types.proto
_DTYPES = {1: np.dtype('<f4'), 2: np.dtype('<f8'), 3: np.dtype('<i4'), 4: np.dtype('u1'), 5: np.dtype('<i2'),
           6: np.dtype('i1'), 9: np.dtype('<i8'), 10: np.dtype('bool'), 17: np.dtype('<u2'), 19: np.dtype('<f2'),
           22: np.dtype('<u4'), 23: np.dtype('<u8')}
_DTYPE_ENUM = {v: k for k, v in _DTYPES.items()}


def crc32c(data, crc=0):
  from .. import _lib
  if isinstance(data, np.ndarray):
    data = np.ascontiguousarray(data)
    return int(_lib.lib().immb_crc32c(ctypes.c_void_p(data.ctypes.data), data.nbytes, crc))
  data = bytes(data)
  return int(_lib.lib().immb_crc32c(ctypes.c_char_p(data), len(data), crc))


def mask_crc(c):
  return (((c >> 15) | (c << 17)) + MASK_DELTA) & 0xffffffff


def unmask_crc(m):
  r = (m - MASK_DELTA) & 0xffffffff
  return ((r >> 17) | (r << 15)) & 0xffffffff


# ---- varints / minimal protobuf wire format ---------------------------------------------------------------
def _varint(v):
  v &= (1 << 64) - 1          # negative int64 -> 10-byte two's complement, as protobuf does
  out = bytearray()
  while v >= 0x80:
    out.append((v & 0x7f) | 0x80)
    v >>= 7
  out.append(v)
  return bytes(out)


def _read_varint(buf, pos):
  shift = v = 0
  while True:
    b = buf[pos]
    pos += 1
    v |= (b & 0x7f) << shift
    if not b & 0x80:
      return v, pos
    shift += 7
    if shift > 63:
      raise ValueError('malformed varint')


def _parse_proto(buf):
  """-> list of (field number, wire type, value); value is int (varint / fixed) or bytes (length-delimited)."""
  out, pos, n = [], 0, len(buf)
  while pos < n:
    tag, pos = _read_varint(buf, pos)
    field, wt = tag >> 3, tag & 7
    if wt == 0:
      v, pos = _read_varint(buf, pos)
    elif wt == 1:
      v = struct.unpack_from('<Q', buf, pos)[0]
      pos += 8
    elif wt == 2:
      ln, pos = _read_varint(buf, pos)
      v = bytes(buf[pos:pos + ln])
      pos += ln
    elif wt == 5:
      v = struct.unpack_from('<I', buf, pos)[0]
      pos += 4
    else:
      raise ValueError('unsupported protobuf wire type %d' % wt)
    out.append((field, wt, v))
  return out


def _header_proto():
  version = b'\x08' + _varint(1)                               # VersionDef.producer = 1 (kTensorBundleVersion)
  return b'\x08' + _varint(1) + b'\x1a' + _varint(len(version)) + version       # num_shards = 1, LITTLE endian


def _entry_proto(dtype_enum, shape, offset, size, masked_crc):
  dims = b''
  for d in shape:
    dim = b'\x08' + _varint(int(d))
    dims += b'\x12' + _varint(len(dim)) + dim
  out = b'\x08' + _varint(dtype_enum) + b'\x12' + _varint(len(dims)) + dims
  if offset:
    out += b'\x20' + _varint(offset)
  if size:
    out += b'\x28' + _varint(size)
  return out + b'\x35' + struct.pack('<I', masked_crc)


def _parse_entry(buf):
  e = {'dtype': 0, 'shape': (), 'shard_id': 0, 'offset': 0, 'size': 0, 'crc32c': None, 'slices': 0}
  for field, wt, v in _parse_proto(buf):
    if field == 1:
      e['dtype'] = v
    elif field == 2:
      shape = []
      for f2, _, dim in _parse_proto(v):
        if f2 == 2:
          size = 0
          for f3, _, s in _parse_proto(dim):
            if f3 == 1:
              size = s - (1 << 64) if s >= (1 << 63) else s
          shape.append(size)
        elif f2 == 3 and dim:
          raise ValueError('tensor of unknown rank in a checkpoint')
      e['shape'] = tuple(shape)
    elif field == 3:
      e['shard_id'] = v
    elif field == 4:
      e['offset'] = v
    elif field == 5:
      e['size'] = v
    elif field == 6:
      e['crc32c'] = v
    elif field == 7:
      e['slices'] += 1
  return e


# ---- table (SSTable) writer ---------------------------------------------------------------------------------
class _BlockBuilder(object):
  def __init__(self, restart_interval):
    self.interval = restart_interval
    self.reset()

  def reset(self):
    self.buf, self.restarts, self.counter, self.last_key = bytearray(), [0], 0, b''

  def empty(self):
    return not self.buf

  def size_estimate(self):
    return len(self.buf) + 4 * len(self.restarts) + 4

  def add(self, key, value):
    shared = 0
    if self.counter < self.interval:
      m = min(len(key), len(self.last_key))
      while shared < m and key[shared] == self.last_key[shared]:
        shared += 1
    else:
      self.restarts.append(len(self.buf))
      self.counter = 0
    self.buf += _varint(shared) + _varint(len(key) - shared) + _varint(len(value)) + key[shared:] + value
    self.last_key = key
    self.counter += 1

  def finish(self):
    return bytes(self.buf) + b''.join(struct.pack('<I', r) for r in self.restarts) + struct.pack('<I', len(self.restarts))


def _shortest_separator(start, limit):
  """A key k with start <= k < limit, as short as the common prefix allows (LevelDB BytewiseComparator)."""
  m = min(len(start), len(limit))
  i = 0
  while i < m and start[i] == limit[i]:
    i += 1
  if i < m and start[i] < 0xff and start[i] + 1 < limit[i]:
    return start[:i] + bytes([start[i] + 1])
  return start


def _short_successor(key):
  for i, b in enumerate(key):
    if b != 0xff:
      return key[:i] + bytes([b + 1])
  return key


def _write_table(path, items):
  """items: list of (key bytes, value bytes), keys strictly increasing."""
  with open(path, 'wb') as f:
    offset = 0

    def write_block(contents):
      nonlocal offset
      trailer_type = b'\x00'
      crc = mask_crc(crc32c(trailer_type, crc32c(contents)))
      f.write(contents + trailer_type + struct.pack('<I', crc))
      handle = _varint(offset) + _varint(len(contents))
      offset += len(contents) + 5
      return handle

    data, index = _BlockBuilder(RESTART_INTERVAL), _BlockBuilder(1)
    pending = None            # (last key of the finished block, its handle)
    for key, value in items:
      if pending is not None:
        index.add(_shortest_separator(pending[0], key), pending[1])
        pending = None
      data.add(key, value)
      if data.size_estimate() >= BLOCK_SIZE:
        pending = (data.last_key, write_block(data.finish()))
        data.reset()
    if not data.empty():
      pending = (data.last_key, write_block(data.finish()))
    if pending is not None:
      index.add(_short_successor(pending[0]), pending[1])
    meta_handle = write_block(_BlockBuilder(RESTART_INTERVAL).finish())
    index_handle = write_block(index.finish())
    footer = meta_handle + index_handle
    footer += b'\x00' * (40 - len(footer)) + struct.pack('<Q', TABLE_MAGIC)
    f.write(footer)


# ---- table reader -------------------------------------------------------------------------------------------
def _read_block(buf, offset, size, verify=True):
  contents, typ = buf[offset:offset + size], buf[offset + size]
  if verify:
    want = struct.unpack_from('<I', buf, offset + size + 1)[0]
    got = mask_crc(crc32c(bytes(buf[offset + size:offset + size + 1]), crc32c(bytes(contents))))
    if want != got:
      raise IOError('checkpoint index: block checksum mismatch at offset %d' % offset)
  if typ != 0:
    raise IOError('checkpoint index: compressed table blocks (type %d) are not supported' % typ)
  return contents


def _block_entries(block):
  n_restarts = struct.unpack_from('<I', block, len(block) - 4)[0]
  end = len(block) - 4 - 4 * n_restarts
  pos, key = 0, b''
  while pos < end:
    shared, pos = _read_varint(block, pos)
    non_shared, pos = _read_varint(block, pos)
    vlen, pos = _read_varint(block, pos)
    key = key[:shared] + bytes(block[pos:pos + non_shared])
    pos += non_shared
    yield key, bytes(block[pos:pos + vlen])
    pos += vlen


def _read_table(path, verify=True):
  with open(path, 'rb') as f:
    buf = f.read()
  if len(buf) < 48 or struct.unpack_from('<Q', buf, len(buf) - 8)[0] != TABLE_MAGIC:
    raise IOError('%s is not a TensorFlow checkpoint index (bad magic number)' % path)
  footer = buf[len(buf) - 48:]
  pos = 0
  _, pos = _read_varint(footer, pos)
  _, pos = _read_varint(footer, pos)
  ioff, pos = _read_varint(footer, pos)
  isize, pos = _read_varint(footer, pos)
  out = OrderedDict()
  for _, handle in _block_entries(_read_block(buf, ioff, isize, verify)):
    boff, p = _read_varint(handle, 0)
    bsize, p = _read_varint(handle, p)
    for k, v in _block_entries(_read_block(buf, boff, bsize, verify)):
      out[k] = v
  return out


# ---- public API -----------------------------------------------------------------------------------------------
def data_filename(prefix, shard=0, num_shards=1):
  return '%s.data-%05d-of-%05d' % (prefix, shard, num_shards)


def write_checkpoint(prefix, tensors):
  """Writes `<prefix>.index` + `<prefix>.data-00000-of-00001` from {variable name: array-like}.
  float64 inputs are kept as DT_DOUBLE; pass float32 arrays for DT_FLOAT variables."""
  items = []
  for name in sorted(tensors, key=lambda s: s.encode('utf-8')):
    a = np.asarray(tensors[name])
    if a.dtype.newbyteorder('<') not in _DTYPE_ENUM and a.dtype not in _DTYPE_ENUM:
      raise TypeError('variable %r: dtype %s has no checkpoint encoding here' % (name, a.dtype))
    shape = a.shape                     # (np.ascontiguousarray promotes 0-d to 1-d: keep the scalar's shape [])
    a = np.ascontiguousarray(a.astype(a.dtype.newbyteorder('<'), copy=False)).reshape(-1)
    items.append((name, a, shape))
  os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
  table = [(b'', _header_proto())]
  tmp = data_filename(prefix) + '.tempstate'
  with open(tmp, 'wb') as f:
    offset = 0
    for name, a, shape in items:
      table.append((name.encode('utf-8'), _entry_proto(_DTYPE_ENUM[a.dtype], shape, offset, a.nbytes,
                                                       mask_crc(crc32c(a)))))
      f.write(a.tobytes())
      offset += a.nbytes
  os.replace(tmp, data_filename(prefix))
  _write_table(prefix + '.index.tempstate', table)
  os.replace(prefix + '.index.tempstate', prefix + '.index')
  return prefix


class CheckpointReader(object):
  """tf.train.NewCheckpointReader look-alike: get_variable_to_shape_map / has_tensor / get_tensor."""

  def __init__(self, prefix, verify=True):
    if not os.path.exists(prefix + '.index'):
      raise IOError('checkpoint index %s.index not found' % prefix)
    self.prefix, self.verify = prefix, verify
    raw = _read_table(prefix + '.index', verify)
    if b'' not in raw:
      raise IOError('%s.index has no bundle header' % prefix)
    hdr = {f: v for f, _, v in _parse_proto(raw[b''])}
    self.num_shards = hdr.get(1, 0)
    if hdr.get(2, 0) != 0:
      raise IOError('big-endian checkpoints are not supported')
    self.entries = OrderedDict((k.decode('utf-8'), _parse_entry(v)) for k, v in raw.items() if k != b'')
    self._files = {}

  def get_variable_to_shape_map(self):
    return {k: list(e['shape']) for k, e in self.entries.items()}

  def get_variable_to_dtype_map(self):
    return {k: _DTYPES.get(e['dtype']) for k, e in self.entries.items()}

  def has_tensor(self, name):
    return name in self.entries

  def get_tensor(self, name):
    e = self.entries[name]
    if e['slices']:
      raise NotImplementedError('partitioned variable %r' % name)
    if e['dtype'] not in _DTYPES:
      raise NotImplementedError('variable %r has DataType %d (no numeric encoding here)' % (name, e['dtype']))
    dt = _DTYPES[e['dtype']]
    fn = data_filename(self.prefix, e['shard_id'], self.num_shards)
    if fn not in self._files:
      self._files[fn] = np.memmap(fn, dtype=np.uint8, mode='r') if os.path.getsize(fn) else np.zeros(0, np.uint8)
    raw = np.asarray(self._files[fn][e['offset']:e['offset'] + e['size']])
    n = int(np.prod(e['shape'], dtype=np.int64)) if e['shape'] else 1
    if raw.nbytes != n * dt.itemsize:
      raise IOError('variable %r: %d bytes on disk, shape %s needs %d' % (name, raw.nbytes, e['shape'], n * dt.itemsize))
    if self.verify and e['crc32c'] is not None and mask_crc(crc32c(raw)) != e['crc32c']:
      raise IOError('variable %r: data checksum mismatch' % name)
    return raw.view(dt).reshape(e['shape']).copy()

  def read_all(self, skip_unsupported=True):
    """Every variable of the bundle.  Entries this reader has no decoding for (non-numeric dtypes such as the DT_STRING
    `_CHECKPOINTABLE_OBJECT_GRAPH` of later TF1 savers, partitioned variables) are skipped with a warning, or raise
    NotImplementedError with skip_unsupported=False."""
    out = OrderedDict()
    for k in self.entries:
      try:
        out[k] = self.get_tensor(k)
      except NotImplementedError as e:
        if not skip_unsupported:
          raise
        import warnings
        warnings.warn('checkpoint entry skipped: %s' % (e,))
    return out


def checkpoint_exists(fname):
  """cnn_train_multi.py:404: tf.gfile.Exists(fname) or tf.gfile.Exists(fname + '.index')."""
  return os.path.exists(fname) or os.path.exists(fname + '.index')


def update_checkpoint_state(save_dir, prefix):
  """Maintains `<save_dir>/checkpoint` as tf.train.Saver(max_to_keep=None) does (cnn_train_multi.py:439): the latest
  path plus the list of all paths, relative to save_dir."""
  rel = os.path.relpath(prefix, save_dir)
  state = os.path.join(save_dir, 'checkpoint')
  paths = []
  if os.path.exists(state):
    for line in open(state):
      if line.startswith('all_model_checkpoint_paths:'):
        paths.append(line.split(':', 1)[1].strip().strip('"'))
  paths = [p for p in paths if p != rel] + [rel]
  with open(state + '.tmp', 'w') as f:
    f.write('model_checkpoint_path: "%s"\n' % rel)
    for p in paths:
      f.write('all_model_checkpoint_paths: "%s"\n' % p)
  os.replace(state + '.tmp', state)


def latest_checkpoint(save_dir):
  """tf.train.latest_checkpoint."""
  state = os.path.join(save_dir, 'checkpoint')
  if not os.path.exists(state):
    return None
  for line in open(state):
    if line.startswith('model_checkpoint_path:'):
      p = line.split(':', 1)[1].strip().strip('"')
      p = p if os.path.isabs(p) else os.path.join(save_dir, p)
      return p if os.path.exists(p + '.index') else None
  return None
