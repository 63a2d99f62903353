"""TensorBundle checkpoint reader/writer (cnn_train_multi.py:404-439,511-513, eval_imm.py:81-94).
CPU-only: exercises the C-ABI host function immb_crc32c and the pure-Python table / protobuf encoders."""
import os
import struct

import numpy as np
import pytest

from imm_b200.utils import tf_checkpoint as T


def test_crc32c_known_answers():
  # RFC 3720 B.4 test vectors for CRC-32C (Castagnoli)
  assert T.crc32c(b'123456789') == 0xe3069283
  assert T.crc32c(bytes(32)) == 0x8a9136aa
  assert T.crc32c(bytes([0xff] * 32)) == 0x62a8ab43
  assert T.crc32c(bytes(range(32))) == 0x46dd794e
  assert T.crc32c(bytes(range(31, -1, -1))) == 0x113fdb5c
  # continuation + unaligned starts
  d = os.urandom(1001)
  for cut in (0, 1, 7, 8, 500, 1001):
    assert T.crc32c(d[cut:], T.crc32c(d[:cut])) == T.crc32c(d)
  a = np.frombuffer(d, dtype=np.uint8)
  assert T.crc32c(a[3:]) == T.crc32c(d[3:])
  assert T.unmask_crc(T.mask_crc(0xdeadbeef)) == 0xdeadbeef
  # the same masked checksum TensorBoard's event-file writer computes (independent implementation)
  from tensorboard.compat.tensorflow_stub.pywrap_tensorflow import masked_crc32c
  assert T.mask_crc(T.crc32c(d)) == masked_crc32c(d)


def test_index_file_bytes_known_answer(tmp_path):
  """One variable a = float32 [1, 2]: the index file spelled out byte by byte from the format definition."""
  p = T.write_checkpoint(str(tmp_path / 'm.ckpt-0'), {'a': np.array([1.0, 2.0], np.float32)})
  data = open(T.data_filename(p), 'rb').read()
  assert data == struct.pack('<2f', 1.0, 2.0)
  header = bytes([0x08, 0x01, 0x1a, 0x02, 0x08, 0x01])
  entry = bytes([0x08, 0x01, 0x12, 0x04, 0x12, 0x02, 0x08, 0x02, 0x28, 0x08, 0x35]) + \
      struct.pack('<I', T.mask_crc(T.crc32c(data)))
  block = bytes([0, 0, len(header)]) + header + bytes([0, 1, len(entry)]) + b'a' + entry + struct.pack('<II', 0, 1)

  def with_trailer(b):
    return b + b'\x00' + struct.pack('<I', T.mask_crc(T.crc32c(b + b'\x00')))
  meta = struct.pack('<II', 0, 1)
  index_block = bytes([0, 1, 2]) + b'b' + bytes([0, len(block)]) + struct.pack('<II', 0, 1)     # successor('a') = 'b'
  off_meta = len(block) + 5
  off_index = off_meta + len(meta) + 5
  footer = bytes([off_meta, len(meta), off_index, len(index_block)])
  footer += bytes(40 - len(footer)) + bytes([0x57, 0xfb, 0x80, 0x8b, 0x24, 0x75, 0x47, 0xdb])
  want = with_trailer(block) + with_trailer(meta) + with_trailer(index_block) + footer
  assert open(p + '.index', 'rb').read() == want


def test_round_trip_many_variables_multi_block(tmp_path):
  rng = np.random.default_rng(0)
  tens = {'model/image_encoder/encoder/conv_1/conv_1/w': rng.normal(size=(7, 7, 3, 32)).astype(np.float32),
          'global_step': np.float32(41), 'beta1_power': np.float32(0.9 ** 43), 'counter': np.arange(5, dtype=np.int64),
          'empty': np.zeros((0, 4), np.float32), 'd': rng.normal(size=(3,)), 'flag': np.array([True, False])}
  for i in range(12000):           # > 256 KiB of index entries -> several data blocks, prefix-compressed keys
    tens['SelfSupReconstructionLoss/pad/%05d/Adam_1' % i] = rng.normal(size=(i % 5,)).astype(np.float32)
  p = T.write_checkpoint(str(tmp_path / 'model.ckpt-41'), tens)
  assert os.path.getsize(p + ".index") > T.BLOCK_SIZE + 4096
  r = T.CheckpointReader(p)
  assert set(r.get_variable_to_shape_map()) == set(tens)
  assert r.get_variable_to_shape_map()['global_step'] == [] and r.has_tensor('counter') and not r.has_tensor('nope')
  assert list(r.entries) == sorted(tens, key=lambda s: s.encode())          # bytewise key order
  for k, v in tens.items():
    got = r.get_tensor(k)
    assert got.dtype == np.asarray(v).dtype and got.shape == np.asarray(v).shape and np.array_equal(got, v), k
  # tensors lie back to back in key order in the data file
  off = 0
  for k, e in r.entries.items():
    assert e['offset'] == off and e['shard_id'] == 0
    off += e['size']
  assert off == os.path.getsize(T.data_filename(p))


def test_corruption_is_detected(tmp_path):
  p = T.write_checkpoint(str(tmp_path / 'c'), {'a': np.arange(100, dtype=np.float32), 'b': np.ones(3, np.float32)})
  raw = bytearray(open(T.data_filename(p), 'rb').read())
  raw[17] ^= 0x40
  open(T.data_filename(p), 'wb').write(bytes(raw))
  r = T.CheckpointReader(p)
  with pytest.raises(IOError, match='checksum'):
    r.get_tensor('a')
  assert np.array_equal(r.get_tensor('b'), np.ones(3, np.float32))
  assert T.CheckpointReader(p, verify=False).get_tensor('a').shape == (100,)
  idx = bytearray(open(p + '.index', 'rb').read())
  idx[5] ^= 1
  open(p + '.index', 'wb').write(bytes(idx))
  with pytest.raises(IOError, match='checksum'):
    T.CheckpointReader(p)
  open(p + '.index', 'wb').write(b'not a table at all, definitely not forty-eight bytes of footer......')
  with pytest.raises(IOError, match='magic'):
    T.CheckpointReader(p)
  with pytest.raises(IOError, match='not found'):
    T.CheckpointReader(str(tmp_path / 'missing'))


def test_checkpoint_state_file(tmp_path):
  d = str(tmp_path)
  assert T.latest_checkpoint(d) is None
  for step in (0, 2, 4, 2):
    p = T.write_checkpoint(os.path.join(d, 'model.ckpt-%d' % step), {'global_step': np.float32(step)})
    T.update_checkpoint_state(d, p)
  lines = open(os.path.join(d, 'checkpoint')).read().splitlines()
  assert lines[0] == 'model_checkpoint_path: "model.ckpt-2"'
  assert lines[1:] == ['all_model_checkpoint_paths: "model.ckpt-%d"' % s for s in (0, 4, 2)]     # max_to_keep=None
  assert T.latest_checkpoint(d) == os.path.join(d, 'model.ckpt-2')
  assert T.checkpoint_exists(os.path.join(d, 'model.ckpt-4')) and not T.checkpoint_exists(os.path.join(d, 'model.ckpt-9'))


def test_reader_accepts_foreign_writer_choices(tmp_path):
  """A TensorFlow-written index may use other (equally valid) separators, restart intervals and explicit default
  fields; the reader must not depend on this writer's choices."""
  data = np.arange(6, dtype=np.float32)
  open(T.data_filename(str(tmp_path / 'f')), 'wb').write(b'\x00' * 4 + data.tobytes())
  hdr = b'\x08\x01\x10\x00\x1a\x04\x08\x01\x10\x00'                       # endianness / min_consumer written explicitly
  ent = b'\x08\x01\x12\x08\x12\x02\x08\x02\x12\x02\x08\x03\x18\x00\x20\x04\x28\x18\x35' + \
      struct.pack('<I', T.mask_crc(T.crc32c(data)))
  b0 = T._BlockBuilder(1)                   # restart at every key
  b0.add(b'', hdr)
  b1 = T._BlockBuilder(1)
  b1.add(b'w', ent)
  path = str(tmp_path / 'f.index')
  with open(path, 'wb') as f:
    off, handles = 0, []
    for blk in (b0.finish(), b1.finish()):
      f.write(blk + b'\x00' + struct.pack('<I', T.mask_crc(T.crc32c(blk + b'\x00'))))
      handles.append(T._varint(off) + T._varint(len(blk)))
      off += len(blk) + 5
    ib = T._BlockBuilder(1)
    ib.add(b'\x00', handles[0])              # any separator in ['', 'w')
    ib.add(b'w\xff\xff', handles[1])         # any key >= 'w'
    meta = T._BlockBuilder(16).finish()
    f.write(meta + b'\x00' + struct.pack('<I', T.mask_crc(T.crc32c(meta + b'\x00'))))
    mh = T._varint(off) + T._varint(len(meta))
    off += len(meta) + 5
    iblk = ib.finish()
    f.write(iblk + b'\x00' + struct.pack('<I', T.mask_crc(T.crc32c(iblk + b'\x00'))))
    foot = mh + T._varint(off) + T._varint(len(iblk))
    f.write(foot + bytes(40 - len(foot)) + struct.pack('<Q', T.TABLE_MAGIC))
  r = T.CheckpointReader(str(tmp_path / 'f'))
  assert r.get_variable_to_shape_map() == {'w': [2, 3]}
  assert np.array_equal(r.get_tensor('w'), data.reshape(2, 3))


def test_round_trip_property_random_variable_sets(tmp_path):
  """Randomised round trips (names with shared prefixes, scalars, empty tensors, several dtypes, enough entries to
  cross restart intervals): what is written is what is read, for any variable set."""
  from hypothesis import given, settings, strategies as st, HealthCheck
  dtypes = [np.float32, np.float64, np.int32, np.int64, np.uint8, np.bool_]
  name = st.text(alphabet='abcxyz/_01', min_size=1, max_size=24)
  shape = st.lists(st.integers(0, 5), min_size=0, max_size=3)
  counter = {'n': 0}

  @settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
  @given(st.dictionaries(name, st.tuples(shape, st.integers(0, len(dtypes) - 1), st.integers(0, 2 ** 31 - 1)),
                         min_size=1, max_size=40))
  def run(spec):
    counter['n'] += 1
    tens = {}
    for k, (shp, di, seed) in spec.items():
      rng = np.random.default_rng(seed)
      tens[k] = (rng.integers(0, 2, size=shp).astype(dtypes[di]) if dtypes[di] in (np.bool_, np.uint8)
                 else (rng.normal(size=shp) * 100).astype(dtypes[di]))
    p = T.write_checkpoint(str(tmp_path / ('c%d' % counter['n'])), tens)
    r = T.CheckpointReader(p)
    assert list(r.entries) == sorted(tens, key=lambda s: s.encode())
    for k, v in tens.items():
      got = r.get_tensor(k)
      assert got.dtype == v.dtype and got.shape == v.shape and np.array_equal(got, v)
  run()


def test_read_all_skips_entries_without_a_numeric_decoding(tmp_path):
  """A bundle may carry entries this reader cannot decode (DT_STRING object graphs of later TF1 savers, partitioned
  variables): read_all() skips them with a warning instead of making the whole checkpoint unreadable, get_tensor() and
  read_all(skip_unsupported=False) still refuse them by name."""
  import warnings
  prefix = str(tmp_path / 'model.ckpt-7')
  T.write_checkpoint(prefix, {'a/w': np.arange(6, dtype=np.float32).reshape(2, 3), 'global_step': np.float32(7)})
  reader = T.CheckpointReader(prefix)
  reader.entries['_CHECKPOINTABLE_OBJECT_GRAPH'] = dict(reader.entries['a/w'], dtype=7)      # DT_STRING
  with warnings.catch_warnings(record=True) as w:
    warnings.simplefilter('always')
    got = reader.read_all()
  assert sorted(got) == ['a/w', 'global_step'] and any('_CHECKPOINTABLE_OBJECT_GRAPH' in str(x.message) for x in w)
  with pytest.raises(NotImplementedError):
    reader.read_all(skip_unsupported=False)
  with pytest.raises(NotImplementedError):
    reader.get_tensor('_CHECKPOINTABLE_OBJECT_GRAPH')
