"""TF V2 checkpoint reader/writer + variable-name map.  CPU only: libcis_b200.so is loaded for its host CRC-32C
routine, no GPU call is made.

Parity note: no TF-written checkpoint is available offline, so the format is pinned by (a) the published CRC-32C vectors,
(b) a table assembled by hand in this file, byte by byte, from the LevelDB/TensorBundle layout, (c) round trips."""
import os
import struct

import numpy as np
import pytest
import torch

from unsupervised_detection_b200.checkpoint import tf_bundle as tb
from unsupervised_detection_b200.checkpoint import tf_names as tn


# ---------------------------------------------------------------------------------------------------------------- crc32c
def test_crc32c_known_answers():
    # RFC 3720 B.4 test vectors + the classic check value
    assert tb.crc32c(b'123456789') == 0xE3069283
    assert tb.crc32c(b'\x00' * 32) == 0x8A9136AA
    assert tb.crc32c(b'\xff' * 32) == 0x62A8AB43
    assert tb.crc32c(bytes(range(32))) == 0x46DD794E
    assert tb.crc32c(bytes(range(31, -1, -1))) == 0x113FDB5C
    assert tb.crc32c(b'') == 0


def test_crc32c_extend_and_alignment():
    rng = np.random.RandomState(0)
    buf = rng.randint(0, 256, 4099).astype(np.uint8).tobytes()
    whole = tb.crc32c(buf)
    for cut in (0, 1, 7, 8, 9, 1000, 4098, 4099):
        assert tb.crc32c(buf[cut:], tb.crc32c(buf[:cut])) == whole
    # bit-serial restatement on a short prefix
    c = 0xffffffff
    for b in buf[:257]:
        c ^= b
        for _ in range(8):
            c = (c >> 1) ^ (0x82F63B78 if c & 1 else 0)
    assert tb.crc32c(buf[:257]) == c ^ 0xffffffff
    a = np.frombuffer(buf, dtype=np.uint8)[3:]                # unaligned ndarray view
    assert tb.crc32c(a) == tb.crc32c(buf[3:])


def test_crc_mask():
    c = tb.crc32c(b'foo')
    assert tb.mask_crc(c) != c and tb.unmask_crc(tb.mask_crc(c)) == c
    assert tb.mask_crc(0) == 0xa282ead8
    assert tb.mask_crc(0x00008000) == (1 + 0xa282ead8)         # rotate right by 15
    assert tb.unmask_crc(tb.mask_crc(0xffffffff)) == 0xffffffff


# ----------------------------------------------------------------------------------------------------------------- table
def _hand_block(entries, restarts):
    body = b''.join(bytes([s, len(k), len(v)]) + k + v for s, k, v in entries)
    body += b''.join(struct.pack('<I', r) for r in restarts) + struct.pack('<I', len(restarts))
    return body


def _with_trailer(body):
    return body + b'\x00' + struct.pack('<I', tb.mask_crc(tb.crc32c(body + b'\x00')))


def test_read_hand_assembled_table(tmp_path):
    """A two-block table written out byte by byte from the format description (no use of write_table)."""
    b0 = _hand_block([(0, b'apple', b'1'), (3, b'ly', b'22'), (0, b'banana', b'')], [0])      # 'app'+'ly' shares 3 bytes
    b1 = _hand_block([(0, b'cherry', b'xyz')], [0])
    meta = _hand_block([], [0])
    off0, off1 = 0, len(b0) + 5
    offm = off1 + len(b1) + 5
    offi = offm + len(meta) + 5
    idx = _hand_block([(0, b'bb', bytes([off0, len(b0)])), (0, b'd', bytes([off1, len(b1)]))], [0, 5 + 2])
    # second index entry starts after: 3 header bytes + 'bb' + 2 handle bytes = 7
    footer = bytes([offm, len(meta), offi, len(idx)])
    footer += b'\x00' * (40 - len(footer)) + struct.pack('<Q', 0xdb4775248b80fb57)
    p = tmp_path / 'hand.index'
    p.write_bytes(_with_trailer(b0) + _with_trailer(b1) + _with_trailer(meta) + _with_trailer(idx) + footer)
    assert tb.read_table(str(p)) == [(b'apple', b'1'), (b'apply', b'22'), (b'banana', b''), (b'cherry', b'xyz')]
    # corrupt one payload byte -> block checksum error
    raw = bytearray(p.read_bytes())
    raw[8] ^= 1
    p.write_bytes(bytes(raw))
    with pytest.raises(IOError, match='checksum'):
        tb.read_table(str(p))
    assert tb.read_table(str(p), verify=False)[0][0] == b'apple'


def test_write_table_bytes_of_tiny_table(tmp_path):
    """write_table's output for one entry equals the layout assembled by hand."""
    p = tmp_path / 't.index'
    tb.write_table(str(p), [(b'k', b'v')])
    b0 = _hand_block([(0, b'k', b'v')], [0])
    meta = _hand_block([], [0])
    idx = _hand_block([(0, b'k', bytes([0, len(b0)]))], [0])
    offm = len(b0) + 5
    offi = offm + len(meta) + 5
    footer = bytes([offm, len(meta), offi, len(idx)])
    footer += b'\x00' * (40 - len(footer)) + struct.pack('<Q', tb.TABLE_MAGIC)
    assert p.read_bytes() == _with_trailer(b0) + _with_trailer(meta) + _with_trailer(idx) + footer


def test_table_many_keys_multi_block(tmp_path):
    items = [(('scope/layer_%04d/kernel' % i).encode(), os.urandom(i % 37)) for i in range(500)]
    items.sort()
    p = str(tmp_path / 'm.index')
    tb.write_table(p, items, block_size=512)                   # forces dozens of blocks and restart points
    assert tb.read_table(p) == items
    with pytest.raises(ValueError):
        tb.write_table(p, [(b'b', b''), (b'a', b'')])


def test_bad_magic(tmp_path):
    p = tmp_path / 'x.index'
    p.write_bytes(b'\x00' * 64)
    with pytest.raises(IOError, match='magic'):
        tb.read_table(str(p))


def test_snappy_block_decoder():
    comp = bytes([10, 0x04, ord('a'), ord('b'), 0x11, 0x02])    # len 10; literal "ab"; copy len 8 offset 2
    assert tb._snappy_uncompress(comp) == b'ababababab'
    long_lit = bytes(range(70))
    comp = bytes([70, 60 << 2, 69]) + long_lit                 # literal with 1-byte length
    assert tb._snappy_uncompress(comp) == long_lit


# ---------------------------------------------------------------------------------------------------------------- bundle
def test_bundle_roundtrip_dtypes_and_shapes(tmp_path):
    rng = np.random.RandomState(1)
    t = {'a/kernel': rng.randn(3, 3, 5, 7).astype(np.float32), 'a/bias': np.zeros(7, np.float32),
         'train_op/global_step': np.asarray(175, np.int32), 'gs64': np.asarray(-3, np.int64), 'empty': np.zeros((0, 4), np.float32),
         'MaskNet//conv1/kernel': rng.randn(5, 5, 5, 32).astype(np.float32), 'd': rng.randn(4).astype(np.float64)}
    prefix = str(tmp_path / 'sub' / 'model-5')
    tb.write_bundle(prefix, t)
    assert tb.is_bundle(prefix) and os.path.isfile(prefix + '.data-00000-of-00001')
    r = tb.read_bundle(prefix)
    assert set(r) == set(t)
    for k in t:
        assert r[k].shape == t[k].shape and r[k].dtype == t[k].dtype and np.array_equal(r[k], t[k]), k
    lv = dict((n, (s, d)) for n, s, d in tb.list_variables(prefix))
    assert lv['a/kernel'] == ((3, 3, 5, 7), np.float32) and lv['train_op/global_step'] == ((), np.int32)
    assert set(tb.read_bundle(prefix, names=['a/bias'])) == {'a/bias'}
    # header entry: BundleHeaderProto{num_shards=1, version.producer=1}; entries sorted bytewise; data packed back to back
    items = tb.read_table(prefix + '.index')
    assert items[0] == (b'', b'\x08\x01\x1a\x02\x08\x01')
    assert [k for k, _ in items] == sorted(k for k, _ in items)
    assert os.path.getsize(prefix + '.data-00000-of-00001') == sum(v.nbytes for v in t.values())
    e = tb._parse_entry(dict(items)[b'a/kernel'])
    assert e['dtype'] == 1 and e['shape'] == (3, 3, 5, 7) and e['size'] == 3 * 3 * 5 * 7 * 4
    assert tb.unmask_crc(e['crc32c']) == tb.crc32c(t['a/kernel'])


def test_bundle_detects_corrupt_tensor(tmp_path):
    prefix = str(tmp_path / 'm')
    tb.write_bundle(prefix, {'w': np.arange(100, dtype=np.float32)})
    dp = prefix + '.data-00000-of-00001'
    raw = bytearray(open(dp, 'rb').read())
    raw[17] ^= 0x40
    open(dp, 'wb').write(bytes(raw))
    with pytest.raises(IOError, match='checksum'):
        tb.read_bundle(prefix)
    open(dp, 'wb').write(bytes(raw[:100]))
    with pytest.raises(IOError, match='truncated'):
        tb.read_bundle(prefix, verify=False)


def test_bfloat16_entries_are_widened(tmp_path):
    prefix = str(tmp_path / 'bf')
    vals = np.array([1.0, -2.5, 0.15625], np.float32)
    bits = (vals.view(np.uint32) >> 16).astype(np.uint16)
    with open(prefix + '.data-00000-of-00001', 'wb') as f:
        f.write(bits.tobytes())
    ent = tb._encode_entry(14, (3,), 0, 6, tb.mask_crc(tb.crc32c(bits)))
    tb.write_table(prefix + '.index', [(b'', b'\x08\x01\x1a\x02\x08\x01'), (b'x', ent)])
    r = tb.read_bundle(prefix)
    assert r['x'].dtype == np.float32 and np.array_equal(r['x'], vals)


def test_checkpoint_state_file(tmp_path):
    d = str(tmp_path)
    assert tb.latest_checkpoint(d) is None
    for i in range(5):
        tb.write_bundle(os.path.join(d, 'model-%d' % i), {'w': np.zeros(1, np.float32)})
        dropped = tb.update_checkpoint_state(d, 'model-%d' % i, keep=3)
    assert dropped == ['model-1']
    txt = open(os.path.join(d, 'checkpoint')).read().splitlines()
    assert txt[0] == 'model_checkpoint_path: "model-4"'
    assert txt[1:] == ['all_model_checkpoint_paths: "model-%d"' % i for i in (2, 3, 4)]
    assert tb.latest_checkpoint(d) == os.path.join(d, 'model-4')


# ----------------------------------------------------------------------------------------------------------------- names
def test_generator_tf_names():
    from unsupervised_detection_b200.models.nets import GEN_LAYERS
    assert tuple(l[0] for l in GEN_LAYERS) == tn.GEN_LAYER_NAMES
    m = tn.generator_tf_names()
    assert m['MaskNet/conv1/kernel'] == 'MaskNet//conv1/kernel'
    assert m['MaskNet/conv1/gamma'] == 'MaskNet//batch_normalization/gamma'
    assert m['MaskNet/conv2_downsample/beta'] == 'MaskNet//batch_normalization_1/beta'
    assert m['MaskNet/conv12/gamma'] == 'MaskNet//batch_normalization_11/gamma'
    assert m['MaskNet/conv13_upsample/kernel'] == 'MaskNet//conv13_upsample/conv13_upsample_conv/kernel'
    assert m['MaskNet/conv13_upsample/gamma'] == 'MaskNet//conv13_upsample/batch_normalization/gamma'
    assert m['MaskNet/conv14/gamma'] == 'MaskNet//batch_normalization_12/gamma'
    assert m['MaskNet/conv15_upsample/bias'] == 'MaskNet//conv15_upsample/conv15_upsample_conv/bias'
    assert m['MaskNet/conv17/beta'] == 'MaskNet//batch_normalization_14/beta'
    assert len(set(m.values())) == len(m) == 17 * 4
    assert tn.to_tf_name('FlownetS/aconv1/weights') == 'FlownetS//aconv1/weights'
    assert tn.to_tf_name('FlownetS/flow1/biases', '/') == 'FlownetS/flow1/biases'
    assert tn.to_tf_name('pwcnet/ctxt/dc_conv21/kernel') == 'pwcnet/ctxt/dc_conv21/kernel'


def test_import_accepts_both_scope_spellings_and_reports_missing():
    from oracle.params import make_params
    p = make_params(3)
    for sep in ('//', '/'):
        tfv = tn.export_params(p, global_step=42, sep=sep)
        tfv['pwcnet/featpyr/conv1a/kernel/Adam'] = np.zeros(1, np.float32)        # optimizer slots are ignored
        got, gs = tn.import_params(tfv, list(p))
        assert gs == 42 and set(got) == set(p)
        assert all(np.array_equal(got[k], p[k].numpy()) for k in p)
    tfv.pop('MaskNet/batch_normalization_3/gamma')
    with pytest.raises(KeyError, match='conv4_downsample/gamma'):
        tn.import_params(tfv, list(p))
    got, _ = tn.import_params(tfv, list(p), strict=False)
    assert len(got) == len(p) - 1
    _, gs = tn.import_params({'global_step': np.asarray(7, np.int64)}, [])
    assert gs == 7


def test_normalize_prefix():
    assert tn.normalize_prefix('/a/pwcnet.ckpt-595000.data-00000-of-00001') == '/a/pwcnet.ckpt-595000'
    assert tn.normalize_prefix('/a/model.best.index') == '/a/model.best'
    assert tn.normalize_prefix('/a/model-175') == '/a/model-175'


# --------------------------------------------------------------------------------------------------------------- learner
class _StubStore:
    def __init__(self, names):
        self.entries = [(n, None, 0, 0, None) for n in names]


class _StubGraph:
    def __init__(self, params):
        self.p = params
        self.gen_store = _StubStore([k for k in params if k.startswith('MaskNet/')])
        self.rec_store = _StubStore([k for k in params if k.startswith('FlownetS/')])
        self.pwc_store = _StubStore([k for k in params if k.startswith('pwcnet/')])

    def export_params(self):
        return self.p


def test_learner_save_writes_saver_layout_and_reads_back(tmp_path, capsys):
    from oracle.params import make_params
    from unsupervised_detection_b200.models.adversarial_learner import AdversarialLearner
    p = make_params(5)
    L = object.__new__(AdversarialLearner)
    L.rank, L.global_step, L.graph = 0, 350, _StubGraph(p)
    d = str(tmp_path / 'ckpts')
    L.save(None, d, 5)
    L.save(None, d, 'best')
    files = sorted(os.listdir(d))
    assert files == ['checkpoint', 'model-5.data-00000-of-00001', 'model-5.index', 'model-5.pt', 'model.best.data-00000-of-00001',
                     'model.best.index', 'model.best.pt']
    names = [n for n, _, _ in tb.list_variables(os.path.join(d, 'model-5'))]
    assert 'train_op/global_step' in names and 'MaskNet//batch_normalization_14/gamma' in names
    assert 'FlownetS//deconv5/weights' in names and 'pwcnet/upsample/up_feat2/kernel' not in names   # make_params has no level-2 upsampler
    assert len(names) == len(p) + 1
    assert AdversarialLearner._latest_checkpoint(d) == os.path.join(d, 'model.best')
    for path in (os.path.join(d, 'model-5'), os.path.join(d, 'model-5.index'), os.path.join(d, 'model-5.data-00000-of-00001'),
                 os.path.join(d, 'model-5.pt')):
        assert AdversarialLearner._is_ckpt(path)
        got, gs = AdversarialLearner._read_ckpt(path, L._names('MaskNet', 'FlownetS', 'pwcnet'))
        assert gs == 350 and set(got) == set(p)
        assert all(torch.equal(got[k], p[k]) for k in p)
    got, _ = AdversarialLearner._read_ckpt(os.path.join(d, 'model-5'), L._names('FlownetS'))
    assert set(got) == set(k for k in p if k.startswith('FlownetS/'))
    assert not AdversarialLearner._is_ckpt(os.path.join(d, 'model-6')) and not AdversarialLearner._is_ckpt('')
    with pytest.raises(KeyError):
        AdversarialLearner._read_ckpt(os.path.join(d, 'model-5'), ['MaskNet/conv1/kernel', 'pwcnet/nope/kernel'])


# ----------------------------------------------------------------------------------------------------------- property tests
from hypothesis import given, settings, strategies as st   # noqa: E402


@settings(max_examples=60, deadline=None)
@given(st.lists(st.integers(min_value=0, max_value=(1 << 64) - 1), min_size=1, max_size=20))
def test_varint_roundtrip(values):
    buf = bytearray()
    for v in values:
        tb._put_varint(buf, v)
    pos, out = 0, []
    for _ in values:
        v, pos = tb._get_varint(buf, pos)
        out.append(v)
    assert out == values and pos == len(buf)


@settings(max_examples=25, deadline=None)
@given(st.dictionaries(st.binary(min_size=0, max_size=40), st.binary(min_size=0, max_size=200), min_size=1, max_size=60),
       st.sampled_from([64, 300, 4096]))
def test_table_roundtrip_random_keys(tmp_path_factory, kv, block_size):
    items = sorted(kv.items())
    p = str(tmp_path_factory.mktemp('tbl') / 'r.index')
    tb.write_table(p, items, block_size=block_size)
    assert tb.read_table(p) == items


@settings(max_examples=20, deadline=None)
@given(st.lists(st.tuples(st.sampled_from([np.float32, np.int32, np.int64, np.float64]),
                          st.lists(st.integers(min_value=0, max_value=5), min_size=0, max_size=4)), min_size=1, max_size=8),
       st.integers(min_value=0, max_value=2 ** 31 - 1))
def test_bundle_roundtrip_random_tensors(tmp_path_factory, specs, seed):
    rng = np.random.RandomState(seed)
    t = {}
    for i, (dt, shape) in enumerate(specs):
        a = (rng.randn(*shape) * 100).astype(dt) if shape else np.asarray(rng.randint(-1000, 1000), dtype=dt)
        t['scope_%d/var:%d' % (i % 3, i)] = a
    prefix = str(tmp_path_factory.mktemp('bndl') / 'ck')
    tb.write_bundle(prefix, t)
    r = tb.read_bundle(prefix)
    assert set(r) == set(t)
    for k in t:
        assert r[k].dtype == t[k].dtype and r[k].shape == t[k].shape and np.array_equal(r[k], t[k])


# ------------------------------------------------------------------------------------------- foreign bytes: a whole bundle by hand
def _py_crc32c(data):
    """Bit-by-bit CRC-32C (Castagnoli, reflected polynomial 0x82F63B78) -- deliberately independent of libcis_b200's table-driven
    routine and of every encoder in checkpoint/tf_bundle.py."""
    crc = 0xffffffff
    for b in bytes(data):
        crc ^= b
        for _ in range(8):
            crc = (crc >> 1) ^ (0x82F63B78 & -(crc & 1))
    return crc ^ 0xffffffff


def _py_mask(c):
    return ((((c >> 15) | (c << 17)) & 0xffffffff) + 0xa282ead8) & 0xffffffff


def _pb_varint(v):
    out = bytearray()
    while True:
        b = v & 0x7f
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _pb_field(num, wt, payload):
    return _pb_varint((num << 3) | wt) + payload


def _entry_proto(dtype, dims, offset, size, crc):
    # BundleEntryProto: dtype = 1 (varint), shape = 2 (TensorShapeProto: repeated Dim dim = 2 {int64 size = 1}), shard_id = 3,
    # offset = 4, size = 5, crc32c = 6 (fixed32); proto3 omits zero-valued scalars, as TF's serializer does
    shape = b''.join(_pb_field(2, 2, _pb_varint(len(d)) + d) for d in (_pb_field(1, 0, _pb_varint(n)) for n in dims))
    out = _pb_field(1, 0, _pb_varint(dtype)) + _pb_field(2, 2, _pb_varint(len(shape)) + shape)
    if offset:
        out += _pb_field(4, 0, _pb_varint(offset))
    out += _pb_field(5, 0, _pb_varint(size)) + _pb_field(6, 5, struct.pack('<I', crc))
    return out


def _foreign_block(entries):
    """LevelDB data block with shared-prefix compression and a restart point every 2 entries (TF uses 16; any interval is legal)."""
    body, restarts, prev = bytearray(), [], b''
    for i, (k, v) in enumerate(entries):
        shared = 0
        if i % 2 == 0:
            restarts.append(len(body))
        else:
            while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                shared += 1
        body += _pb_varint(shared) + _pb_varint(len(k) - shared) + _pb_varint(len(v)) + k[shared:] + v
        prev = k
    for r in restarts:
        body += struct.pack('<I', r)
    body += struct.pack('<I', len(restarts))
    return bytes(body)


def _foreign_trailer(body):
    return body + b'\x00' + struct.pack('<I', _py_mask(_py_crc32c(body + b'\x00')))


def test_read_a_bundle_assembled_by_hand_from_the_published_format(tmp_path):
    """`.index` + `.data-00000-of-00001` built byte by byte here (own varints, protos, block builder with prefix compression, own bitwise
    CRC-32C) -- nothing from write_bundle / write_table / cis_crc32c -- with the reference's variable naming (`MaskNet//...` double
    slash, adversarial_learner.py:211-214; `global_step`), two data blocks, non-zero offsets: the reader and the name map must take it."""
    rng = np.random.RandomState(3)
    tensors = [('FlownetS//aconv1/biases', rng.randn(16).astype(np.float32)),
               ('MaskNet//conv1/bias', rng.randn(32).astype(np.float32)),
               ('MaskNet//conv1/kernel', rng.randn(5, 5, 5, 32).astype(np.float32)),
               ('global_step', np.asarray(1234, dtype=np.int64))]
    data, entries = bytearray(), []
    for name, arr in tensors:                       # keys are already in sorted (byte) order, as a table requires
        raw = arr.tobytes()
        dt = 1 if arr.dtype == np.float32 else 9
        entries.append((name.encode(), _entry_proto(dt, list(arr.shape), len(data), len(raw), _py_mask(_py_crc32c(raw)))))
        data += raw
    header = _pb_field(1, 0, _pb_varint(1)) + _pb_field(3, 2, _pb_varint(2) + _pb_field(1, 0, _pb_varint(1)))   # num_shards = 1, version {producer 1}
    b0 = _foreign_block([(b'', header)] + entries[:2])
    b1 = _foreign_block(entries[2:])
    meta = _foreign_block([])
    off1 = len(b0) + 5
    offm = off1 + len(b1) + 5
    offi = offm + len(meta) + 5
    handle = lambda o, n: _pb_varint(o) + _pb_varint(n)
    # index block: one entry per data block, key >= last key of the block (the separator TF/LevelDB would shorten; any such key is legal)
    idx = _foreign_block([(entries[1][0] + b'\x00', handle(0, len(b0))), (b'h', handle(off1, len(b1)))])
    footer = handle(offm, len(meta)) + handle(offi, len(idx))
    footer += b'\x00' * (40 - len(footer)) + struct.pack('<Q', 0xdb4775248b80fb57)
    prefix = str(tmp_path / 'model-1234')
    with open(prefix + '.index', 'wb') as f:
        f.write(_foreign_trailer(b0) + _foreign_trailer(b1) + _foreign_trailer(meta) + _foreign_trailer(idx) + footer)
    with open(prefix + '.data-00000-of-00001', 'wb') as f:
        f.write(bytes(data))
    assert tb.is_bundle(prefix)
    got = tb.read_bundle(prefix)
    assert sorted(got) == sorted(n for n, _ in tensors)
    for name, arr in tensors:
        assert got[name].dtype == arr.dtype and got[name].shape == arr.shape and np.array_equal(got[name], arr), name
    # the independent CRC agrees with the library routine the reader verifies with
    assert _py_crc32c(b'123456789') == 0xE3069283 == tb.crc32c(b'123456789')
    # name map: the internal (single-slash) names of this repo pick the double-slash TF variables
    from unsupervised_detection_b200.checkpoint import tf_names
    picked, gs = tf_names.import_params(got, ['MaskNet/conv1/kernel', 'MaskNet/conv1/bias', 'FlownetS/aconv1/biases'])
    assert gs == 1234 and np.array_equal(picked['MaskNet/conv1/kernel'], tensors[2][1])
    # a flipped payload byte in the data shard is caught by the per-tensor checksum written by the foreign writer
    raw = bytearray(open(prefix + '.data-00000-of-00001', 'rb').read())
    raw[70] ^= 0x10
    open(prefix + '.data-00000-of-00001', 'wb').write(bytes(raw))
    with pytest.raises(IOError, match='checksum'):
        tb.read_bundle(prefix)
