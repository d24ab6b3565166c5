"""TensorFlow V2 checkpoint ("tensor bundle") reader / writer without TensorFlow  (SURVEY 8f-1).

The reference saves and restores with `tf.train.Saver` (models/adversarial_learner.py:300-310, 326-360; test_generator.py:45-58),
i.e. a pair of files `<prefix>.index` + `<prefix>.data-00000-of-00001` plus a text `checkpoint` state file.  TensorFlow 1.13 is a
third-party dependency of the reference (environment.yml) that is not vendored and not installable here, so this module restates
the published on-disk format:

  * `<prefix>.index` is an immutable sorted string table (tensorflow/core/lib/io/table_builder.cc, block_builder.cc, format.cc --
    the LevelDB table format): data blocks of prefix-compressed entries with a restart array, each followed by a 5-byte trailer
    (compression type, masked CRC-32C), then a meta-index block, an index block and a 48-byte footer ending in the magic
    0xdb4775248b80fb57.  TF writes bundle indexes uncompressed; snappy blocks are still decoded for robustness.
  * key "" holds a `BundleHeaderProto` (num_shards, endianness, version); every other key is a variable name whose value is a
    `BundleEntryProto` {dtype=1, shape=2, shard_id=3, offset=4, size=5, crc32c=6 (fixed32, masked), slices=7}
    (tensorflow/core/protobuf/tensor_bundle.proto, tensorflow/core/util/tensor_bundle/tensor_bundle.cc).
  * the data shard is the raw little-endian tensor bytes at [offset, offset+size).

PARITY UNPINNED: no TF-written checkpoint exists in this environment (no TensorFlow, no network, the authors' files are a
separate download), so the byte-level format is checked only against this module's own writer, the CRC-32C known-answer vectors,
TensorBoard's independent masked-CRC implementation (tests/test_summary_tensorboard.py) and a hand-assembled table (tests/test_tf_bundle.py).  The first real `model.best` / `pwcnet.ckpt-595000` that is loaded should be
treated as the pinning test.
"""
import ctypes as C
import os
import struct

import numpy as np

from .. import _lib

TABLE_MAGIC = 0xdb4775248b80fb57
FOOTER_LEN = 48
BLOCK_TRAILER = 5
RESTART_INTERVAL = 16
BLOCK_SIZE = 256 * 1024          # tensorflow/core/lib/io/table_options.h default block_size
MASK_DELTA = 0xa282ead8

# tensorflow/core/framework/types.proto
_DT = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_, 17: np.uint16,
       19: np.float16, 22: np.uint32, 23: np.uint64}
_DT_BFLOAT16 = 14
_NP2DT = {np.dtype(v): k for k, v in _DT.items()}


# ------------------------------------------------------------------------------------------------------------------- crc
def crc32c(data, crc=0):
    """CRC-32C of `data` (bytes / bytearray / contiguous ndarray), computed by libcis_b200's host routine."""
    lib = _lib.load()
    if isinstance(data, np.ndarray):
        a = np.ascontiguousarray(data)
        return int(lib.cis_crc32c(crc, a.ctypes.data_as(C.c_void_p), a.nbytes))
    b = bytes(data)
    return int(lib.cis_crc32c(crc, C.cast(C.c_char_p(b), C.c_void_p), len(b)))


def mask_crc(c):
    """tensorflow/core/lib/hash/crc32c.h Mask(): rotate right by 15 and add a constant."""
    return (((c >> 15) | (c << 17)) + MASK_DELTA) & 0xffffffff


def unmask_crc(m):
    r = (m - MASK_DELTA) & 0xffffffff
    return ((r >> 17) | (r << 15)) & 0xffffffff


# ------------------------------------------------------------------------------------------------------------- varint / proto
def _put_varint(out, v):
    v &= (1 << 64) - 1
    while v >= 0x80:
        out.append((v & 0x7f) | 0x80)
        v >>= 7
    out.append(v)


def _get_varint(buf, pos):
    r, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        r |= (b & 0x7f) << shift
        if b < 0x80:
            return r, pos
        shift += 7
        if shift > 63:
            raise ValueError('malformed varint')


def _parse_proto(buf):
    """Minimal protobuf wire parser -> list of (field, wiretype, value)."""
    out, pos, n = [], 0, len(buf)
    while pos < n:
        tag, pos = _get_varint(buf, pos)
        f, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from('<Q', buf, pos)[0]
            pos += 8
        elif wt == 2:
            ln, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + ln])
            pos += ln
        elif wt == 5:
            v = struct.unpack_from('<I', buf, pos)[0]
            pos += 4
        else:
            raise ValueError('unsupported protobuf wire type %d' % wt)
        out.append((f, wt, v))
    return out


def _signed64(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def _parse_shape(buf):
    """TensorShapeProto: repeated Dim dim = 2 {int64 size = 1; string name = 2}; bool unknown_rank = 3."""
    dims = []
    for f, wt, v in _parse_proto(buf):
        if f == 2 and wt == 2:
            size = 0
            for g, gw, gv in _parse_proto(v):
                if g == 1 and gw == 0:
                    size = _signed64(gv)
            dims.append(size)
        elif f == 3 and v:
            raise ValueError('tensor of unknown rank in checkpoint')
    return tuple(dims)


def _encode_shape(shape):
    out = bytearray()
    for d in shape:
        dim = bytearray()
        if d != 0:
            dim.append(0x08)
            _put_varint(dim, int(d))
        out.append(0x12)
        _put_varint(out, len(dim))
        out += dim
    return bytes(out)


def _encode_entry(dtype, shape, offset, size, crc_masked):
    out = bytearray()
    out.append(0x08)
    _put_varint(out, dtype)
    sh = _encode_shape(shape)
    out.append(0x12)
    _put_varint(out, len(sh))
    out += sh
    # shard_id = 0 is the proto3 default and is omitted, like TF's serializer does
    if offset:
        out.append(0x20)
        _put_varint(out, offset)
    if size:
        out.append(0x28)
        _put_varint(out, size)
    out.append(0x35)
    out += struct.pack('<I', crc_masked)
    return bytes(out)


def _parse_entry(buf):
    e = {'dtype': 0, 'shape': (), 'shard_id': 0, 'offset': 0, 'size': 0, 'crc32c': None, 'slices': 0}
    for f, wt, v in _parse_proto(buf):
        if f == 1:
            e['dtype'] = v
        elif f == 2:
            e['shape'] = _parse_shape(v)
        elif f == 3:
            e['shard_id'] = v
        elif f == 4:
            e['offset'] = v
        elif f == 5:
            e['size'] = v
        elif f == 6:
            e['crc32c'] = v
        elif f == 7:
            e['slices'] += 1
    return e


# -------------------------------------------------------------------------------------------------------------------- snappy
def _snappy_uncompress(buf):
    """Raw snappy block decoder (format_description.txt of google/snappy); only needed if an index was written compressed."""
    n, pos = _get_varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[pos:pos + nb], 'little')
                pos += nb
            ln += 1
            out += buf[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | buf[pos]
            pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = buf[pos] | (buf[pos + 1] << 8)
            pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 4], 'little')
            pos += 4
        if off == 0 or off > len(out):
            raise ValueError('corrupt snappy block')
        for _ in range(ln):
            out.append(out[-off])
    if len(out) != n:
        raise ValueError('snappy length mismatch')
    return bytes(out)


# --------------------------------------------------------------------------------------------------------------------- table
def _read_block(f, offset, size, verify=True):
    f.seek(offset)
    raw = f.read(size + BLOCK_TRAILER)
    if len(raw) != size + BLOCK_TRAILER:
        raise IOError('truncated table block')
    body, ctype = raw[:size], raw[size]
    if verify:
        want = unmask_crc(struct.unpack_from('<I', raw, size + 1)[0])
        if crc32c(raw[:size + 1]) != want:
            raise IOError('block checksum mismatch in checkpoint index')
    if ctype == 1:
        body = _snappy_uncompress(body)
    elif ctype != 0:
        raise IOError('unknown block compression type %d' % ctype)
    return body


def _block_entries(body):
    """Yield (key, value) of one block (block_builder.cc layout: entries, restart offsets u32[], num_restarts u32)."""
    if len(body) < 4:
        raise IOError('bad table block')
    nrestart = struct.unpack_from('<I', body, len(body) - 4)[0]
    end = len(body) - 4 - 4 * nrestart
    if end < 0:
        raise IOError('bad restart array')
    pos, key = 0, b''
    while pos < end:
        shared, pos = _get_varint(body, pos)
        non_shared, pos = _get_varint(body, pos)
        vlen, pos = _get_varint(body, pos)
        key = key[:shared] + bytes(body[pos:pos + non_shared])
        pos += non_shared
        yield key, bytes(body[pos:pos + vlen])
        pos += vlen


def _decode_handle(buf, pos=0):
    off, pos = _get_varint(buf, pos)
    size, pos = _get_varint(buf, pos)
    return off, size, pos


def read_table(path, verify=True):
    """All (key, value) pairs of a table file in key order."""
    out = []
    with open(path, 'rb') as f:
        f.seek(0, os.SEEK_END)
        n = f.tell()
        if n < FOOTER_LEN:
            raise IOError('%s: too short to be a table file' % path)
        f.seek(n - FOOTER_LEN)
        footer = f.read(FOOTER_LEN)
        if struct.unpack_from('<Q', footer, 40)[0] != TABLE_MAGIC:
            raise IOError('%s: not a TensorFlow checkpoint index (bad magic)' % path)
        _, _, p = _decode_handle(footer, 0)                 # meta-index handle (unused)
        ioff, isize, _ = _decode_handle(footer, p)
        for _, handle in _block_entries(_read_block(f, ioff, isize, verify)):
            boff, bsize, _ = _decode_handle(handle)
            out.extend(_block_entries(_read_block(f, boff, bsize, verify)))
    return out


class _BlockBuilder:
    def __init__(self, interval):
        self.interval = interval
        self.reset()

    def reset(self):
        self.buf = bytearray()
        self.restarts = [0]
        self.count = 0
        self.last = b''

    def add(self, key, value):
        shared = 0
        if self.count < self.interval:
            m = min(len(key), len(self.last))
            while shared < m and key[shared] == self.last[shared]:
                shared += 1
        else:
            self.restarts.append(len(self.buf))
            self.count = 0
        _put_varint(self.buf, shared)
        _put_varint(self.buf, len(key) - shared)
        _put_varint(self.buf, len(value))
        self.buf += key[shared:]
        self.buf += value
        self.last = key
        self.count += 1

    def size(self):
        return len(self.buf) + 4 * len(self.restarts) + 4

    def empty(self):
        return not self.buf

    def finish(self):
        out = bytes(self.buf) + b''.join(struct.pack('<I', r) for r in self.restarts) + struct.pack('<I', len(self.restarts))
        return out


def write_table(path, items, block_size=BLOCK_SIZE):
    """Write sorted (key, value) byte pairs as an uncompressed table file (table_builder.cc)."""
    keys = [k for k, _ in items]
    if keys != sorted(keys) or len(set(keys)) != len(keys):
        raise ValueError('table keys must be unique and sorted')
    with open(path, 'wb') as f:
        pos = [0]

        def emit(body):
            trailer = b'\x00'
            c = mask_crc(crc32c(body + trailer))
            f.write(body + trailer + struct.pack('<I', c))
            h = bytearray()
            _put_varint(h, pos[0])
            _put_varint(h, len(body))
            pos[0] += len(body) + BLOCK_TRAILER
            return bytes(h)

        data, index = _BlockBuilder(RESTART_INTERVAL), _BlockBuilder(1)
        for k, v in items:
            data.add(k, v)
            if data.size() >= block_size:
                index.add(data.last, emit(data.finish()))   # the block's last key is a valid separator
                data.reset()
        if not data.empty():
            index.add(data.last, emit(data.finish()))
        meta_h = emit(_BlockBuilder(RESTART_INTERVAL).finish())
        index_h = emit(index.finish())
        footer = meta_h + index_h
        footer += b'\x00' * (40 - len(footer)) + struct.pack('<Q', TABLE_MAGIC)
        f.write(footer)


# -------------------------------------------------------------------------------------------------------------------- bundle
def _data_path(prefix, shard, num_shards):
    return '%s.data-%05d-of-%05d' % (prefix, shard, num_shards)


def is_bundle(prefix):
    return bool(prefix) and os.path.isfile(prefix + '.index')


def list_variables(prefix):
    """[(name, shape, numpy dtype)] like tf.train.list_variables."""
    out = []
    for k, v in read_table(prefix + '.index'):
        if k == b'':
            continue
        e = _parse_entry(v)
        out.append((k.decode(), e['shape'], 'bfloat16' if e['dtype'] == _DT_BFLOAT16 else _DT.get(e['dtype'])))
    return out


def read_bundle(prefix, names=None, verify=True):
    """Read a checkpoint -> {variable name: numpy array}.  `names`: optional iterable or predicate selecting variables."""
    items = read_table(prefix + '.index', verify)
    if not items or items[0][0] != b'':
        raise IOError('%s.index: missing bundle header' % prefix)
    num_shards, endian = 1, 0
    for f, wt, v in _parse_proto(items[0][1]):
        if f == 1:
            num_shards = v
        elif f == 2:
            endian = v
    if endian != 0:
        raise IOError('big-endian checkpoints are not supported')
    pred = names if callable(names) else ((lambda n, s=set(names): n in s) if names is not None else (lambda n: True))
    files, out = {}, {}
    try:
        for k, v in items[1:]:
            name = k.decode()
            if not pred(name):
                continue
            e = _parse_entry(v)
            if e['slices']:
                raise IOError('%s: partitioned (sliced) variables are not supported' % name)
            if e['dtype'] == _DT_BFLOAT16:
                dt = np.dtype(np.uint16)
            elif e['dtype'] in _DT:
                dt = np.dtype(_DT[e['dtype']])
            else:
                raise IOError('%s: unsupported dtype enum %d' % (name, e['dtype']))
            cnt = int(np.prod(e['shape'], dtype=np.int64)) if e['shape'] else 1
            if cnt * dt.itemsize != e['size']:
                raise IOError('%s: size %d does not match shape %s' % (name, e['size'], e['shape']))
            sh = e['shard_id']
            if sh not in files:
                files[sh] = open(_data_path(prefix, sh, num_shards), 'rb')
            files[sh].seek(e['offset'])
            raw = files[sh].read(e['size'])
            if len(raw) != e['size']:
                raise IOError('%s: data shard truncated' % name)
            arr = np.frombuffer(raw, dtype=dt).reshape(e['shape'])
            if verify and e['crc32c'] is not None and crc32c(arr) != unmask_crc(e['crc32c']):
                raise IOError('%s: tensor checksum mismatch' % name)
            if e['dtype'] == _DT_BFLOAT16:
                arr = (arr.astype(np.uint32) << 16).view(np.float32)
            out[name] = arr
    finally:
        for fh in files.values():
            fh.close()
    return out


def write_bundle(prefix, tensors):
    """Write {name: array} as a single-shard V2 checkpoint (BundleWriter: tensors stored back to back in key order)."""
    d = os.path.dirname(prefix)
    if d:
        os.makedirs(d, exist_ok=True)
    names = sorted(tensors, key=lambda s: s.encode())
    items = [(b'', b'\x08\x01\x1a\x02\x08\x01')]             # BundleHeaderProto{num_shards:1, version{producer:1}}
    off = 0
    tmp = _data_path(prefix, 0, 1) + '.tmp'
    with open(tmp, 'wb') as f:
        for n in names:
            if not n:
                raise ValueError('empty variable name')
            a = np.asarray(tensors[n])
            if a.dtype not in _NP2DT:
                raise ValueError('%s: dtype %s cannot be stored' % (n, a.dtype))
            shape = a.shape                                  # ascontiguousarray would turn a scalar into shape (1,)
            a = np.ascontiguousarray(a)
            if a.dtype.byteorder == '>':
                a = a.astype(a.dtype.newbyteorder('<'))
            f.write(a.tobytes())
            items.append((n.encode(), _encode_entry(_NP2DT[a.dtype], shape, off, a.nbytes, mask_crc(crc32c(a)))))
            off += a.nbytes
    os.replace(tmp, _data_path(prefix, 0, 1))
    write_table(prefix + '.index.tmp', items)
    os.replace(prefix + '.index.tmp', prefix + '.index')


# ---------------------------------------------------------------------------------------------------------- checkpoint state
def update_checkpoint_state(checkpoint_dir, latest, keep=None):
    """The text-proto `checkpoint` file tf.train.Saver maintains (CheckpointState: model_checkpoint_path +
    all_model_checkpoint_paths).  `keep`: max number of listed paths (Saver's max_to_keep)."""
    path = os.path.join(checkpoint_dir, 'checkpoint')
    allp = [p for p in read_checkpoint_state(checkpoint_dir)[1] if p != latest]
    allp.append(latest)
    dropped = []
    if keep and len(allp) > keep:
        dropped, allp = allp[:-keep], allp[-keep:]
    with open(path + '.tmp', 'w') as f:
        f.write('model_checkpoint_path: "%s"\n' % latest)
        for p in allp:
            f.write('all_model_checkpoint_paths: "%s"\n' % p)
    os.replace(path + '.tmp', path)
    return dropped


def read_checkpoint_state(checkpoint_dir):
    path = os.path.join(checkpoint_dir, 'checkpoint')
    latest, allp = None, []
    if os.path.isfile(path):
        for line in open(path):
            line = line.strip()
            if ':' not in line:
                continue
            k, v = line.split(':', 1)
            v = v.strip().strip('"')
            if k.strip() == 'model_checkpoint_path':
                latest = v
            elif k.strip() == 'all_model_checkpoint_paths':
                allp.append(v)
    return latest, allp


def latest_checkpoint(checkpoint_dir):
    """tf.train.latest_checkpoint: prefix named by the state file (relative paths resolve against the directory)."""
    latest, _ = read_checkpoint_state(checkpoint_dir)
    if latest is None:
        return None
    p = latest if os.path.isabs(latest) else os.path.join(checkpoint_dir, latest)
    return p if is_bundle(p) else None
