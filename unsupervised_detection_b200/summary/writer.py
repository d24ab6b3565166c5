"""TensorBoard `events.out.tfevents.*` writer/reader without TensorFlow  (SURVEY 8f-4).

The reference logs through `tf.summary.scalar/image/histogram` merged into `step_sum` / `validation_summary`
(models/adversarial_learner.py:260-298) and `sv.summary_writer.add_summary(summary, step)` (:403, :439).  TensorFlow is a
third-party dependency that is not available here; this restates its published formats:

  * an event file is a TFRecord stream (tensorflow/core/lib/io/record_writer.cc): per record
    `u64 length | u32 masked_crc32c(length) | bytes | u32 masked_crc32c(bytes)`;
  * each record is an `Event` proto (tensorflow/core/util/event.proto): wall_time=1 (double), step=2 (int64),
    file_version=3 (string, first record "brain.Event:2"), summary=5;
  * `Summary.Value` (tensorflow/core/framework/summary.proto): tag=1, simple_value=2 (float), image=4 {height=1,width=2,
    colorspace=3, encoded_image_string=4 (PNG)}, histo=5 {min=1,max=2,num=3,sum=4,sum_squares=5, bucket_limit=6 packed,
    bucket=7 packed};
  * float images are normalised like summary_image_op.cc NormalizeFloatImage and tagged `<name>/image` for max_outputs=1;
  * histograms use tensorflow/core/lib/histogram/histogram.cc's default bucket limits (+-1e-12 * 1.1^k up to 1e20, DBL_MAX ends)
    with runs of empty buckets collapsed.

Pinned against TensorBoard's own implementation of the format: tests/test_summary_tensorboard.py reads these files back with
`tensorboard.backend.event_processing` (TensorBoard 2.x is in the image; its TensorFlow stub re-implements the TFRecord framing and
masked CRC-32C) and parses the payloads with TensorBoard's proto classes.  TensorFlow's own bucket table / image normalisation is not
available offline; those two rules are restated from histogram.cc / summary_image_op.cc and checked structurally.
"""
import os
import socket
import struct
import sys
import time

import numpy as np

from ..checkpoint.tf_bundle import crc32c, mask_crc, unmask_crc, _parse_proto, _put_varint


# ------------------------------------------------------------------------------------------------------------------ protobuf
def _key(field, wt):
    return bytes([(field << 3) | wt]) if field < 16 else bytes([((field << 3) | wt) & 0x7f | 0x80, field >> 4])


def _pb_bytes(field, b):
    out = bytearray(_key(field, 2))
    _put_varint(out, len(b))
    return bytes(out) + bytes(b)


def _pb_varint(field, v):
    out = bytearray(_key(field, 0))
    _put_varint(out, v)
    return bytes(out)


def _pb_double(field, v):
    return _key(field, 1) + struct.pack('<d', v)


def _pb_float(field, v):
    return _key(field, 5) + struct.pack('<f', v)


# ----------------------------------------------------------------------------------------------------------------- histogram
def _default_limits():
    pos, v = [], 1.0e-12
    while v < 1.0e20:
        pos.append(v)
        v *= 1.1
    pos.append(sys.float_info.max)
    return np.array([-x for x in reversed(pos)] + [0.0] + pos, dtype=np.float64)


_LIMITS = _default_limits()


def histogram_proto(values):
    """HistogramProto bytes of a tensor with TF's default buckets (Histogram::Add: bucket = upper_bound(limits, v);
    EncodeToProto(preserve_zero_buckets=false))."""
    v = np.asarray(values, dtype=np.float64).reshape(-1)
    if v.size == 0:
        mn, mx = sys.float_info.max, -sys.float_info.max        # Histogram::Clear()
    else:
        mn, mx = float(v.min()), float(v.max())
    idx = np.minimum(np.searchsorted(_LIMITS, v, side='right'), len(_LIMITS) - 1)
    counts = np.bincount(idx, minlength=len(_LIMITS)).astype(np.float64)
    lim, cnt, i, n = [], [], 0, len(_LIMITS)
    while i < n:
        end, c = _LIMITS[i], counts[i]
        i += 1
        if c <= 0.0:
            while i < n and counts[i] <= 0.0:
                end, c = _LIMITS[i], counts[i]
                i += 1
        lim.append(end)
        cnt.append(c)
    out = _pb_double(1, mn) + _pb_double(2, mx) + _pb_double(3, float(v.size)) + _pb_double(4, float(v.sum())) \
        + _pb_double(5, float(np.dot(v, v)))
    out += _pb_bytes(6, struct.pack('<%dd' % len(lim), *lim)) + _pb_bytes(7, struct.pack('<%dd' % len(cnt), *cnt))
    return out


# --------------------------------------------------------------------------------------------------------------------- image
def normalize_float_image(img):
    """summary_image_op.cc NormalizeFloatImage for ONE image [H,W,C] float -> uint8 (non-finite pixels -> the op's bad_color
    red for 3 channels)."""
    a = np.asarray(img, dtype=np.float32)
    fin = np.isfinite(a)
    if fin.any():
        mn, mx = float(a[fin].min()), float(a[fin].max())
    else:
        mn, mx = 0.0, 0.0
    if mn < 0:
        m = max(abs(mn), abs(mx))
        scale, offset = (0.0 if m < 1e-6 else 127.0 / m), 128.0
    else:
        scale, offset = (0.0 if mx < 1e-6 else 255.0 / mx), 0.0
    out = (np.where(fin, a, 0.0) * np.float32(scale) + np.float32(offset)).astype(np.uint8)
    bad = ~fin.all(axis=-1)
    if bad.any():
        out[bad] = ([255, 0, 0] + [255] * (a.shape[-1] - 3))[:a.shape[-1]] if a.shape[-1] >= 3 else 255
    return out


def _png(u8):
    import cv2
    a = u8 if u8.shape[-1] != 3 else u8[..., ::-1]               # cv2 expects BGR
    ok, buf = cv2.imencode('.png', np.ascontiguousarray(a))
    if not ok:
        raise RuntimeError('PNG encoding failed')
    return buf.tobytes()


# -------------------------------------------------------------------------------------------------------------------- writer
class SummaryWriter(object):
    """Minimal tf.summary.FileWriter: `add_scalar/add_image/add_histogram` buffer Summary.Value protos, `flush_step(step)` emits
    them as ONE Event (like one merged `step_sum` evaluation, adversarial_learner.py:291,403)."""

    def __init__(self, logdir):
        os.makedirs(logdir, exist_ok=True)
        self.path = os.path.join(logdir, 'events.out.tfevents.%010d.%s' % (int(time.time()), socket.gethostname()))
        self._f = open(self.path, 'ab')
        self._values = []
        self._record(_pb_double(1, time.time()) + _pb_bytes(3, b'brain.Event:2'))

    def _record(self, data):
        hdr = struct.pack('<Q', len(data))
        self._f.write(hdr + struct.pack('<I', mask_crc(crc32c(hdr))) + data + struct.pack('<I', mask_crc(crc32c(data))))
        self._f.flush()

    def add_scalar(self, tag, value):
        self._values.append(_pb_bytes(1, tag.encode()) + _pb_float(2, float(value)))

    def add_image(self, name, image):
        """tf.summary.image(name, batch, max_outputs=1): first image of the batch, tag `<name>/image`."""
        a = np.asarray(image)
        if a.ndim == 4:
            a = a[0]
        u8 = a if a.dtype == np.uint8 else normalize_float_image(a)
        img = _pb_varint(1, u8.shape[0]) + _pb_varint(2, u8.shape[1]) + _pb_varint(3, u8.shape[2]) + _pb_bytes(4, _png(u8))
        self._values.append(_pb_bytes(1, (name + '/image').encode()) + _pb_bytes(4, img))

    def add_histogram(self, tag, values):
        self._values.append(_pb_bytes(1, tag.encode()) + _pb_bytes(5, histogram_proto(values)))

    def flush_step(self, step):
        summ = b''.join(_pb_bytes(1, v) for v in self._values)
        self._values = []
        self._record(_pb_double(1, time.time()) + _pb_varint(2, int(step)) + _pb_bytes(5, summ))

    def close(self):
        self._f.close()


# -------------------------------------------------------------------------------------------------------------------- reader
def read_events(path, verify=True):
    """-> list of {'wall_time', 'step', 'file_version'?, 'values': [{tag, simple_value | image{h,w,c,png} | histo{...}}]}."""
    out = []
    with open(path, 'rb') as f:
        raw = f.read()
    pos = 0
    while pos < len(raw):
        ln, = struct.unpack_from('<Q', raw, pos)
        lcrc, = struct.unpack_from('<I', raw, pos + 8)
        data = raw[pos + 12:pos + 12 + ln]
        dcrc, = struct.unpack_from('<I', raw, pos + 12 + ln)
        if verify and (unmask_crc(lcrc) != crc32c(raw[pos:pos + 8]) or unmask_crc(dcrc) != crc32c(data)):
            raise IOError('corrupt record at byte %d' % pos)
        pos += 16 + ln
        ev = {'values': []}
        for fld, wt, v in _parse_proto(data):
            if fld == 1:
                ev['wall_time'] = struct.unpack('<d', struct.pack('<Q', v))[0]
            elif fld == 2:
                ev['step'] = v
            elif fld == 3:
                ev['file_version'] = v.decode()
            elif fld == 5:
                for f2, _, val in _parse_proto(v):
                    if f2 != 1:
                        continue
                    d = {}
                    for f3, w3, x in _parse_proto(val):
                        if f3 == 1:
                            d['tag'] = x.decode()
                        elif f3 == 2:
                            d['simple_value'] = struct.unpack('<f', struct.pack('<I', x))[0]
                        elif f3 == 4:
                            im = {}
                            for f4, _, y in _parse_proto(x):
                                im[{1: 'height', 2: 'width', 3: 'colorspace', 4: 'png'}.get(f4, f4)] = y
                            d['image'] = im
                        elif f3 == 5:
                            h = {}
                            for f4, w4, y in _parse_proto(x):
                                if f4 <= 5:
                                    h[{1: 'min', 2: 'max', 3: 'num', 4: 'sum', 5: 'sum_squares'}[f4]] = \
                                        struct.unpack('<d', struct.pack('<Q', y))[0]
                                else:
                                    h['bucket_limit' if f4 == 6 else 'bucket'] = list(struct.unpack('<%dd' % (len(y) // 8), y))
                            d['histo'] = h
                    ev['values'].append(d)
        out.append(ev)
    return out
