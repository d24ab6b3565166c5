"""TensorBoard event files + flow colour code, CPU only."""
import os
import struct
import sys

import numpy as np
import pytest

from unsupervised_detection_b200.summary import SummaryWriter, read_events, normalize_float_image, histogram_proto
from unsupervised_detection_b200.summary.writer import _LIMITS
from unsupervised_detection_b200.checkpoint.tf_bundle import crc32c, mask_crc, _parse_proto
from unsupervised_detection_b200.models.utils import flow_utils as fu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'flowviz.npz')


def test_flow_colour_code_matches_reference_golden():
    """Vectors produced by the reference's own numpy code (tests/golden/make_golden_flowviz.py): bit-exact."""
    g = np.load(GOLD)
    assert np.array_equal(fu.color_wheel().astype(np.uint8), g['wheel']) and fu.color_wheel().shape == (55, 3)
    out = fu.flow_to_image(g['flow'].copy())
    assert out.dtype == np.float32 and np.array_equal(out.astype(np.uint8), g['image'])
    pm = fu.flow_to_image_pm(g['flow'].copy())
    assert pm.min() >= -0.5 and pm.max() <= 0.5
    # zero flow is white, the batch-running maximum radius makes element order matter (reference quirk kept)
    z = fu.flow_to_image(np.zeros((1, 4, 4, 2), np.float32))
    assert (z == 255).all()
    a = fu.flow_to_image(g['flow'][[1, 2]].copy())[1]
    b = fu.flow_to_image(g['flow'][[2]].copy())[0]
    assert not np.array_equal(a, b)


def test_tfrecord_framing_and_event_fields(tmp_path):
    w = SummaryWriter(str(tmp_path))
    assert os.path.basename(w.path).startswith('events.out.tfevents.')
    w.add_scalar('recover', 0.25)
    w.add_scalar('generator', -1.5)
    w.flush_step(7)
    w.add_scalar('IoU on Validation', 0.5)
    w.flush_step(1)
    w.close()
    raw = open(w.path, 'rb').read()
    ln, = struct.unpack_from('<Q', raw, 0)
    assert struct.unpack_from('<I', raw, 8)[0] == mask_crc(crc32c(raw[:8]))
    assert struct.unpack_from('<I', raw, 12 + ln)[0] == mask_crc(crc32c(raw[12:12 + ln]))
    first = _parse_proto(raw[12:12 + ln])
    assert [f for f, _, _ in first] == [1, 3] and first[1][2] == b'brain.Event:2'
    ev = read_events(w.path)
    assert ev[0]['file_version'] == 'brain.Event:2' and len(ev) == 3
    assert ev[1]['step'] == 7 and [(v['tag'], v['simple_value']) for v in ev[1]['values']] == [('recover', 0.25), ('generator', -1.5)]
    assert ev[2]['step'] == 1 and ev[2]['values'][0]['tag'] == 'IoU on Validation'
    bad = bytearray(raw)
    bad[-6] ^= 1
    open(w.path, 'wb').write(bytes(bad))
    with pytest.raises(IOError):
        read_events(w.path)


def test_float_image_normalisation_rule():
    # any negative value: scale 127/max|x| around 128; otherwise 255/max
    a = np.array([[[-0.5, 0.0, 0.25]]], np.float32)
    assert normalize_float_image(a).tolist() == [[[int(128 - 127), 128, int(128 + 63.5)]]]
    b = np.array([[[0.0, 1.0, 2.0]]], np.float32)
    assert normalize_float_image(b).tolist() == [[[0, 127, 255]]]
    assert normalize_float_image(np.zeros((2, 2, 3), np.float32)).max() == 0
    c = np.array([[[np.nan, 0.0, 1.0], [0.5, 0.5, 1.0]]], np.float32)
    assert normalize_float_image(c)[0, 0].tolist() == [255, 0, 0] and normalize_float_image(c)[0, 1].tolist() == [127, 127, 255]


def test_image_summary_roundtrip(tmp_path):
    import cv2
    w = SummaryWriter(str(tmp_path))
    rng = np.random.RandomState(0)
    batch = rng.rand(2, 12, 20, 3).astype(np.float32) - 0.5
    w.add_image('input_image', batch)
    w.flush_step(3)
    w.close()
    v = read_events(w.path)[1]['values'][0]
    assert v['tag'] == 'input_image/image' and (v['image']['height'], v['image']['width'], v['image']['colorspace']) == (12, 20, 3)
    dec = cv2.imdecode(np.frombuffer(v['image']['png'], np.uint8), cv2.IMREAD_COLOR)[..., ::-1]
    assert np.array_equal(dec, normalize_float_image(batch[0]))


def test_histogram_buckets_follow_tf_rule():
    assert len(_LIMITS) % 2 == 1
    assert _LIMITS[len(_LIMITS) // 2] == 0.0 and _LIMITS[-1] == sys.float_info.max and _LIMITS[0] == -sys.float_info.max
    assert abs(_LIMITS[len(_LIMITS) // 2 + 1] - 1e-12) < 1e-24 and abs(_LIMITS[len(_LIMITS) // 2 + 2] / 1.1e-12 - 1) < 1e-12
    vals = np.array([0.0, 0.0, 0.15, 0.15, 0.15, -0.2, 3.0], np.float64)
    h = {}
    for f, wt, v in _parse_proto(histogram_proto(vals)):
        h[f] = struct.unpack('<d', struct.pack('<Q', v))[0] if f <= 5 else np.frombuffer(v, np.float64)
    assert (h[1], h[2], h[3]) == (-0.2, 3.0, 7.0) and abs(h[4] - vals.sum()) < 1e-12 and abs(h[5] - (vals ** 2).sum()) < 1e-12
    lim, cnt = h[6], h[7]
    assert len(lim) == len(cnt) and cnt.sum() == 7 and (np.diff(lim) > 0).all() and lim[-1] == sys.float_info.max
    # every value sits in the first bucket whose limit is > value (upper_bound); zeros land in the bucket bounded by 1e-12
    nz = [(l, c) for l, c in zip(lim, cnt) if c > 0]
    assert [c for _, c in nz] == [1, 2, 3, 1]
    assert nz[1][0] == pytest.approx(1e-12) and nz[0][0] > -0.2 and nz[0][0] <= -0.2 / 1.1 + 1e-9
    assert nz[2][0] > 0.15 and nz[2][0] <= 0.15 * 1.1 and nz[3][0] > 3.0
    # runs of empty buckets are collapsed to one entry: no two consecutive zero counts
    assert not any(cnt[i] == 0 and cnt[i + 1] == 0 for i in range(len(cnt) - 1))
    e = {f: v for f, wt, v in _parse_proto(histogram_proto(np.zeros(0)))}
    assert np.frombuffer(e[7], np.float64).tolist() == [0.0]


class _Store:
    def __init__(self, names, seed):
        import torch
        g = torch.Generator().manual_seed(seed)
        self.entries, off = [], 0
        for n in names:
            self.entries.append((n, (10,), 10, off, None))
            off += 16                                   # padded slots, like engine.ParamStore
        self.grad = torch.randn(off, generator=g) * 0.3


class _Graph:
    def __init__(self):
        import torch
        g = torch.Generator().manual_seed(0)
        self.B = 2
        self.flow = torch.randn(2, 8, 12, 2, generator=g)
        self.mask = torch.rand(2, 8, 12, 1, generator=g)
        self.pred = torch.randn(6, 8, 12, 2, generator=g)
        self.image = torch.rand(2, 8, 12, 3, generator=g) - 0.5
        self.img2 = torch.rand(2, 16, 24, 3, generator=g) - 0.5
        self.rec_store = _Store(['FlownetS/aconv1/weights', 'FlownetS/aconv1/biases'], 1)
        self.gen_store = _Store(['MaskNet/conv1/kernel', 'MaskNet/conv1/gamma', 'MaskNet/conv13_upsample/beta'], 2)

    def losses(self, full=False, reduce=None):
        assert full
        return dict(generator=1.0, recover=2.0, red_rate=0.5, red_rate_compl=0.5, reconstruction_loss=3.0,
                    reconstruction_compl_loss=4.0, denominator_red_rate=80.0, denominator_red_rate_compl=81.0)


def test_learner_step_summary_layout(tmp_path):
    from unsupervised_detection_b200.common_flags import Config
    from unsupervised_detection_b200.models.adversarial_learner import AdversarialLearner
    L = object.__new__(AdversarialLearner)
    L.rank, L.graph = 0, _Graph()
    L.config = Config(checkpoint_dir=str(tmp_path))
    w = L.collect_summaries()
    other = L._net_gradients('G')
    L._write_step_summary(12, 'R', other)
    w.close()
    ev = read_events(w.path)[1]
    assert ev['step'] == 12
    tags = [v['tag'] for v in ev['values']]
    assert tags[:8] == ['generator', 'recover', 'red_rate', 'red_rate_compl', 'reconstruction_loss', 'reconstruction_compl_loss',
                        'denominator_red_rate', 'denominator_red_rate_compl']
    assert tags[8:14] == ['input_image/image', 'next_image/image', 'masked_flow/image', 'PWC_Flow/image', 'Rec_flow/image',
                          'Rec_flow_compl/image']
    assert tags[14:] == ['FlownetS//aconv1/weights/gradients', 'FlownetS//aconv1/biases/gradients', 'MaskNet//conv1/kernel/gradients',
                         'MaskNet//batch_normalization/gamma/gradients', 'MaskNet//conv13_upsample/batch_normalization/beta/gradients']
    h = ev['values'][14]['histo']
    assert h['num'] == 10 and h['min'] >= -0.2001 and h['max'] <= 0.2001           # clipped like train_op's return value, pads excluded
    assert ev['values'][9]['image']['height'] == 16 and ev['values'][8]['image']['height'] == 8
    # rank != 0 or no checkpoint_dir: no writer
    L.rank = 1
    assert L.collect_summaries() is None
