"""The event files written by unsupervised_detection_b200.summary are read back with TensorBoard's OWN reader and proto classes
(`tensorboard` 2.x is installed in this image; its TensorFlow stub implements the TFRecord framing and masked CRC-32C independently
of this package).  This pins the event-file format -- record framing, Event / Summary protos, image and histogram payloads --
against a third-party implementation of the format the reference writes through tf.summary (adversarial_learner.py:260-298,403)."""
import struct

import numpy as np
import pytest

tb_loader = pytest.importorskip('tensorboard.backend.event_processing.event_file_loader')
from tensorboard.compat.proto import summary_pb2                                   # noqa: E402
from tensorboard.compat.tensorflow_stub import pywrap_tensorflow as tb_stub        # noqa: E402

from unsupervised_detection_b200.summary import SummaryWriter, histogram_proto, normalize_float_image   # noqa: E402
from unsupervised_detection_b200.checkpoint.tf_bundle import crc32c, mask_crc      # noqa: E402


def test_masked_crc_matches_tensorboards_implementation():
    rng = np.random.RandomState(0)
    for n in (0, 1, 8, 9, 63, 1000, 4097):
        data = rng.randint(0, 256, n).astype(np.uint8).tobytes()
        assert crc32c(data) == tb_stub.crc32c(data)
        assert mask_crc(crc32c(data)) == tb_stub.masked_crc32c(data)


def test_tensorboard_reads_our_event_file(tmp_path):
    import cv2
    w = SummaryWriter(str(tmp_path))
    rng = np.random.RandomState(1)
    img = rng.rand(1, 10, 14, 3).astype(np.float32) - 0.5
    grads = rng.randn(5000) * 0.05
    w.add_scalar('recover', 0.125)
    w.add_scalar('generator', -2.5)
    w.add_image('PWC_Flow', img)
    w.add_histogram('FlownetS//aconv1/weights/gradients', grads)
    w.flush_step(42)
    w.add_scalar('IoU on Validation', 0.375)
    w.flush_step(3)
    w.close()
    events = list(tb_loader.LegacyEventFileLoader(w.path).Load())
    assert events[0].file_version == 'brain.Event:2' and events[0].wall_time > 1e9
    e = events[1]
    assert e.step == 42 and [v.tag for v in e.summary.value] == ['recover', 'generator', 'PWC_Flow/image', 'FlownetS//aconv1/weights/gradients']
    assert e.summary.value[0].simple_value == 0.125 and e.summary.value[1].simple_value == -2.5
    im = e.summary.value[2].image
    assert (im.height, im.width, im.colorspace) == (10, 14, 3)
    dec = cv2.imdecode(np.frombuffer(im.encoded_image_string, np.uint8), cv2.IMREAD_COLOR)[..., ::-1]
    assert np.array_equal(dec, normalize_float_image(img[0]))
    h = e.summary.value[3].histo
    assert h.num == 5000 and abs(h.sum - grads.sum()) < 1e-9 and abs(h.sum_squares - (grads ** 2).sum()) < 1e-9
    assert h.min == grads.min() and h.max == grads.max()
    assert len(h.bucket) == len(h.bucket_limit) and sum(h.bucket) == 5000
    # every value lies in the bucket whose limit is the first one above it
    lim = np.array(h.bucket_limit)
    cnt = np.bincount(np.searchsorted(lim, grads, side='right'), minlength=len(lim))
    assert np.array_equal(cnt, np.array(h.bucket, dtype=np.int64))
    assert events[2].step == 3 and events[2].summary.value[0].tag == 'IoU on Validation'
    # the modern loader (data-compat layer: scalars/images/histograms become tensors) accepts the file as well
    assert sum(1 for _ in tb_loader.EventFileLoader(w.path).Load()) == 3


def test_histogram_proto_parses_with_tensorboards_proto_class():
    vals = np.array([-0.2, -1e-13, 0.0, 1e-13, 0.1999, 3.5])
    h = summary_pb2.HistogramProto()
    h.ParseFromString(histogram_proto(vals))
    assert (h.min, h.max, h.num) == (-0.2, 3.5, 6.0)
    assert list(h.bucket_limit) == sorted(h.bucket_limit) and sum(h.bucket) == 6
    # TF's default bucket limits: +-1e-12 * 1.1^k; zero and |v| < 1e-12 share the bucket bounded by 1e-12 resp. 0 ... check the edges
    lim = list(h.bucket_limit)
    i0 = next(i for i, l in enumerate(lim) if l > 0.0)
    assert abs(lim[i0] - 1e-12) < 1e-24                                # first positive limit
    assert 0.0 in lim                                                  # bucket (-1e-12, 0] exists because -1e-13 falls in it
    assert struct.pack('<d', lim[-1]) == struct.pack('<d', 1.7976931348623157e308)
