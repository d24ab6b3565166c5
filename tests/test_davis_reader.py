"""Host-side DAVIS2016 reader on a tiny synthetic dataset tree written to tmp."""
import os
import cv2
import numpy as np
import pytest
import torch

from oracle import tf_ops as T
from unsupervised_detection_b200.data import davis2016_data_utils as D


@pytest.fixture(scope='module')
def root(tmp_path_factory):
    r = str(tmp_path_factory.mktemp('davis'))
    rng = np.random.RandomState(0)
    lines = {'train': [], 'val': []}
    for part, seqs in (('train', ['bear', 'bus']), ('val', ['cows'])):
        for s in seqs:
            os.makedirs(os.path.join(r, 'JPEGImages/480p', s))
            os.makedirs(os.path.join(r, 'Annotations/480p', s))
            for i in range(6):
                img = (rng.rand(48, 80, 3) * 255).astype(np.uint8)
                img[:, :, 0] = i * 20                       # blue channel encodes the frame index (BGR order for cv2)
                cv2.imwrite(os.path.join(r, 'JPEGImages/480p', s, '%05d.jpg' % i), img, [cv2.IMWRITE_JPEG_QUALITY, 100])
                ann = np.zeros((48, 80), np.uint8)
                ann[10:30, 20 + i:50 + i] = 255
                cv2.imwrite(os.path.join(r, 'Annotations/480p', s, '%05d.png' % i), ann)
                lines[part].append('/JPEGImages/480p/%s/%05d.jpg /Annotations/480p/%s/%05d.png' % (s, i, s, i))
    os.makedirs(os.path.join(r, 'ImageSets/480p'))
    for part in ('train', 'val'):
        open(os.path.join(r, 'ImageSets/480p', part + '.txt'), 'w').write('\n'.join(lines[part]) + '\n')
    open(os.path.join(r, 'ImageSets/480p', 'trainval.txt'), 'w').write('\n'.join(lines['train'] + lines['val']) + '\n')
    return r


def test_directory_iterator(root):
    it = D.DirectoryIterator(root, 'train')
    assert it.samples == 12 and it.num_experiments == 2 and len(it.image_filenames[0]) == 6
    assert it.image_filenames[1][0].endswith('JPEGImages/480p/bus/00000.jpg')
    with pytest.raises(IOError):
        D.DirectoryIterator(root, 'nope')
    with pytest.raises(IOError):
        D.DirectoryIterator('/nonexistent', 'train')


def test_resizes_match_the_oracle():
    x = np.random.RandomState(1).rand(7, 9, 3).astype(np.float32)
    a = D.legacy_resize(x, 12, 16)
    b = T.resize_bilinear_legacy(torch.from_numpy(x)[None], 12, 16)[0].numpy()
    assert np.abs(a - b).max() < 1e-6
    m = np.random.RandomState(2).rand(7, 9, 1).astype(np.float32)
    assert np.array_equal(D.nn_resize(m, 5, 4), T.resize_nn_legacy(torch.from_numpy(m)[None], 5, 4)[0].numpy())
    assert D.central_crop_box(384, 640, 0.9) == (19, 32, 346, 576)


def test_training_pairs(root):
    rd = D.Davis2016Reader(root, max_temporal_len=2, min_temporal_len=1, num_threads=2, seed=3)
    it = rd.image_inputs(batch_size=4, partition='train', train_crop=0.9)
    # per sequence of 6 frames: forward heads 0..3, backward tails 2..5 -> 8 pairs x 2 sequences
    assert len(it.pairs) == 16
    assert sorted(p[0] for p in it.pairs if p[1] > 0) == [0, 1, 2, 3, 6, 7, 8, 9]
    assert sorted(p[0] for p in it.pairs if p[1] < 0) == [2, 3, 4, 5, 8, 9, 10, 11]
    i1, i2, seg, names = it.batch(4, pinned=False)
    assert i1.shape == (4, 384, 640, 3) and i2.shape == (4, 384, 640, 3) and i1.dtype == torch.float32
    assert float(i1.min()) >= -0.5 - 1e-6 and float(i1.max()) <= 0.5 + 1e-6
    assert len(names) == 4 and all(n.endswith('.jpg') for n in names)
    # the pair never crosses a sequence boundary: frame index encoded in the red..blue channel differs by 1 or 2 steps of 20/255
    d = (i1[..., 2].mean(dim=(1, 2)) - i2[..., 2].mean(dim=(1, 2))).abs() * 255 / 20
    assert all(0.8 < float(v) < 2.2 for v in d)
    for _ in range(6):                       # endless (repeat + reshuffle)
        it.batch(4, pinned=False)


def test_test_iterator_order_and_boundaries(root):
    rd = D.Davis2016Reader(root, num_threads=1)
    it = rd.test_inputs(batch_size=2, partition='val', t_len=1, with_fname=True, test_crop=0.9)
    assert rd.val_samples == 6
    assert it.pairs == [(0, 1.0), (1, 1.0), (2, 1.0), (3, 1.0), (4, 1.0), (5, -1.0)]   # last frame looks backward
    i1, i2, seg, names = it.batch(6, pinned=False)
    assert [os.path.basename(n) for n in names] == ['%05d.jpg' % i for i in range(6)]
    assert seg.shape == (6, 384, 640, 1) and float(seg.max()) == 1.0 and float(seg.min()) == 0.0
    blue = lambda t: t[..., 2].mean(dim=(1, 2)) * 255 + 127.5
    # JPEG chroma coding blurs the exact value; direction and rough size are what matter: +1 frame forward, last frame backward
    assert 5 < float(blue(i2)[0] - blue(i1)[0]) < 35 and -35 < float(blue(i2)[5] - blue(i1)[5]) < -5
    it2 = rd.test_inputs(partition='val', t_len=-2)
    assert it2.pairs[:2] == [(0, 1.0), (1, 1.0)] and it2.pairs[2] == (2, -1.0)


def test_video_to_davis_layout_feeds_the_reader(tmp_path):
    """scripts/create_data_frvideo.py: a short synthetic clip becomes a DAVIS-style tree the reader accepts."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('create_data_frvideo', os.path.join(os.path.dirname(os.path.dirname(
        os.path.abspath(__file__))), 'scripts', 'create_data_frvideo.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    clip = str(tmp_path / 'clip.avi')
    vw = cv2.VideoWriter(clip, cv2.VideoWriter_fourcc(*'MJPG'), 12.0, (96, 64))
    if not vw.isOpened():
        pytest.skip('no MJPG video writer in this OpenCV build')
    for i in range(10):
        f = np.full((64, 96, 3), 40, np.uint8)
        f[:, 8 * i:8 * i + 8] = 220
        vw.write(f)
    vw.release()
    out = str(tmp_path / 'ds')
    n = mod.convert(clip, out, fps=12.0, size=(80, 48))
    assert n == 10 and len(os.listdir(os.path.join(out, 'JPEGImages/480p/clip'))) == 10
    rd = D.Davis2016Reader(out, num_threads=1)
    it = rd.test_inputs(batch_size=2, partition='val', t_len=1, test_crop=1.0)
    i1, i2, seg, names = it.batch(3, pinned=False)
    assert i1.shape == (3, 384, 640, 3) and float(seg.max()) == 0.0 and names[0].endswith('clip/00000.jpg')
    with pytest.raises(IOError):
        mod.convert(str(tmp_path / 'missing.mp4'), out)


def test_learner_wires_the_davis_reader(root):
    """AdversarialLearner.load_training_data (adversarial_learner.py:22-70 / :454-470): train + validation iterators, inference iterator."""
    from unsupervised_detection_b200.common_flags import Config
    from unsupervised_detection_b200.models.adversarial_learner import AdversarialLearner
    L = object.__new__(AdversarialLearner)
    L.rank = 0
    L.config = Config(dataset='DAVIS2016', root_dir=root, batch_size=2, train_partition='train', max_temporal_len=2, min_temporal_len=1)
    L.load_training_data()
    assert L.num_samples_val == 6 and len(L.reader.pairs) == 16 and len(L.val_reader.pairs) == 6
    L._inference, L.aug_test = True, True
    L.config = Config(dataset='DAVIS2016', root_dir=root, batch_size=1, test_partition='val', test_temporal_shift=1, test_crop=0.9)
    L.load_training_data()
    assert L.reader.val_samples == 6 and L.dataset_reader.test_crop == 1.0          # aug_test crops later (0.85..1.0), reader stays uncropped
    with pytest.raises(IOError):
        L.config = Config(dataset='DAVIS2016', root_dir='/nonexistent')
        L.load_training_data()


def test_validation_iterator_keeps_its_own_file_list_after_the_training_iterator_is_built(root):
    """load_training_data builds the validation iterator first and the training iterator (another partition) second on the SAME
    reader: validation batches must still come from the validation list, frames and annotations alike."""
    from unsupervised_detection_b200.common_flags import Config
    from unsupervised_detection_b200.models.adversarial_learner import AdversarialLearner
    L = object.__new__(AdversarialLearner)
    L.rank = 0
    L.config = Config(dataset='DAVIS2016', root_dir=root, batch_size=2, train_partition='train', test_partition='val', max_temporal_len=2,
                      min_temporal_len=1)
    L.load_training_data()
    i1, i2, seg, names = L.val_reader.batch(6, pinned=False)
    assert all('/cows/' in n for n in names), names
    assert float(seg.max()) > 0.5                                           # the cows annotations, decoded with the cows frames
    ref = D.Davis2016Reader(root, num_threads=1).test_inputs(batch_size=2, partition='val', t_len=L.config.test_temporal_shift,
                                                             test_crop=L.config.test_crop)
    r1, r2, rseg, rnames = ref.batch(6, pinned=False)
    assert names == rnames and torch.equal(i1, r1) and torch.equal(i2, r2) and torch.equal(seg, rseg)
    t1, _, _, tnames = L.reader.batch(4, pinned=False)
    assert all('/bear/' in n or '/bus/' in n for n in tnames), tnames


def test_host_c_preprocessing_is_bit_identical_to_the_numpy_restatement(root):
    """cis_host_resize_bilinear_legacy / cis_host_bgr8_to_rgb_resized (libcis_b200 host routines used by every reader) against the numpy
    gather version and, through it, the oracle's legacy resize."""
    rng = np.random.RandomState(5)
    for (h, w, oh, ow, c) in [(48, 80, 384, 640, 3), (230, 390, 384, 640, 3), (7, 9, 12, 16, 3), (345, 576, 384, 640, 1), (5, 3, 11, 2, 2),
                              (1, 1, 4, 4, 3), (20, 20, 7, 33, 4)]:
        x = rng.rand(h, w, c).astype(np.float32) - 0.5
        assert np.array_equal(D.legacy_resize(x, oh, ow), D.legacy_resize_numpy(x, oh, ow))
        assert np.array_equal(D.legacy_resize(x[::-1, ::-1], oh, ow), D.legacy_resize_numpy(np.ascontiguousarray(x[::-1, ::-1]), oh, ow))
    x = rng.rand(6, 7, 3).astype(np.float32)
    assert np.abs(D.legacy_resize(x, 9, 5) - T.resize_bilinear_legacy(torch.from_numpy(x)[None], 9, 5)[0].numpy()).max() < 1e-6
    path = os.path.join(root, 'JPEGImages/480p/bear/00002.jpg')
    bgr = cv2.imread(path)
    rgb = cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB).astype(np.float32) / np.float32(255.0) - np.float32(0.5)
    assert np.array_equal(D.Davis2016Reader.preprocess_image(path), D.legacy_resize_numpy(rgb, 384, 640))


def test_prefetching_iterator_yields_the_same_stream(root):
    """Background prefetch (dataset.prefetch of the reference) changes neither order nor content, survives a batch-size change and
    propagates reader errors to the consumer."""
    def stream(prefetch):
        rd = D.Davis2016Reader(root, max_temporal_len=2, min_temporal_len=1, num_threads=3, seed=11)
        rd.prefetch = prefetch
        it = rd.image_inputs(batch_size=2, partition='train', train_crop=0.8)
        out = [it.batch(2, pinned=False) for _ in range(5)] + [it.batch(3, pinned=False) for _ in range(2)]
        it.close()
        return out
    a, b = stream(0), stream(3)
    for x, y in zip(a[:5], b[:5]):                                  # constant batch size: identical stream
        assert x[3] == y[3] and torch.equal(x[0], y[0]) and torch.equal(x[1], y[1])
    # a batch-size change restarts the producer; batches it had decoded ahead are dropped, so only the shapes are comparable
    assert a[0][0].shape == (2, 384, 640, 3) and a[-1][0].shape == (3, 384, 640, 3)
    rd = D.Davis2016Reader(root, num_threads=2, seed=1)
    rd.prefetch = 2
    it = rd.image_inputs(batch_size=2, partition='train', train_crop=0.9)
    it.batch(2, pinned=False)
    it.view.filenames = ['/nonexistent.jpg'] * len(rd.filenames)      # every later sample of THIS iterator fails to decode
    with pytest.raises(IOError):
        for _ in range(6):
            it.batch(2, pinned=False)
    it.close()


def test_multi_crop_matches_the_readers_central_cropping():
    """data/crops.py (aug_test path of AdversarialLearner.inference) cuts and resizes exactly like Davis2016Reader.central_cropping,
    i.e. with tf.image.central_crop's geometry, for images and masks."""
    from unsupervised_detection_b200.data.crops import central_crops
    rng = np.random.RandomState(9)
    img1, img2 = (rng.rand(1, 48, 80, 3).astype(np.float32) - 0.5 for _ in range(2))
    gt = (rng.rand(1, 48, 80, 1) > 0.5).astype(np.float32)
    crops = [0.85, 0.9, 0.95, 1.0]
    o1, o2, og = central_crops(torch.from_numpy(img1), torch.from_numpy(img2), torch.from_numpy(gt), crops)
    assert o1.shape == (4, 48, 80, 3) and og.shape == (4, 48, 80, 1)
    for i, c in enumerate(crops):
        assert np.abs(o1[i].numpy() - D.Davis2016Reader.central_cropping(img1[0], c)).max() < 1e-6
        assert np.abs(o2[i].numpy() - D.Davis2016Reader.central_cropping(img2[0], c)).max() < 1e-6
        assert np.abs(og[i].numpy() - D.Davis2016Reader.central_cropping(gt[0], c)).max() < 1e-6
    assert torch.equal(o1[3], torch.from_numpy(img1[0]))                     # crop 1.0 is the identity
    assert D.central_crop_box(384, 640, 0.85) == (28, 48, 328, 544)          # int((384 - 326.4) / 2) = 28, not (384 - 326) // 2 = 29


def test_validation_iterator_shards_across_ranks(root):
    """DP validation: with world_size 2 and a global batch of 4 the two ranks read complementary halves of every global batch."""
    names = {}
    for rank in (0, 1):
        rd = D.Davis2016Reader(root, num_threads=1, seed=8964 + rank)
        it = rd.test_inputs(batch_size=4, partition='val', t_len=1, test_crop=1.0).shard(rank, 2, 4)
        names[rank] = [os.path.basename(n) for _ in range(2) for n in it.batch(2, pinned=False)[3]]
        it.close()
    rd = D.Davis2016Reader(root, num_threads=1)
    full = rd.test_inputs(batch_size=4, partition='val', t_len=1, test_crop=1.0)
    glob = [os.path.basename(n) for _ in range(2) for n in full.batch(4, pinned=False)[3]]
    full.close()
    # val has 6 frames: global batches [0..3] and [4,5,0,1] (wrap like dataset.repeat)
    assert names[0] == [glob[0], glob[1], glob[4], glob[5]] and names[1] == [glob[2], glob[3], glob[6], glob[7]]
    assert full.shard(0, 1, 4) is full
