"""Host-side FBMS59 and SegTrackV2 readers on tiny dataset trees written to tmp."""
import os

import cv2
import numpy as np
import pytest

from unsupervised_detection_b200.data import fbms_data_utils as F
from unsupervised_detection_b200.data import segtrackv2_data_utils as S


def _frame(i, h=40, w=64):
    img = np.full((h, w, 3), 90, np.uint8)
    img[:, :, 0] = i * 10                                   # blue (BGR) encodes the frame index
    return img


@pytest.fixture(scope='module')
def fbms_root(tmp_path_factory):
    r = str(tmp_path_factory.mktemp('FBMS'))
    for part, cats in (('Trainingset', ['cars1', 'marple2']), ('Testset', ['cats01', 'marple7'])):
        for c in cats:
            d = os.path.join(r, part, c)
            os.makedirs(os.path.join(d, 'GroundTruth'))
            n = 8
            with open(os.path.join(d, c + '.bmf'), 'w') as f:
                f.write('%d 1\n' % n + ''.join('%s_%02d.ppm\n' % (c, i + 1) for i in range(n)))
            for i in range(n):
                cv2.imwrite(os.path.join(d, '%s_%02d.jpg' % (c, i + 1)), _frame(i), [cv2.IMWRITE_JPEG_QUALITY, 100])
            if c == 'cats01':                                # "weird" flavour: colour .ppm labels, white = unlabelled -> 0
                for k in (1, 5, 8):
                    g = np.zeros((40, 64, 3), np.uint8)
                    g[5:20, 5:30] = 128
                    g[30:, :] = 255
                    cv2.imwrite(os.path.join(d, 'GroundTruth', '%s_%02d_gt.ppm' % (c, k)), g)
                cv2.imwrite(os.path.join(d, 'GroundTruth', 'cats01_01_PROB.ppm'), np.zeros((40, 64, 3), np.uint8))
            else:                                            # grey .pgm labels numbered by frame
                for k in (1, 4, 8):
                    g = np.zeros((40, 64), np.uint8)
                    g[10:25, 8:40] = 20                      # 20/255 = 0.078: object for marple7 (thr 0.05), background at 0.1
                    g[0:5, 0:5] = 200
                    cv2.imwrite(os.path.join(d, 'GroundTruth', '%s_%03d.pgm' % (c, k)), g)
    return r


def test_fbms_directory_and_offsets(fbms_root):
    it = F.DirectoryIterator(fbms_root, 'train')
    assert it.samples == 16 and it.num_experiments == 2
    assert it.image_filenames[0][0].endswith('Trainingset/cars1/cars1_01.jpg')           # .ppm in the .bmf -> .jpg
    assert F.DirectoryIterator(fbms_root, 'trainval').samples == 32
    with pytest.raises(IOError):
        F.DirectoryIterator(os.path.join(fbms_root, 'nope'), 'train')
    # annotated frames 1,4,8 -> rebased 0,3,7; shift +2 mirrored at the tail; shift -2 mirrored at the head
    n, o = F.test_offsets([1, 4, 8], 2)
    assert n.tolist() == [0, 3, 7] and o.tolist() == [2, 5, 5]
    n, o = F.test_offsets([1, 4, 8], -2)
    assert o.tolist() == [2, 1, 5]
    n, o = F.test_offsets([3, 4], 9)                          # clamped into the sequence
    assert o.tolist() == [1, 0] or (o >= 0).all() and (o <= 1).all()
    files, nums, weird = F.find_gt(os.path.join(fbms_root, 'Testset', 'cats01', 'GroundTruth'))
    assert weird and nums == [1, 5, 8] and all('PROB' not in f for f in files)
    files, nums, weird = F.find_gt(os.path.join(fbms_root, 'Testset', 'marple7', 'GroundTruth'))
    assert not weird and nums == [1, 4, 8]


def test_fbms_test_batches(fbms_root):
    rd = F.FBMS59Reader(fbms_root)
    it = rd.test_inputs(batch_size=3, partition='val', t_len=2, test_crop=1.0)
    assert rd.val_samples == 6 and rd.num_categories == 2 and rd.samples_per_cat == {'cats01': 3, 'marple7': 3}
    i1, i2, seg, names = it.batch(6, pinned=False)
    assert i1.shape == (6, 384, 640, 3) and seg.shape == (6, 384, 640, 1)
    assert [os.path.basename(n) for n in names] == ['cats01_01.jpg', 'cats01_05.jpg', 'cats01_08.jpg', 'marple7_01.jpg', 'marple7_04.jpg',
                                                    'marple7_08.jpg']
    assert rd.batch_samples_per_cat(names).tolist() == [3.0] * 6

    def idx(t):                                               # frame index from the blue channel (img*255+127.5)/10
        return int(round(float((t[..., 2].mean() + 0.5) * 255) / 10))
    assert [idx(i1[k]) for k in range(6)] == [0, 4, 7, 0, 3, 7]
    assert [idx(i2[k]) for k in range(6)] == [2, 6, 5, 2, 5, 5]
    # weird flavour: grey 128 is object, pure white is "unlabelled" -> background
    s = seg[0, :, :, 0].numpy()
    assert set(np.unique(s)) == {0.0, 1.0} and s[100, 100] == 1.0 and s[380, 10] == 0.0
    # marple7 threshold 0.05 keeps the faint object that the default 0.1 would drop
    s = seg[3, :, :, 0].numpy()
    assert s[150, 200] == 1.0 and s[10, 10] == 1.0 and s[300, 500] == 0.0
    m = F.binarise_gt(os.path.join(fbms_root, 'Trainingset', 'cars1', 'GroundTruth', 'cars1_001.pgm'), 'cars1', False)
    assert m[15, 20] == 0 and m[2, 2] == 255
    assert m.dtype == np.uint8 and not os.path.exists(os.path.join(fbms_root, 'Trainingset', 'cars1', 'GroundTruth', 'cars1_001.jpg'))


def test_fbms_training_pairs(fbms_root):
    rd = F.FBMS59Reader(fbms_root, max_temporal_len=3, min_temporal_len=2, num_threads=2, seed=5)
    it = rd.image_inputs(batch_size=2, partition='train', train_crop=0.9)
    assert len(it.pairs) == 2 * (5 + 5)
    a, b, ones, names = it.batch(4, pinned=False)
    assert a.shape == (4, 384, 640, 3) and float(ones.min()) == 1.0
    assert float(a.min()) >= -0.5 - 1e-5 and float(a.max()) <= 0.5 + 1e-5
    with pytest.raises(AssertionError):
        F.FBMS59Reader(fbms_root, max_temporal_len=2, min_temporal_len=2)


@pytest.fixture(scope='module')
def seg_root(tmp_path_factory):
    r = str(tmp_path_factory.mktemp('SegTrackv2'))
    os.makedirs(os.path.join(r, 'ImageSets'))
    seqs = {'bird': 5, 'frog': 6}
    open(os.path.join(r, 'ImageSets', 'all.txt'), 'w').write(''.join('*%s\n' % s for s in seqs))
    for s, n in seqs.items():
        os.makedirs(os.path.join(r, 'JPEGImages', s))
        os.makedirs(os.path.join(r, 'GroundTruth', s))
        with open(os.path.join(r, 'ImageSets', s + '.txt'), 'w') as f:
            f.write('%s 1 %d\n' % (s, n) + ''.join('%s_%05d\n' % (s, i) for i in range(n)))
        for i in range(n):
            cv2.imwrite(os.path.join(r, 'JPEGImages', s, '%s_%05d.png' % (s, i)), _frame(i))
            g = np.zeros((40, 64), np.uint8)
            g[5:15, 10 + i:30 + i] = 255
            cv2.imwrite(os.path.join(r, 'GroundTruth', s, '%s_%05d.png' % (s, i)), g)
    return r


def test_segtrack_reader(seg_root):
    it = S.DirectoryIterator(seg_root)
    assert it.components == ['bird', 'frog'] and it.samples == 11 and it.num_experiments == 2
    with pytest.raises(IOError):
        S.DirectoryIterator('/nonexistent')
    rd = S.SegTrackV2Reader(seg_root, max_temporal_len=2, min_temporal_len=1, num_threads=2)
    tr = rd.image_inputs(batch_size=2, train_crop=1.0)
    assert len(tr.pairs) == (3 + 3) + (4 + 4)
    a, b, _, _ = tr.batch(2, pinned=False)
    assert a.shape == (2, 384, 640, 3)
    te = rd.test_inputs(batch_size=4, t_len=1, with_fname=True, test_crop=0.9)
    assert rd.val_samples == 11 and len(te.pairs) == 11
    i1, i2, seg, names = te.batch(11, pinned=False)

    def idx(t):
        return int(round(float((t[..., 2].mean() + 0.5) * 255) / 10))
    first = [idx(i1[k]) for k in range(11)]
    second = [idx(i2[k]) for k in range(11)]
    # forward pairs for all but the last frame of each sequence, then the tails looking backwards
    assert sorted(first[:9]) == [0, 0, 1, 1, 2, 2, 3, 3, 4] and all(s == f + 1 for f, s in zip(first[:9], second[:9]))
    assert first[9:] == [4, 5] and second[9:] == [3, 4]
    assert seg.shape == (11, 384, 640, 1) and 0.0 < float(seg.mean()) < 0.5
    assert names[0].endswith('JPEGImages/bird/bird_00000.png')


def test_learner_selects_readers(fbms_root, seg_root):
    from unsupervised_detection_b200.common_flags import Config
    from unsupervised_detection_b200.models.adversarial_learner import AdversarialLearner
    L = object.__new__(AdversarialLearner)
    L.rank = 0
    L.config = Config(dataset='FBMS', root_dir=fbms_root, batch_size=2, min_temporal_len=2, max_temporal_len=3)
    L.load_training_data()
    assert L.num_samples_val == 6 and L.num_categories == 2 and len(L.reader.pairs) == 40   # train_partition defaults to trainval
    L.config = Config(dataset='SEGTRACK', root_dir=seg_root, batch_size=2, min_temporal_len=1, max_temporal_len=2)
    L.load_training_data()
    assert L.num_samples_val == 11
    L._inference, L.aug_test = True, False
    L.config = Config(dataset='SEGTRACK', root_dir=seg_root, batch_size=2, test_temporal_shift=1)
    L.load_training_data()
    assert L.reader.val_samples == 11
    L.config = Config(dataset='KITTI', root_dir=seg_root)
    with pytest.raises(IOError):
        L.load_training_data()
