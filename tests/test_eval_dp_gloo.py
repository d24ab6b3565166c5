"""Batch-sharded evaluation (BASELINE config 4) on CPU: two gloo ranks run the DP code path of test_generator.py and
test_generator_ensemble.py with a stub learner (no GPU); the merged scores and the written files must equal the single-process run."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from unsupervised_detection_b200 import eval_dp   # noqa: E402

NAMES = ['ds/catA/%05d.jpg' % i for i in range(5)] + ['ds/catB/%05d.jpg' % i for i in range(4)] + ['ds/catA/%05d.jpg' % i for i in range(5, 7)]
H, W = 16, 24


def _frame(gidx, crop=1.0):
    """Deterministic fake network output for list position gidx."""
    rng = np.random.RandomState(1000 * gidx + int(crop * 100))
    gt = np.zeros((H, W, 1), np.float32)
    gt[4:12, 6:18] = 1.0
    pred = np.clip(gt * 0.8 + rng.rand(H, W, 1).astype(np.float32) * 0.25 - 0.1 * (gidx % 3), 0, 1)
    img = rng.rand(H, W, 3).astype(np.float32) - 0.5
    return pred, gt, img


class _Iter(object):
    def __init__(self, rank, world, per_rank):
        n = len(NAMES)
        gb = per_rank * world
        self.global_names = list(NAMES)
        self.order = [(k * gb + rank * per_rank + j) % n for k in range(-(-n // gb)) for j in range(per_rank)]
        self.pos = 0

    def next(self, k):
        out = [self.order[(self.pos + i) % len(self.order)] for i in range(k)]
        self.pos += k
        return out


class StubLearner(object):
    """Mimics the attributes the scripts read from AdversarialLearner in inference mode."""

    def setup_inference(self, flags, aug_test=False):
        self.rank, self.world = eval_dp.dist_info()
        self.aug_test, self.flags = aug_test, flags
        self.test_crops = [0.85, 0.9, 0.95, 1.0]
        self.test_samples = len(NAMES)
        self.test_iterator = _Iter(self.rank, self.world, 1 if aug_test else flags.batch_size)

    def restore(self, ckpt):
        pass

    def inference(self, sess):
        if self.aug_test:
            g = self.test_iterator.next(1)[0]
            outs = {'pred_masks': {}, 'gt_masks': {}, 'img_1s': {}}
            for c in self.test_crops:
                outs['pred_masks'][c], outs['gt_masks'][c], outs['img_1s'][c] = _frame(g, c)
            return {'outs': outs, 'img_fname': np.array(NAMES[g].encode())}
        idx = self.test_iterator.next(self.flags.batch_size)
        fr = [_frame(g) for g in idx]
        return {'gen_masks': np.stack([f[0] for f in fr]), 'gt_masks': np.stack([f[1] for f in fr]), 'input_image': np.stack([f[2] for f in fr]),
                'gt_flow': np.zeros((len(idx), H, W, 2), np.float32), 'pred_flow': np.zeros((len(idx), H, W, 2), np.float32),
                'img_fname': np.array([NAMES[g].encode() for g in idx])}


def _parse(out_dir, batch):
    from unsupervised_detection_b200.common_flags import FLAGS
    FLAGS(['prog', '--dataset=SYNTHETIC', '--ckpt_file=stub', '--batch_size=%d' % batch, '--generate_visualization', '--test_save_dir=' + out_dir])


def _worker(rank, world, port, out_dir, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), WORLD_SIZE=str(world), RANK=str(rank))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import test_generator as TG
    import test_generator_ensemble as TE
    TG.AdversarialLearner = TE.AdversarialLearner = StubLearner
    _parse(os.path.join(out_dir, 'single'), 2)
    s1 = TG._test_masks_dp()
    _parse(os.path.join(out_dir, 'ens'), 1)
    s2 = TE._test_masks_dp()
    if rank == 0:
        q.put((s1, s2))
    dist.barrier()
    dist.destroy_process_group()


def _listing(d):
    return sorted(os.path.join(os.path.relpath(r, d), f) for r, _, fs in os.walk(d) for f in fs)


def test_two_rank_evaluation_equals_single_process(tmp_path, capsys):
    import scipy.io as sio
    import test_generator as TG
    import test_generator_ensemble as TE
    # ---- single-process reference run of the SAME dp code path (world = 1) and of the original single-process functions
    TG.AdversarialLearner = TE.AdversarialLearner = StubLearner
    ref_dir = str(tmp_path / 'ref')
    _parse(os.path.join(ref_dir, 'single'), 2)
    r1 = TG._test_masks_dp()
    # the original single-process loop with the SAME batch size, which does not divide the list: its last batch wraps around and the
    # head frame is scored twice (the reference does the same); the sharded loop scores the same virtual sequence
    assert len(NAMES) % 2 == 1
    _parse(str(tmp_path / 'orig'), 2)
    TG._test_masks()
    out = capsys.readouterr().out
    _parse(os.path.join(ref_dir, 'ens'), 1)
    r2 = TE._test_masks_dp()
    avg = [l for l in out.splitlines() if l.startswith('The Average over the dataset')]
    nums = [[float(t) for t in l.replace('The Average over the dataset: IoU is ', '').split(' and MAE is ')] for l in avg]
    assert len(avg) == 2 and np.allclose(nums[0], nums[1], rtol=0, atol=1e-6)      # dp report (world 1) == original loop (fp32 vs fp64 sums)
    assert [s[0] for s in r1] == list(range(len(NAMES) + 1))          # 11 frames + the wrap-around duplicate of frame 0
    assert r1[-1][1:] == r1[0][1:] and [s[1] for s in r1].count('catA') == 8
    # ---- two gloo ranks
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    dp_dir = str(tmp_path / 'dp')
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, dp_dir, q)) for r in range(2)]
    for p in procs:
        p.start()
    s1, s2 = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert s1 == r1 and s2 == r2                                   # every frame exactly once, same scores, same order after the merge
    assert _listing(dp_dir) == _listing(ref_dir)                   # same files: per-category frame numbering is global, not per rank
    a = sio.loadmat(os.path.join(dp_dir, 'ens', 'catA', 'result_6.mat'))
    b = sio.loadmat(os.path.join(ref_dir, 'ens', 'catA', 'result_6.mat'))
    assert np.array_equal(a['pred_mask_100'], b['pred_mask_100']) and np.array_equal(a['img_1_085'], b['img_1_085'])


def test_ownership_helpers():
    assert eval_dp.steps_for(11, 1, 2) == 6 and eval_dp.steps_for(11, 2, 2) == 3 and eval_dp.steps_for(8, 4, 1) == 2
    assert eval_dp.owned_indices(2, 2, 1, 2, 11) == [10, 11]        # 11 = the wrap-around duplicate of frame 0 (batch 2 scores 12)
    assert eval_dp.virtual_total(11, 2) == 12 and eval_dp.virtual_total(11, 1) == 11 and eval_dp.virtual_total(8, 4) == 8
    assert eval_dp.category_counters(['d/a/0', 'd/a/1', 'd/b/0', 'd/a/2']) == [1, 2, 1, 3]
    assert eval_dp.merge_scores([(1, 'a', 0.5, 0.1), (0, 'a', 0.2, 0.3)]) == [(0, 'a', 0.2, 0.3), (1, 'a', 0.5, 0.1)]
    lines = []
    iou, mae = eval_dp.report([(0, 'a', 0.2, 0.3), (1, 'b', 0.6, 0.1)], sequence_average=True, out=lines.append)
    assert abs(iou - 0.4) < 1e-12 and abs(mae - 0.2) < 1e-12 and len(lines) == 5
