"""Batch-sharded evaluation (BASELINE config 4 / SURVEY 8e: "frames round-robin, no collective except a final gather of IoU sums").

Under torchrun every rank evaluates its contiguous slice of every global batch of the ordered test list
(`data.davis2016_data_utils._Iter.shard`) and the per-category score lists are merged once at the end with one
`all_gather_object`; rank 0 prints the report the single-process scripts print.  The helpers here are free of GPU calls, so the
sharding / ownership / merge logic is exercised on CPU with gloo (tests/test_eval_dp_gloo.py)."""
import os


def dist_info():
    """(rank, world) of an initialised torch.distributed job, else (0, 1)."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return 0, 1


def owned_indices(step, local_n, rank, world, total):
    """Global list positions of the `local_n` samples rank `rank` reads in step `step` (global batch = local_n * world); position g is
    frame g % total of the endless iterator; positions >= virtual_total(total, batch) are beyond what the single-process script scores."""
    base = step * local_n * world + rank * local_n
    return [base + j for j in range(local_n)]


def steps_for(total, local_n, world):
    gb = local_n * world
    return -(-total // gb)


def virtual_total(total, batch):
    """Number of frames the single-process script scores: ceil(total / batch) batches of the endless ordered iterator, i.e. the last
    batch wraps around and its head frames are scored (and written) a second time -- the reference does exactly that
    (test_generator.py:62-64 with dataset.repeat()).  The sharded loop scores the same virtual sequence, position g -> frame g % total,
    so 1-rank and N-rank reports agree with the single-process one for any total % batch."""
    return -(-total // batch) * batch


def category_counters(names):
    """Per-frame running index inside its category in GLOBAL list order (what `len(CategoryIou[category])` is in the single-process
    scripts at the moment a frame is written) -> list parallel to `names`."""
    seen, out = {}, []
    for n in names:
        c = n.split('/')[-2]
        seen[c] = seen.get(c, 0) + 1
        out.append(seen[c])
    return out


def merge_scores(local):
    """local: list of (global_index, category, iou, mae).  Returns the same list for ALL frames of the job, sorted by global index,
    on every rank (one all_gather_object; identity when not distributed)."""
    rank, world = dist_info()
    if world == 1:
        return sorted(local)
    import torch.distributed as dist
    parts = [None] * world
    dist.all_gather_object(parts, local)
    return sorted(x for p in parts for x in p)


def report(scores, sequence_average=False, out=print):
    """The summary lines of test_generator.py:120-132 / test_generator_ensemble.py:113-122 from merged (index, category, iou, mae)."""
    import numpy as np
    cat_iou, cat_mae = {}, {}
    for _, c, iou, mae in scores:
        cat_iou.setdefault(c, []).append(iou)
        cat_mae.setdefault(c, []).append(mae)
    tot_i = tot_m = 0.0
    per_cat = []
    for c, li in cat_iou.items():
        out("Category {}: IoU is {} and MAE is {}".format(c, np.mean(li), np.mean(cat_mae[c])))
        tot_i += np.sum(li)
        tot_m += np.sum(cat_mae[c])
        per_cat.append(np.mean(li))
    n = len(scores)
    out("The Average over the dataset: IoU is {} and MAE is {}".format(tot_i / float(n), tot_m / float(n)))
    if sequence_average:
        out("The Average over sequences IoU is {}".format(np.mean(per_cat)))
    out("Success: Processed {} frames".format(n))
    return tot_i / float(n), tot_m / float(n)


def global_names(learner):
    """First-frame file names of the whole ordered test list (set by the reader's shard())."""
    names = getattr(learner.test_iterator, 'global_names', None)
    if not names:
        raise RuntimeError('batch-sharded evaluation needs an ordered dataset reader (DAVIS2016 / FBMS / SEGTRACK); the %s reader has no '
                           'fixed sample list' % type(learner.test_iterator).__name__)
    return names


def is_distributed_launch():
    return int(os.environ.get('WORLD_SIZE', '1')) > 1
