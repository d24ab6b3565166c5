"""Golden vectors for the host-side metric / dump helpers: executes the REFERENCE's own numpy functions in this container --
`compute_boundary_score`, `postprocess_image`, `postprocess_mask` (models/utils/general_utils.py:22-51,117-132) and
`compute_IoU`, `compute_mae` (test_generator.py:19-40) -- by compiling only those function definitions out of the reference files
(the TensorFlow / gflags / keras imports at the top of those files are never executed).  `np.bool` (removed from numpy >= 1.24) is
aliased to `bool` for the run.  Usage: python tests/golden/make_golden_metrics.py  (needs /root/reference; the .npz is committed)."""
import ast
import os

import cv2
import numpy as np


def load_functions(path, names, extra=None):
    tree = ast.parse(open(path).read())
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    ns = {'np': np, 'cv2': cv2}
    ns.update(extra or {})
    exec(compile(ast.Module(body=body, type_ignores=[]), path, 'exec'), ns)
    return ns


if not hasattr(np, 'bool'):
    np.bool = bool
gu = load_functions('/root/reference/models/utils/general_utils.py', {'compute_boundary_score', 'postprocess_image', 'postprocess_mask'})
tg = load_functions('/root/reference/test_generator.py', {'compute_IoU', 'compute_mae'},
                    {'compute_boundary_score': gu['compute_boundary_score'], 'mask_threshold': 0.6})

rng = np.random.RandomState(77)
H, W = 24, 40
preds, gts, ious, anns, maes, scores = [], [], [], [], [], []
for case in range(8):
    pred = rng.rand(H, W, 1).astype(np.float32) * 0.3
    gt = np.zeros((H, W, 1), np.float32)
    if case == 0:                                   # compact object away from the borders
        pred[6:16, 10:25] = 0.9
        gt[5:15, 12:26] = 1.0
        pred[pred < 0.3] *= 0.2
    elif case == 1:                                 # mask hugging the borders -> complemented
        pred[:] = 0.8
        pred[8:14, 15:22] = 0.01
        gt[8:14, 15:22] = 1.0
    elif case == 2:                                 # both empty -> the bare `1` return value
        pred[:] = 0.0
    elif case == 3:                                 # empty prediction, non-empty ground truth
        pred[:] = 0.05
        gt[3:9, 3:9] = 1.0
    elif case == 4:                                 # border score just under / over 0.6 is decided by the frame pixels
        pred[:] = 0.0
        pred[0:2, :] = 1.0
        pred[:, 0:2] = 1.0
        pred[H - 2:H, : W // 2] = 1.0
        gt[0:4, 0:10] = 1.0
    else:
        pred = rng.rand(H, W, 1).astype(np.float32)
        gt = (rng.rand(H, W, 1) > 0.6).astype(np.float32)
    r = tg['compute_IoU'](gt_mask=gt.copy(), pred_mask_f=pred.copy())
    if isinstance(r, tuple):
        iou, ann = float(r[0]), r[1].astype(np.uint8)
    else:                                           # reference arity quirk: plain 1 when both masks are empty
        iou, ann = float(r), np.zeros((H, W, 1), np.uint8)
    preds.append(pred); gts.append(gt); ious.append(iou); anns.append(ann)
    maes.append(float(tg['compute_mae'](gt, pred)))
    scores.append(float(gu['compute_boundary_score'](pred > 0.1)))
img = rng.rand(H, W, 3).astype(np.float32) - 0.5
msk = rng.rand(H, W, 1).astype(np.float32)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'metrics.npz'), pred=np.stack(preds), gt=np.stack(gts),
                    iou=np.array(ious, np.float64), ann=np.stack(anns), mae=np.array(maes, np.float64), score=np.array(scores, np.float64),
                    img=img, img_out=gu['postprocess_image'](img.copy()), msk=msk, msk_out=gu['postprocess_mask'](msk.copy()))
print('metrics.npz', ious, scores)
