"""Host-side metric / dump helpers pinned to vectors produced by the reference's own numpy code
(tests/golden/make_golden_metrics.py runs general_utils.py:22-51,117-132 and test_generator.py:19-40 in this container)."""
import os

import numpy as np

from unsupervised_detection_b200.models.utils import general_utils as G

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'metrics.npz'))


def test_iou_mae_boundary_score_match_reference_vectors():
    n = GOLD['pred'].shape[0]
    for i in range(n):
        pred, gt = GOLD['pred'][i], GOLD['gt'][i]
        assert G.compute_boundary_score(pred > 0.1) == GOLD['score'][i]
        iou, ann = G.compute_IoU(gt.copy(), pred.copy())
        assert abs(float(iou) - GOLD['iou'][i]) < 1e-7, i
        if not (GOLD['ann'][i].sum() == 0 and GOLD['iou'][i] == 1.0):      # case 2: the reference returns a bare 1, no annotation
            assert np.array_equal(np.asarray(ann, np.uint8), GOLD['ann'][i]), i
        assert abs(G.compute_mae(gt, pred) - GOLD['mae'][i]) < 1e-7
    # the batched validation metric (compute_all_IoU = TF twin of the same rule, adversarial_learner.py:135) agrees with the
    # per-image numpy one wherever the union is non-empty
    all_iou = G.compute_all_IoU(GOLD['pred'], GOLD['gt'])
    for i in range(n):
        if i != 2:
            assert abs(float(all_iou[i]) - GOLD['iou'][i]) < 1e-6, i
    assert all_iou[2] == 0.0                                               # 0 / (0 + 1e-8)


def test_postprocess_helpers_match_reference_vectors():
    assert np.array_equal(G.postprocess_image(GOLD['img'].copy()), GOLD['img_out'])
    assert np.array_equal(G.postprocess_mask(GOLD['msk'].copy()), GOLD['msk_out'])
