"""Mask-IoU metric of the reference (models/utils/general_utils.py:89-150, test_generator.py:19-40), host side (numpy).
The metric runs on the D2H copy of the generated masks exactly like test_generator.py does; it is not a device kernel."""
import numpy as np


def compute_boundary_score(segmentation):
    """general_utils.py:117-132: fraction of the 2-px image frame covered by the mask (corners counted twice)."""
    H, W = segmentation.shape[0], segmentation.shape[1]
    up, bottom = segmentation[0:2, :], segmentation[H - 2:H, :]
    left, right = segmentation[:, 0:2], segmentation[:, W - 2:W]
    occ = np.sum(up) + np.sum(bottom) + np.sum(left) + np.sum(right)
    return occ / (1.0 * (up.size + bottom.size + left.size + right.size))


def disambiguate_forw_back(pred_masks, threshold=0.1):
    """general_utils.py:100-109 on a batch [B,H,W,1]: complement the mask when it hugs the borders (score >= 0.6)."""
    pm = (pred_masks > threshold).astype(np.float32)
    out = np.empty_like(pm)
    for b in range(pm.shape[0]):
        score = compute_boundary_score(pm[b])
        out[b] = pm[b] if score < 0.6 else 1.0 - pm[b]
    return out


def compute_all_IoU(pred_masks, gt_masks, threshold=0.1):
    """general_utils.py:111-115 + tf_iou_computation :89-98 -> [B] (epsilon 1e-8 in the union)."""
    gt = gt_masks > 0.01
    obj = disambiguate_forw_back(pred_masks, threshold) > 0.5
    union = np.sum(gt | obj, axis=(1, 2, 3)).astype(np.float32) + 1e-8
    return np.sum(gt & obj, axis=(1, 2, 3)).astype(np.float32) / union


def compute_IoU(gt_mask, pred_mask_f, threshold=0.1, mask_threshold=0.6):
    """test_generator.py:19-35.  The reference returns a bare `1` when both masks are empty (arity bug); the value is
    kept and the arity fixed to (iou, annotation)."""
    gt_mask = gt_mask.astype(bool)
    pred_mask = pred_mask_f > threshold
    annotation = pred_mask if compute_boundary_score(pred_mask) < mask_threshold else np.logical_not(pred_mask)
    if np.isclose(np.sum(annotation), 0) and np.isclose(np.sum(gt_mask), 0):
        return 1.0, annotation
    return np.sum((annotation & gt_mask)) / np.sum((annotation | gt_mask), dtype=np.float32), annotation


def compute_mae(gt_mask, pred_mask_f):
    """test_generator.py:38-40."""
    return np.mean(np.abs(gt_mask.astype(np.float32) - pred_mask_f.astype(np.float32)))


# ---- dump helpers of test_generator*.py (models/utils/general_utils.py:22-51 of the reference), host side
def postprocess_image(image):
    """[H,W,3] in [-0.5,0.5] RGB -> uint8 BGR (general_utils.py:22-35)."""
    import cv2
    un = np.asarray((image + 0.5) * 255, np.uint8)
    return cv2.cvtColor(un, cv2.COLOR_RGB2BGR)


def postprocess_mask(mask):
    """[H,W,1] in [0,1] -> uint8 [H,W,3] with the mask in the green slot (general_utils.py:37-51)."""
    un = np.asarray(mask * 255.0, np.uint8)
    tile = np.zeros_like(un, dtype=np.uint8)
    return np.concatenate((tile, un, tile), axis=-1)
