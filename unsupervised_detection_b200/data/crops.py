"""Central multi-crop of a frame pair (data/davis2016_data_utils.py:328-354 augmented_inputs): each crop fraction is cut
from the centre and resized back to the working resolution with the legacy bilinear rule.  Host side."""
import torch


def _legacy_resize(x, oh, ow):
    """tf.image.resize_images legacy bilinear (App. A.6) on an NHWC CPU tensor."""
    n, h, w, c = x.shape

    def ax(n_in, n_out):
        s = torch.arange(n_out, dtype=torch.float32) * (n_in / n_out)
        lo = torch.floor(s).long()
        return lo, torch.clamp(lo + 1, max=n_in - 1), (s - lo.float())
    hl, hh, hf = ax(h, oh)
    wl, wh, wf = ax(w, ow)
    t, b = x[:, hl], x[:, hh]
    wf_, hf_ = wf.view(1, 1, -1, 1), hf.view(1, -1, 1, 1)
    top = t[:, :, wl] + (t[:, :, wh] - t[:, :, wl]) * wf_
    bot = b[:, :, wl] + (b[:, :, wh] - b[:, :, wl]) * wf_
    return top + (bot - top) * hf_


def central_crops(img1, img2, gt, crops):
    """img1, img2 [N,H,W,3], gt [N,H,W,1] -> the crops stacked along the batch axis.  Geometry of tf.image.central_crop
    (offset = int((dim - dim*frac)/2), size = dim - 2*offset); all three tensors go back to HxW with the legacy bilinear rule --
    the reference's `central_cropping` resizes the mask bilinearly as well (davis2016_data_utils.py:130-134)."""
    from .davis2016_data_utils import central_crop_box
    n, h, w, _ = img1.shape
    o1, o2, og = [], [], []
    for c in crops:
        y0, x0, ch, cw = central_crop_box(h, w, c)
        sl = (slice(None), slice(y0, y0 + ch), slice(x0, x0 + cw))
        o1.append(_legacy_resize(img1[sl], h, w))
        o2.append(_legacy_resize(img2[sl], h, w))
        og.append(_legacy_resize(gt[sl], h, w))
    return torch.cat(o1), torch.cat(o2), torch.cat(og)
