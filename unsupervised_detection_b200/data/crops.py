"""Central multi-crop of a frame pair (data/davis2016_data_utils.py:328-354 augmented_inputs): each crop fraction is cut
from the centre and resized back to the working resolution with the legacy bilinear rule.  Host side."""
import torch
import torch.nn.functional as F


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
    n, h, w, _ = img1.shape
    o1, o2, og = [], [], []
    for c in crops:
        ch, cw = int(h * c), int(w * c)
        y0, x0 = (h - ch) // 2, (w - cw) // 2
        sl = (slice(None), slice(y0, y0 + ch), slice(x0, x0 + cw))
        o1.append(_legacy_resize(img1[sl], h, w))
        o2.append(_legacy_resize(img2[sl], h, w))
        og.append(F.interpolate(gt[sl].permute(0, 3, 1, 2), size=(h, w), mode='nearest').permute(0, 2, 3, 1))
    return torch.cat(o1), torch.cat(o2), torch.cat(og)
