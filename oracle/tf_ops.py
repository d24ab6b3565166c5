"""TensorFlow 1.13.1 op semantics restated on torch-CPU (oracle; test infra only).

All public functions take/return NHWC tensors like the reference's TF graph.
Semantics follow SURVEY.md Appendix A; call sites cited per function.
"""
import math
import torch
import torch.nn.functional as F


def same_pad(n_in, k, s=1, d=1):
    """TF 'SAME' padding (asymmetric, extra goes bottom/right).  App. A.2."""
    out = -(-n_in // s)
    total = max((out - 1) * s + (k - 1) * d + 1 - n_in, 0)
    return total // 2, total - total // 2


def conv2d_same(x, w_hwio, stride=1, dilation=1, bias=None):
    """tf.nn.conv2d / tf.layers.conv2d with padding='SAME', NHWC x HWIO.
    Call sites: models/utils/convolution_utils.py:46,81; models/PWCNet/model_pwcnet.py:161."""
    kh, kw = w_hwio.shape[0], w_hwio.shape[1]
    pt, pb = same_pad(x.shape[1], kh, stride, dilation)
    pl, pr = same_pad(x.shape[2], kw, stride, dilation)
    xn = F.pad(x.permute(0, 3, 1, 2), (pl, pr, pt, pb))
    y = F.conv2d(xn, w_hwio.permute(3, 2, 0, 1), bias, stride=stride, dilation=dilation)
    return y.permute(0, 2, 3, 1)


def conv2d_transpose_k4s2(x, w_hwoi, bias=None):
    """tf.layers.conv2d_transpose(x, Cout, 4, 2, 'same'); kernel [kh,kw,Cout,Cin].
    models/PWCNet/model_pwcnet.py:286.  Equals ConvTranspose2d(k4,s2,p1) (App. A.7)."""
    w = w_hwoi.permute(3, 2, 0, 1)  # [Cin, Cout, kh, kw]
    y = F.conv_transpose2d(x.permute(0, 3, 1, 2), w, bias, stride=2, padding=1)
    return y.permute(0, 2, 3, 1)


def elu(x):
    return F.elu(x)


def leaky_relu(x, alpha=0.2):
    """tf.nn.leaky_relu default alpha=0.2 (App. A.4)."""
    return F.leaky_relu(x, alpha)


def resize_bilinear_legacy(x, out_h, out_w):
    """tf.image.resize_images / resize_bilinear, align_corners=False, TF<=1.13 legacy
    (no half-pixel centres): src = dst*in/out, lo=floor, hi=min(lo+1,in-1).  App. A.6.
    Call sites: convolution_utils.py:88, nets.py:108, adversarial_learner.py:87-90,
    model_pwcnet.py:646."""
    n, h, w, c = x.shape
    if (h, w) == (out_h, out_w):
        return x

    def axis(n_in, n_out):
        scale = n_in / n_out
        src = torch.arange(n_out, dtype=torch.float64) * scale
        lo = torch.floor(src).long()
        hi = torch.clamp(lo + 1, max=n_in - 1)
        frac = (src - lo.double()).to(x.dtype)
        return lo, hi, frac

    hl, hh, hf = axis(h, out_h)
    wl, wh, wf = axis(w, out_w)
    top = x[:, hl]
    bot = x[:, hh]
    hf = hf.view(1, -1, 1, 1)
    wf = wf.view(1, 1, -1, 1)
    tl, tr = top[:, :, wl], top[:, :, wh]
    bl, br = bot[:, :, wl], bot[:, :, wh]
    t = tl + (tr - tl) * wf
    b = bl + (br - bl) * wf
    return t + (b - t) * hf


def resize_nn_align_corners(x, out_h, out_w):
    """tf.image.resize_nearest_neighbor(align_corners=True): src=min(roundf(dst*(in-1)/(out-1)),in-1).
    convolution_utils.py:71 via resize():4-24.  App. A.5."""
    n, h, w, c = x.shape

    def idx(n_in, n_out):
        scale = (n_in - 1) / (n_out - 1) if n_out > 1 else 0.0
        scale = float(torch.tensor(scale, dtype=torch.float32))
        d = torch.arange(n_out, dtype=torch.float32)
        # roundf = round half away from zero
        s = torch.floor(d * scale + 0.5)
        return torch.clamp(s.long(), max=n_in - 1)

    return x[:, idx(h, out_h)][:, :, idx(w, out_w)]


def resize_nn_legacy(x, out_h, out_w):
    """tf.image.resize_images(method=NEAREST_NEIGHBOR), align_corners=False: src=floor(dst*in/out).
    adversarial_learner.py:92-94 (GT masks)."""
    n, h, w, c = x.shape

    def idx(n_in, n_out):
        scale = float(torch.tensor(n_in / n_out, dtype=torch.float32))
        d = torch.arange(n_out, dtype=torch.float32)
        return torch.clamp(torch.floor(d * scale).long(), max=n_in - 1)

    return x[:, idx(h, out_h)][:, :, idx(w, out_w)]


BN_EPS = 1e-3


def batch_norm_inference(x, gamma, beta):
    """tf.layers.batch_normalization(x) with defaults (training=False, moving mean 0 / var 1).
    convolution_utils.py:50.  App. A.3: y = gamma * x / sqrt(1 + 1e-3) + beta."""
    return x * (gamma / math.sqrt(1.0 + BN_EPS)) + beta
