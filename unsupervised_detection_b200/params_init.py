"""Seeded parameter initialisation with the reference's initialisers and variable names.
generator: tf.layers.conv2d defaults (glorot-uniform kernel, zero bias) + BN gamma=1/beta=0 (convolution_utils.py:46-50);
recover: xavier_initializer_conv2d + zero bias (convolution_utils.py:78);  PWC-Net: he_normal (model_pwcnet.py:153), its
conv2d_transpose layers glorot-uniform (:286).  Real use restores PWC-Net from --flow_ckpt."""
import math
import torch

from .models.nets import GEN_LAYERS, rec_layer_table


def _glorot(g, kh, kw, cin, cout):
    lim = math.sqrt(6.0 / (kh * kw * (cin + cout)))
    return (torch.rand(kh, kw, cin, cout, generator=g) * 2 - 1) * lim


def init_generator(seed=8964):
    g = torch.Generator().manual_seed(seed)
    p = {}
    for name, cin, cout, k, _, _ in GEN_LAYERS:
        p['MaskNet/%s/kernel' % name] = _glorot(g, k, k, cin, cout)
        p['MaskNet/%s/bias' % name] = torch.zeros(cout)
        p['MaskNet/%s/gamma' % name] = torch.ones(cout)
        p['MaskNet/%s/beta' % name] = torch.zeros(cout)
    return p


def init_recover(seed=8965):
    g = torch.Generator().manual_seed(seed)
    p = {}
    for name, k, cin, cout, _, _ in rec_layer_table():
        p['FlownetS/%s/weights' % name] = _glorot(g, k, k, cin, cout)
        p['FlownetS/%s/biases' % name] = torch.zeros(cout)
    return p


def init_pwcnet(store_entries, seed=8966):
    """he_normal for conv kernels, glorot-uniform for the transposed ones, zero biases (synthetic stand-in for a ckpt)."""
    g = torch.Generator().manual_seed(seed)
    p = {}
    for name, shape, _, _, _ in store_entries:
        if name.endswith('/kernel'):
            kh, kw, a, b = shape
            if '/upsample/' in name:
                p[name] = _glorot(g, kh, kw, a, b)
            else:
                p[name] = torch.randn(*shape, generator=g) * math.sqrt(2.0 / (kh * kw * a))
        else:
            p[name] = torch.zeros(*shape)
    return p
