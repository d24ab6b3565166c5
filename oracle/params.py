"""Seeded synthetic parameters with the reference's initialisers and variable layout (test infra only).
generator: tf.layers defaults = glorot-uniform kernel / zero bias, BN gamma=1 beta=0 (convolution_utils.py:46-50);
recover: xavier_initializer_conv2d (uniform) / zero bias (convolution_utils.py:78);
PWC-Net: he_normal (model_pwcnet.py:153,476,559); conv2d_transpose default glorot-uniform (:286).
`jitter` perturbs biases/gamma/beta so parity tests exercise them (checkpoints are not available offline)."""
import math
import torch

GEN_LAYERS = [  # name, cin, cout, k   (nets.py:19-36)
    ('conv1', 5, 32, 5), ('conv2_downsample', 32, 64, 3), ('conv3', 64, 64, 3), ('conv4_downsample', 64, 128, 3),
    ('conv5', 128, 128, 3), ('conv6', 128, 128, 3), ('conv7_atrous', 128, 128, 3), ('conv8_atrous', 128, 128, 3),
    ('conv9_atrous', 128, 128, 3), ('conv10_atrous', 128, 128, 3), ('conv11', 128, 128, 3), ('conv12', 128, 128, 3),
    ('conv13_upsample', 128, 64, 3), ('conv14', 64, 64, 3), ('conv15_upsample', 64, 32, 3), ('conv16', 32, 16, 3),
    ('conv17', 16, 2, 3)]


def rec_layers(f=0.25, C=2):
    """nets.py:57-107 shapes [kh,kw,cin,cout]."""
    i = lambda v: int(v)
    L = []
    for pre, cin0 in (('a', 3), ('b', C + 2)):
        L += [(pre + 'conv1', 7, cin0, i(64 * f)), (pre + 'conv2', 5, i(64 * f), i(128 * f)),
              (pre + 'conv3', 5, i(128 * f), i(256 * f)), (pre + 'conv31', 3, i(256 * f), i(256 * f)),
              (pre + 'conv4', 3, i(256 * f), i(512 * f)), (pre + 'conv41', 3, i(512 * f), i(512 * f)),
              (pre + 'conv5', 3, i(512 * f), i(512 * f)), (pre + 'conv51', 3, i(512 * f), i(512 * f)),
              (pre + 'conv6', 3, i(512 * f), i(512 * f))]
    L += [('deconv5', 4, i(512 * 2 * f), i(512 * f)), ('flow5', 3, i(512 * 3 * f), C),
          ('deconv4', 4, i(512 * 3 * f), i(512 * f)), ('upflow4', 4, C, C), ('flow4', 3, i(512 * 3 * f + C), C),
          ('deconv3', 4, i(512 * 3 * f + C), i(256 * f)), ('upflow3', 4, C, C), ('flow3', 3, i(256 * 3 * f + C), C),
          ('deconv2', 4, i(256 * 3 * f + C), i(128 * f)), ('upflow2', 4, C, C), ('flow2', 3, i(128 * 3 * f + C), C),
          ('deconv1', 4, i(128 * 3 * f + C), i(64 * f)), ('upflow1', 4, C, C), ('flow1', 5, i(64 * 3 * f + C), C)]
    return L


def pwc_layers():
    """model_pwcnet.py: names -> (k, cin, cout, transposed)."""
    nc = [None, 16, 32, 64, 96, 128, 196]
    L = []
    cin = 3
    for l in range(1, 7):
        L += [(f'featpyr/conv{l}a', 3, cin, nc[l], False), (f'featpyr/conv{l}aa', 3, nc[l], nc[l], False),
              (f'featpyr/conv{l}b', 3, nc[l], nc[l], False)]
        cin = nc[l]
    for l in range(6, 1, -1):
        c0 = 81 if l == 6 else 81 + nc[l] + 4
        c = c0
        for i, co in enumerate((128, 128, 96, 64, 32)):
            L.append((f'predict_flow/conv{l}_{i}', 3, c, co, False))
            c += co
        L.append((f'predict_flow/flow{l}', 3, c, 2, False))
        cc = c
        for i, co in enumerate((128, 128, 128, 96, 64, 32, 2), start=1):
            L.append((f'ctxt/dc_conv{l}{i}', 3, cc, co, False))
            cc = co
        if l != 2:
            L.append((f'upsample/up_flow{l}', 4, 2, 2, True))
            L.append((f'upsample/up_feat{l}', 4, c, 2, True))
    return L


def _glorot_u(g, kh, kw, cin, cout, dtype):
    lim = math.sqrt(6.0 / (kh * kw * cin + kh * kw * cout))
    return (torch.rand(kh, kw, cin, cout, generator=g, dtype=dtype) * 2 - 1) * lim


def make_params(seed=8964, dtype=torch.float32, jitter=0.0, nets=('MaskNet', 'FlownetS', 'pwcnet')):
    g = torch.Generator().manual_seed(seed)
    p = {}
    jit = lambda n: (torch.rand(n, generator=g, dtype=dtype) * 2 - 1) * jitter
    if 'MaskNet' in nets:
        for name, cin, cout, k in GEN_LAYERS:
            p[f'MaskNet/{name}/kernel'] = _glorot_u(g, k, k, cin, cout, dtype)
            p[f'MaskNet/{name}/bias'] = jit(cout)
            p[f'MaskNet/{name}/gamma'] = 1.0 + jit(cout)
            p[f'MaskNet/{name}/beta'] = jit(cout)
    if 'FlownetS' in nets:
        for name, k, cin, cout in rec_layers():
            p[f'FlownetS/{name}/weights'] = _glorot_u(g, k, k, cin, cout, dtype)
            p[f'FlownetS/{name}/biases'] = jit(cout)
    if 'pwcnet' in nets:
        for name, k, cin, cout, tr in pwc_layers():
            if tr:   # conv2d_transpose kernel [kh,kw,Cout,Cin], glorot-uniform default
                p[f'pwcnet/{name}/kernel'] = _glorot_u(g, k, k, cout, cin, dtype)
            else:    # he_normal (truncated in Keras; plain normal is an adequate synthetic stand-in)
                std = math.sqrt(2.0 / (k * k * cin))
                p[f'pwcnet/{name}/kernel'] = torch.randn(k, k, cin, cout, generator=g, dtype=dtype) * std
            p[f'pwcnet/{name}/bias'] = jit(cout)
    return p


def count(p, scope):
    return sum(v.numel() for n, v in p.items() if n.startswith(scope))
