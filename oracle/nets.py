"""Oracle restatement of models/nets.py + models/utils/convolution_utils.py (test infra only)."""
import torch
from . import tf_ops as T

# (name, cout, ksize, stride, rate) for the plain gen_conv layers; see GEN_LAYERS in the
# product package for the same table (kept separate on purpose: the oracle must not
# share code with the thing it checks).
_CNUM = 32


def gen_conv(x, p, scope, name, ksize, stride=1, rate=1, activation=T.elu):
    """convolution_utils.py:26-53: conv2d SAME (+bias) -> BN(inference) -> activation."""
    y = T.conv2d_same(x, p[f'{scope}/{name}/kernel'], stride, rate, p[f'{scope}/{name}/bias'])
    y = T.batch_norm_inference(y, p[f'{scope}/{name}/gamma'], p[f'{scope}/{name}/beta'])
    return activation(y)


def gen_deconv(x, p, scope, name):
    """convolution_utils.py:55-75: NN x2 (align_corners=True) then 3x3 gen_conv '<name>/<name>_conv'."""
    x = T.resize_nn_align_corners(x, 2 * x.shape[1], 2 * x.shape[2])
    return gen_conv(x, p, scope, name, 3, 1)


def generator_net(images, flows, p, scope='MaskNet', return_logits=False):
    """models/nets.py:4-42.  images [B,H,W,3], flows [B,H,W,2] -> mask [B,H,W,1]."""
    c = _CNUM
    x = torch.cat((images, flows), 3)
    x_0 = gen_conv(x, p, scope, 'conv1', 5, 1)
    x = gen_conv(x_0, p, scope, 'conv2_downsample', 3, 2)
    x_1 = gen_conv(x, p, scope, 'conv3', 3, 1)
    x = gen_conv(x_1, p, scope, 'conv4_downsample', 3, 2)
    x = gen_conv(x, p, scope, 'conv5', 3, 1)
    x_2 = gen_conv(x, p, scope, 'conv6', 3, 1)
    x = gen_conv(x_2, p, scope, 'conv7_atrous', 3, rate=2)
    x = gen_conv(x, p, scope, 'conv8_atrous', 3, rate=4)
    x = gen_conv(x, p, scope, 'conv9_atrous', 3, rate=8)
    x = gen_conv(x, p, scope, 'conv10_atrous', 3, rate=16)
    x = gen_conv(x, p, scope, 'conv11', 3, 1) + x_2
    x = gen_conv(x, p, scope, 'conv12', 3, 1)
    x = gen_deconv(x, p, scope, 'conv13_upsample')
    x = gen_conv(x, p, scope, 'conv14', 3, 1) + x_1
    x = gen_deconv(x, p, scope, 'conv15_upsample') + x_0
    x = gen_conv(x, p, scope, 'conv16', 3, 1)
    x = gen_conv(x, p, scope, 'conv17', 3, 1, activation=lambda t: t)
    if return_logits:
        return x
    x = x / 10.0                               # nets.py:38
    m = torch.softmax(x, dim=-1)               # nets.py:39
    return m[..., 0:1]                         # nets.py:41


def conv(x, p, scope, name, stride, activation=T.leaky_relu):
    """convolution_utils.py:77-85: tf.nn.conv2d SAME + bias + leaky_relu(0.2)."""
    y = T.conv2d_same(x, p[f'{scope}/{name}/weights'], stride, 1, p[f'{scope}/{name}/biases'])
    return activation(y)


def deconv(x, p, scope, name, size, activation=T.leaky_relu):
    """convolution_utils.py:87-90: legacy bilinear resize to `size`, then conv stride 1."""
    x = T.resize_bilinear_legacy(x, size[0], size[1])
    return conv(x, p, scope, name, 1, activation)


def recover_net(img1, flow_masked, mask, p, scope='FlownetS', return_pyramid=False):
    """models/nets.py:45-110."""
    ident = lambda t: t
    ones_x = torch.ones_like(flow_masked)[..., 0:1]
    fm = torch.cat([flow_masked, ones_x, 1.0 - mask], 3)          # nets.py:52
    a = {}
    b = {}
    for pre, src, d in (('a', img1, a), ('b', fm, b)):
        x = conv(src, p, scope, pre + 'conv1', 2); d['1'] = x
        x = conv(x, p, scope, pre + 'conv2', 2); d['2'] = x
        x = conv(x, p, scope, pre + 'conv3', 2)
        x = conv(x, p, scope, pre + 'conv31', 1); d['31'] = x
        x = conv(x, p, scope, pre + 'conv4', 2)
        x = conv(x, p, scope, pre + 'conv41', 1); d['41'] = x
        x = conv(x, p, scope, pre + 'conv5', 2)
        x = conv(x, p, scope, pre + 'conv51', 1); d['51'] = x
        x = conv(x, p, scope, pre + 'conv6', 2); d['6'] = x
    sz = lambda t: (t.shape[1], t.shape[2])
    conv6 = torch.cat((a['6'], b['6']), 3)                         # nets.py:78
    deconv5 = deconv(conv6, p, scope, 'deconv5', sz(b['51']))
    concat5 = torch.cat((deconv5, b['51'], a['51']), 3)
    flow5 = conv(concat5, p, scope, 'flow5', 1, ident)
    deconv4 = deconv(concat5, p, scope, 'deconv4', sz(b['41']))
    upflow4 = deconv(flow5, p, scope, 'upflow4', sz(b['41']), ident)
    concat4 = torch.cat((deconv4, b['41'], a['41'], upflow4), 3)
    flow4 = conv(concat4, p, scope, 'flow4', 1, ident)
    deconv3 = deconv(concat4, p, scope, 'deconv3', sz(b['31']))
    upflow3 = deconv(flow4, p, scope, 'upflow3', sz(b['31']), ident)
    concat3 = torch.cat((deconv3, b['31'], a['31'], upflow3), 3)
    flow3 = conv(concat3, p, scope, 'flow3', 1, ident)
    deconv2 = deconv(concat3, p, scope, 'deconv2', sz(b['2']))
    upflow2 = deconv(flow3, p, scope, 'upflow2', sz(b['2']), ident)
    concat2 = torch.cat((deconv2, b['2'], a['2'], upflow2), 3)
    flow2 = conv(concat2, p, scope, 'flow2', 1, ident)
    deconv1 = deconv(concat2, p, scope, 'deconv1', sz(b['1']))
    upflow1 = deconv(flow2, p, scope, 'upflow1', sz(b['1']), ident)
    concat1 = torch.cat((deconv1, b['1'], a['1'], upflow1), 3)
    flow1 = conv(concat1, p, scope, 'flow1', 1, ident)
    pred = T.resize_bilinear_legacy(flow1, img1.shape[1], img1.shape[2])   # nets.py:108
    if return_pyramid:
        return pred, dict(flow1=flow1, flow2=flow2, flow3=flow3, flow4=flow4, flow5=flow5)
    return pred
