"""Oracle restatement of models/PWCNet/{model_pwcnet,core_warp,core_costvol}.py (test infra only).
Forward only (PWC-Net is frozen: adversarial_learner.py:211-234)."""
import torch
from . import tf_ops as T

NUM_CHANN = [None, 16, 32, 64, 96, 128, 196]      # model_pwcnet.py:151
PYR_LVLS, FLOW_PRED_LVL, SEARCH_RANGE = 6, 2, 4   # model_pwcnet.py:8-19


def dense_image_warp(image, flow):
    """core_warp.py:153-202 + _interpolate_bilinear :42-150.  out[b,j,i] = bilerp(image, j-flow0, i-flow1)."""
    b, h, w, c = image.shape
    gy, gx = torch.meshgrid(torch.arange(h, dtype=flow.dtype), torch.arange(w, dtype=flow.dtype), indexing='ij')
    qy = gy.unsqueeze(0) - flow[..., 0]
    qx = gx.unsqueeze(0) - flow[..., 1]

    def prep(q, size):
        fl = torch.clamp(torch.floor(q), 0.0, float(size - 2))       # core_warp.py:104-105
        alpha = torch.clamp(q - fl, 0.0, 1.0)                        # core_warp.py:112-115
        return fl.long(), alpha.unsqueeze(-1)

    fy, ay = prep(qy, h)
    fx, ax = prep(qx, w)
    flat = image.reshape(b, h * w, c)

    def gather(yy, xx):
        lin = (yy * w + xx).reshape(b, h * w, 1).expand(-1, -1, c)
        return torch.gather(flat, 1, lin).reshape(b, h, w, c)

    tl, tr = gather(fy, fx), gather(fy, fx + 1)
    bl, br = gather(fy + 1, fx), gather(fy + 1, fx + 1)
    top = ax * (tr - tl) + tl                                        # core_warp.py:146-148
    bot = ax * (br - bl) + bl
    return ay * (bot - top) + top


def cost_volume(c1, warp, search_range=SEARCH_RANGE):
    """core_costvol.py:20-40: channel 9*dy+dx = mean_c c1 * warp_pad[.., i+dy, j+dx], then leaky 0.1."""
    import torch.nn.functional as F
    r = search_range
    padded = F.pad(warp, (0, 0, r, r, r, r))
    _, h, w, _ = c1.shape
    out = []
    for y in range(2 * r + 1):
        for x in range(2 * r + 1):
            out.append((c1 * padded[:, y:y + h, x:x + w]).mean(dim=3, keepdim=True))
    return T.leaky_relu(torch.cat(out, 3), 0.1)


def _c(x, p, name, stride=1, dil=1, act=True):
    y = T.conv2d_same(x, p[name + '/kernel'], stride, dil, p[name + '/bias'])
    return T.leaky_relu(y, 0.1) if act else y


def extract_features(img, p):
    """model_pwcnet.py:149-168 (one frame; weights shared between frames)."""
    pyr = [None]
    x = img
    for lvl in range(1, PYR_LVLS + 1):
        x = _c(x, p, f'pwcnet/featpyr/conv{lvl}a', 2)
        x = _c(x, p, f'pwcnet/featpyr/conv{lvl}aa')
        x = _c(x, p, f'pwcnet/featpyr/conv{lvl}b')
        pyr.append(x)
    return pyr


def predict_flow(corr, c1, up_flow, up_feat, lvl, p):
    """model_pwcnet.py:476-506 (dense connections: new activation concatenated in FRONT)."""
    x = corr if c1 is None else torch.cat([corr, c1, up_flow, up_feat], 3)
    for i in range(5):
        act = _c(x, p, f'pwcnet/predict_flow/conv{lvl}_{i}')
        x = torch.cat([act, x], 3)
    flow = _c(x, p, f'pwcnet/predict_flow/flow{lvl}', act=False)
    return x, flow


def refine_flow(feat, flow, lvl, p):
    """model_pwcnet.py:559-576."""
    x = feat
    for i, d in enumerate((1, 2, 4, 8, 16, 1), start=1):
        x = _c(x, p, f'pwcnet/ctxt/dc_conv{lvl}{i}', 1, d)
    x = _c(x, p, f'pwcnet/ctxt/dc_conv{lvl}7', act=False)
    return flow + x


def predict_from_img_pairs(img1, img2, p, return_pyr=False):
    """model_pwcnet.py:61-76 -> adapt_x :39-56 -> nn :581-649."""
    c1 = extract_features(img1 + 0.5, p)
    c2 = extract_features(img2 + 0.5, p)
    flow_pyr = []
    up_flow = up_feat = None
    for lvl in range(PYR_LVLS, FLOW_PRED_LVL - 1, -1):
        if lvl == PYR_LVLS:
            corr = cost_volume(c1[lvl], c2[lvl])
            upfeat, flow = predict_flow(corr, None, None, None, lvl, p)
        else:
            scaler = 20.0 / 2 ** lvl
            warp = dense_image_warp(c2[lvl], up_flow * scaler)
            corr = cost_volume(c1[lvl], warp)
            upfeat, flow = predict_flow(corr, c1[lvl], up_flow, up_feat, lvl, p)
        flow = refine_flow(upfeat, flow, lvl, p)
        flow_pyr.append(flow)
        if lvl != FLOW_PRED_LVL:
            up_flow = T.conv2d_transpose_k4s2(flow, p[f'pwcnet/upsample/up_flow{lvl}/kernel'],
                                              p[f'pwcnet/upsample/up_flow{lvl}/bias'])
            up_feat = T.conv2d_transpose_k4s2(upfeat, p[f'pwcnet/upsample/up_feat{lvl}/kernel'],
                                              p[f'pwcnet/upsample/up_feat{lvl}/bias'])
        else:
            s = 2 ** FLOW_PRED_LVL
            flow_pred = T.resize_bilinear_legacy(flow, flow.shape[1] * s, flow.shape[2] * s) * s
    if return_pyr:
        return flow_pred, flow_pyr, c1, c2
    return flow_pred
