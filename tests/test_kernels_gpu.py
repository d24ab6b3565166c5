"""GPU parity of the HBM-bound kernels against the oracle (fp32 reference of the same op on bf16-rounded inputs), incl.
the edge cases the domain has: flows pointing far outside the image, ragged tiles, odd sizes, non-integer resize ratios."""
import os
import numpy as np
import pytest
import torch

from oracle import tf_ops as T, pwcnet as PW
from unsupervised_detection_b200 import _lib
from unsupervised_detection_b200.engine import Act

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
bf = lambda x: x.to(torch.bfloat16).float()
ST = lambda: torch.cuda.current_stream().cuda_stream


def act_from(x, pitch=None, c_off=0):
    n, h, w, c = x.shape
    c8 = (c + 7) // 8 * 8
    pitch = pitch or c8
    buf = torch.zeros(n, h, w, pitch, dtype=torch.bfloat16, device='cuda')
    buf[..., c_off:c_off + c] = x.cuda().to(torch.bfloat16)
    return Act(n, h, w, c, 'cuda', buf=buf, c_off=c_off)


@pytest.mark.parametrize('B,h,w,C,use_flow,amp', [(2, 6, 10, 196, False, 0), (2, 12, 20, 128, True, 2.0), (1, 24, 40, 96, True, 30.0),
                                                  (1, 9, 17, 32, True, 1.0), (3, 48, 80, 64, True, 3.0)])
def test_warp_costvol(B, h, w, C, use_flow, amp):
    g = torch.Generator().manual_seed(h * w + C)
    c1, c2 = bf(torch.randn(B, h, w, C, generator=g)), bf(torch.randn(B, h, w, C, generator=g))
    flow = torch.randn(B, h, w, 2, generator=g) * amp if use_flow else None
    a1, a2 = act_from(c1, pitch=(C + 7) // 8 * 8 + 16, c_off=8), act_from(c2)
    out = torch.zeros(B, h, w, 96, dtype=torch.bfloat16, device='cuda')
    fd = flow.cuda() if use_flow else None
    _lib.call('cis_warp_costvol', a1.ptr, a1.pitch, a1.c_off, a2.ptr, a2.pitch, a2.c_off, fd.data_ptr() if use_flow else None, 1.25, B, h, w, C,
              out.data_ptr(), 96, 8, ST())
    torch.cuda.synchronize()
    warp = PW.dense_image_warp(c2, flow * 1.25) if use_flow else c2
    ref = PW.cost_volume(c1, warp)
    got = out[..., 8:89].float().cpu()
    assert float((got - ref).abs().max()) <= 2 ** -8 * float(ref.abs().max()) + 2e-3
    assert float(out[..., :8].abs().max()) == 0 and float(out[..., 89:].abs().max()) == 0    # neighbours untouched


def test_warp_costvol_golden_and_standalone_warp():
    z = np.load(os.path.join(G, 'warp_costvol_6x7.npz'))
    c1, c2, fl = (torch.from_numpy(z[k]).float() for k in ('c1', 'c2', 'flow'))
    a1, a2 = act_from(c1), act_from(c2)
    out = torch.zeros(1, 6, 7, 88, dtype=torch.bfloat16, device='cuda')
    fd = fl.cuda()
    _lib.call('cis_warp_costvol', a1.ptr, 8, 0, a2.ptr, 8, 0, fd.data_ptr(), 1.0, 1, 6, 7, 8, out.data_ptr(), 88, 0, ST())
    wo = torch.zeros(1, 6, 7, 8, dtype=torch.bfloat16, device='cuda')
    _lib.call('cis_dense_image_warp', a2.ptr, 8, 0, fd.data_ptr(), 1.0, 1, 6, 7, 8, wo.data_ptr(), 8, ST())
    torch.cuda.synchronize()
    assert float((wo.float().cpu() - torch.from_numpy(z['warp']).float()).abs().max()) <= 0.03   # bf16 inputs/outputs of O(3)
    assert float((out[..., :81].float().cpu() - torch.from_numpy(z['cv']).float()).abs().max()) <= 0.03


@pytest.mark.parametrize('H,W,OH,OW', [(4, 7, 8, 14), (2, 4, 4, 7), (8, 14, 16, 28), (5, 3, 7, 9), (6, 6, 6, 6)])
def test_resize_bilinear_bf16_and_transpose(H, W, OH, OW):
    g = torch.Generator().manual_seed(H * OW)
    x = bf(torch.randn(2, H, W, 24, generator=g))
    a = act_from(x, pitch=40, c_off=8)
    out = torch.zeros(2, OH, OW, 24, dtype=torch.bfloat16, device='cuda')
    _lib.call('cis_resize_bilinear_bf16', a.ptr, a.pitch, a.c_off, 2, H, W, out.data_ptr(), 24, 0, OH, OW, 3, ST())
    xr = x.clone().requires_grad_(True)
    ref = T.resize_bilinear_legacy(xr, OH, OW)
    torch.cuda.synchronize()
    assert float((out.float().cpu() - ref.detach()).abs().max()) <= 2 ** -7 * float(ref.abs().max()) + 1e-3
    gy = bf(torch.randn(2, OH, OW, 24, generator=g))
    ref.backward(gy)
    gd = gy.cuda().to(torch.bfloat16).contiguous()
    ds = torch.zeros(2, H, W, 24, dtype=torch.bfloat16, device='cuda')
    _lib.call('cis_resize_bilinear_bf16_bwd', gd.data_ptr(), 24, 0, 2, OH, OW, ds.data_ptr(), 24, 0, H, W, 3, 0, ST())
    _lib.call('cis_resize_bilinear_bf16_bwd', gd.data_ptr(), 24, 0, 2, OH, OW, ds.data_ptr(), 24, 0, H, W, 3, 1, ST())   # accumulate: 2x
    torch.cuda.synchronize()
    assert float((ds.float().cpu() - 2 * xr.grad).abs().max()) <= 2 ** -6 * float(xr.grad.abs().max()) * 2 + 1e-2


def test_upsample_nn2x_and_transpose():
    g = torch.Generator().manual_seed(3)
    x = bf(torch.randn(2, 7, 9, 16, generator=g))
    xd = x.cuda().to(torch.bfloat16).contiguous()
    out = torch.zeros(2, 14, 18, 16, dtype=torch.bfloat16, device='cuda')
    _lib.call('cis_upsample_nn2x', xd.data_ptr(), 2, 7, 9, 16, out.data_ptr(), ST())
    xr = x.clone().requires_grad_(True)
    ref = T.resize_nn_align_corners(xr, 14, 18)
    torch.cuda.synchronize()
    assert torch.equal(out.float().cpu(), ref.detach())
    gy = bf(torch.randn(2, 14, 18, 16, generator=g))
    ref.backward(gy)
    gd = gy.cuda().to(torch.bfloat16).contiguous()
    ds = torch.zeros(2, 7, 9, 16, dtype=torch.bfloat16, device='cuda')
    _lib.call('cis_upsample_nn2x_bwd', gd.data_ptr(), 2, 7, 9, 16, ds.data_ptr(), 0, ST())
    torch.cuda.synchronize()
    assert float((ds.float().cpu() - xr.grad).abs().max()) <= 2 ** -7 * float(xr.grad.abs().max()) + 1e-3


def test_resize_f32_and_nn():
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 12, 20, 2, generator=g)
    out = torch.zeros(2, 8, 14, 2, device='cuda')
    xd = x.cuda()
    _lib.call('cis_resize_bilinear_f32', xd.data_ptr(), 2, 12, 20, 2, out.data_ptr(), 8, 14, 0.0125, ST())
    torch.cuda.synchronize()
    assert float((out.cpu() - T.resize_bilinear_legacy(x, 8, 14) / 80.0).abs().max()) <= 1e-6
    o2 = torch.zeros(2, 5, 7, 2, device='cuda')
    _lib.call('cis_resize_nn_f32', xd.data_ptr(), 2, 12, 20, 2, o2.data_ptr(), 5, 7, ST())
    torch.cuda.synchronize()
    assert torch.equal(o2.cpu(), T.resize_nn_legacy(x, 5, 7))


def test_flow_normalise_and_generator_input():
    from oracle import losses as OL
    g = torch.Generator().manual_seed(6)
    B, H, W = 3, 10, 14
    image = torch.rand(B, H, W, 3, generator=g) - 0.5
    flow = torch.randn(B, H, W, 2, generator=g) * 0.3 + 5.0      # large mean: exercises the variance cancellation
    stats = torch.zeros(B, 4, dtype=torch.float64, device='cuda')
    out = torch.zeros(B, H, W, 8, dtype=torch.bfloat16, device='cuda')
    idv, fdv = image.cuda(), flow.cuda()
    _lib.call('cis_flow_stats', fdv.data_ptr(), B, H * W, stats.data_ptr(), ST())
    _lib.call('cis_pack_generator_input', idv.data_ptr(), fdv.data_ptr(), stats.data_ptr(), B, H * W, out.data_ptr(), ST())
    torch.cuda.synchronize()
    ref = torch.cat([image, OL.preprocess_flow_batch(flow)], 3)
    assert float((out[..., :5].float().cpu() - ref).abs().max()) <= 2 ** -7 * float(ref.abs().max())
    assert float(out[..., 5:].abs().max()) == 0


@pytest.mark.parametrize('size', [(7, 9, 14, 18), (8, 14, 16, 28), (6, 10, 6, 10), (4, 7, 7, 13)])
def test_fused_resize_concat_and_its_transpose(size):
    """cis_resize_concat_bf16(_bwd): three sources of one resolution (the middle one batch-broadcast over 3 replicas, as the a-encoder
    features of the three weight-shared recover_net calls) -> legacy bilinear -> one concatenated slice; the transpose slices the
    channels, applies R^T, folds the replicas and accumulates where asked.  Checked against the oracle's resize + autograd."""
    import ctypes as C
    from unsupervised_detection_b200._lib import CisSrc
    H, W, OH, OW = size
    g = torch.Generator().manual_seed(17)
    N, nm = 6, 2
    chans = (16, 8, 24)
    xs = [bf(torch.randn(N if i != 1 else nm, H, W, c, generator=g)) for i, c in enumerate(chans)]
    xd = [x.cuda().to(torch.bfloat16).contiguous() for x in xs]
    tot = sum(chans)
    out = torch.zeros(N, OH, OW, tot + 8, dtype=torch.bfloat16, device='cuda')        # destination slice at channel offset 8
    arr = (CisSrc * 3)(*[CisSrc(t.data_ptr(), c, 0, c // 8, nm if i == 1 else 0) for i, (t, c) in enumerate(zip(xd, chans))])
    _lib.call('cis_resize_concat_bf16', arr, 3, N, H, W, out.data_ptr(), tot + 8, 8, OH, OW, ST())
    torch.cuda.synchronize()
    xr = [x.clone().requires_grad_(True) for x in xs]
    parts = [T.resize_bilinear_legacy(xr[0], OH, OW), T.resize_bilinear_legacy(xr[1], OH, OW).repeat(N // nm, 1, 1, 1),
             T.resize_bilinear_legacy(xr[2], OH, OW)]
    ref = torch.cat(parts, 3)
    got = out[..., 8:].float().cpu()
    assert float(out[..., :8].abs().max()) == 0
    assert float((got - ref.detach()).abs().max()) <= 2 ** -7 * float(ref.abs().max()) + 1e-3
    # transpose: source 0 plain, source 1 folds the replicas AND accumulates onto an existing gradient, source 2 not wanted
    gy = bf(torch.randn(N, OH, OW, tot, generator=g))
    ref.backward(gy)
    gd = torch.zeros(N, OH, OW, tot + 8, dtype=torch.bfloat16, device='cuda')
    gd[..., 8:] = gy.cuda().to(torch.bfloat16)
    pre = bf(torch.randn(nm, H, W, chans[1], generator=g))
    grads = [torch.zeros(N, H, W, chans[0], dtype=torch.bfloat16, device='cuda'), pre.cuda().to(torch.bfloat16).contiguous(),
             torch.full((N, H, W, chans[2]), 7.0, dtype=torch.bfloat16, device='cuda')]
    ga = (CisSrc * 3)(*[CisSrc(t.data_ptr(), c, 0, c // 8, nm if i == 1 else 0) for i, (t, c) in enumerate(zip(grads, chans))])
    want, acc = (C.c_int32 * 3)(1, 1, 0), (C.c_int32 * 3)(0, 1, 0)
    _lib.call('cis_resize_concat_bf16_bwd', gd.data_ptr(), tot + 8, 8, N, OH, OW, ga, want, acc, 3, H, W, ST())
    torch.cuda.synchronize()
    tol = lambda r: 2 ** -6 * float(r.abs().max()) + 1e-2
    assert float((grads[0].float().cpu() - xr[0].grad).abs().max()) <= tol(xr[0].grad)
    assert float((grads[1].float().cpu() - (xr[1].grad + pre)).abs().max()) <= tol(xr[1].grad + pre)
    assert float((grads[2].float() - 7.0).abs().max()) == 0                      # untouched


@pytest.mark.parametrize('nch,act', [(32, 1), (2, 2), (128, 1), (20, 2)])
def test_dact_colsum_equals_the_two_separate_passes(nch, act):
    """cis_dact_colsum = cis_dact_mul followed by cis_colsum on the same slice: the gradient and every per-block partial bit-identical
    (same block count, same pixel order), with a residual source and a slice at a channel offset inside a wider buffer."""
    g = torch.Generator().manual_seed(5)
    npix, c8, nblocks = 3 * 37 * 29, (nch + 7) // 8 * 8, 17
    pitch = c8 + 16
    grad = torch.zeros(npix, pitch, dtype=torch.bfloat16, device='cuda')
    grad[:, 8:8 + c8] = torch.randn(npix, c8, generator=g).cuda().to(torch.bfloat16)
    y = torch.randn(npix, c8, generator=g).cuda().to(torch.bfloat16)
    res = (0.5 * torch.randn(npix, c8, generator=g)).cuda().to(torch.bfloat16)
    g0, g2 = grad.clone(), grad.clone()
    p1 = torch.full((nblocks, nch), -1.0, device='cuda')
    p2 = torch.full((nblocks, nch), -1.0, device='cuda')
    _lib.call('cis_dact_mul', grad.data_ptr(), pitch, 8, y.data_ptr(), c8, 0, res.data_ptr(), c8, 0, npix, c8 // 8, act, 0.1, ST())
    _lib.call('cis_colsum', grad.data_ptr(), pitch, 8, npix, nch, p1.data_ptr(), nblocks, ST())
    _lib.call('cis_dact_colsum', g2.data_ptr(), pitch, 8, y.data_ptr(), c8, 0, res.data_ptr(), c8, 0, npix, nch, act, 0.1, p2.data_ptr(), nblocks, ST())
    torch.cuda.synchronize()
    assert torch.equal(grad, g2) and torch.equal(p1, p2)
    u = y.float() - res.float()
    d = torch.where(u > 0, torch.ones_like(u), u + 1 if act == 1 else torch.full_like(u, 0.1))
    assert torch.equal(g2[:, 8:8 + c8], (g0[:, 8:8 + c8].float() * d).to(torch.bfloat16))
    assert float((p2.sum(0) - g2[:, 8:8 + nch].float().sum(0)).abs().max()) <= 1e-3 * npix ** 0.5
    assert float(g2[:, :8].abs().max()) == 0 and float(g2[:, 8 + c8:].abs().max()) == 0


def test_device_central_crops_match_the_host_reader_rule():
    """cis_crop_resize_bilinear_f32 (multi-crop ensemble inputs cut on the device) against data/crops.central_crops (host restatement of
    Davis2016Reader.central_cropping, itself pinned in tests/test_davis_reader.py): same boxes, same legacy-bilinear values."""
    from unsupervised_detection_b200.data.crops import central_crops
    from unsupervised_detection_b200.data.davis2016_data_utils import central_crop_box
    g = torch.Generator().manual_seed(23)
    hs, ws = 96, 160
    img = torch.rand(1, hs, ws, 3, generator=g)
    gt = (torch.rand(1, hs, ws, 1, generator=g) > 0.5).float()
    crops = [0.85, 0.9, 0.95, 1.0]
    r1, _, rg = central_crops(img, img, gt, crops)
    di, dg = img.cuda(), gt.cuda()
    o1 = torch.empty(len(crops), hs, ws, 3, device='cuda')
    og = torch.empty(len(crops), hs, ws, 1, device='cuda')
    for i, c in enumerate(crops):
        y0, x0, ch, cw = central_crop_box(hs, ws, c)
        _lib.call('cis_crop_resize_bilinear_f32', di.data_ptr(), hs, ws, 3, y0, x0, ch, cw, o1[i].data_ptr(), hs, ws, ST())
        _lib.call('cis_crop_resize_bilinear_f32', dg.data_ptr(), hs, ws, 1, y0, x0, ch, cw, og[i].data_ptr(), hs, ws, ST())
    torch.cuda.synchronize()
    assert float((o1.cpu() - r1).abs().max()) <= 2e-6
    assert float((og.cpu() - rg).abs().max()) <= 2e-6
    assert torch.equal(o1[3].cpu(), img[0]) and torch.equal(og[3].cpu(), gt[0])          # crop 1.0 is the identity
    with pytest.raises(Exception):
        _lib.call('cis_crop_resize_bilinear_f32', di.data_ptr(), hs, ws, 3, 10, 0, hs, ws, o1[0].data_ptr(), hs, ws, ST())
