"""Function-level API (models/functional.py, SURVEY 8b: generator_net, recover_net, charbonnier_loss, train_op, cost_volume,
dense_image_warp, ModelPWCNet.predict_from_img_pairs under the reference's module paths) against the oracle on the GPU."""
import os

import pytest
import torch

from oracle import nets as ON, pwcnet as OW, losses as OL, params as OP
from unsupervised_detection_b200.models.nets import generator_net, recover_net
from unsupervised_detection_b200.models.utils.loss_utils import charbonnier_loss, train_op
from unsupervised_detection_b200.models.PWCNet.core_costvol import cost_volume
from unsupervised_detection_b200.models.PWCNet.core_warp import dense_image_warp
from unsupervised_detection_b200.models.PWCNet.model_pwcnet import ModelPWCNet

pytestmark = pytest.mark.gpu
bf = lambda x: x.to(torch.bfloat16).float()


@pytest.fixture(scope='module')
def params():
    return OP.make_params(11)


def test_generator_net_matches_oracle(params):
    g = torch.Generator().manual_seed(0)
    img = torch.rand(2, 64, 96, 3, generator=g) - 0.5
    flw = torch.randn(2, 64, 96, 2, generator=g)
    got = generator_net(img.cuda(), flw.cuda(), 'MaskNet/', params=params).cpu()
    ref = ON.generator_net(img, flw, params)
    assert got.shape == ref.shape == (2, 64, 96, 1)
    assert float((got - ref).abs().mean()) < 2e-3 and float((got - ref).abs().max()) < 3e-2     # bf16 activations through 17 layers


def test_recover_net_matches_oracle(params):
    g = torch.Generator().manual_seed(1)
    img = torch.rand(1, 64, 96, 3, generator=g) - 0.5
    m = torch.rand(1, 64, 96, 1, generator=g)
    fm = torch.randn(1, 64, 96, 2, generator=g) * 0.3 * (1 - m)
    got = recover_net(img.cuda(), fm.cuda(), m.cuda(), 'FlownetS/', params=params).cpu()
    ref = ON.recover_net(img, fm, m, params)
    assert got.shape == ref.shape == (1, 64, 96, 2)
    assert float((got - ref).abs().max()) <= 2 ** -5 * float(ref.abs().max()) + 2e-2


def test_predict_from_img_pairs_matches_oracle(params):
    g = torch.Generator().manual_seed(2)
    a = torch.rand(1, 128, 128, 3, generator=g) - 0.5      # level 6 must be at least 2x2 (dense_image_warp, core_warp.py:188)
    b = torch.roll(a, shifts=(1, 2), dims=(1, 2))
    got = ModelPWCNet.predict_from_img_pairs(a.cuda(), b.cuda(), params=params).cpu()
    ref = OW.predict_from_img_pairs(a, b, params)
    assert got.shape == ref.shape == (1, 128, 128, 2)
    assert float((got - ref).abs().mean()) <= 0.02 * float(ref.abs().mean()) + 0.05


def test_charbonnier_loss_matches_oracle():
    g = torch.Generator().manual_seed(3)
    gt, pr = torch.randn(3, 20, 28, 2, generator=g), torch.randn(3, 20, 28, 2, generator=g)
    for mask in (torch.rand(3, 20, 28, 1, generator=g), torch.ones(3, 20, 28, 2)):
        for cbn in (0.5, 1.0, 0.3):
            got = charbonnier_loss(gt.cuda(), pr.cuda(), mask.cuda(), cbn).cpu()
            ref = OL.charbonnier_loss(gt, pr, mask, cbn)
            assert torch.allclose(got, ref, rtol=2e-5, atol=1e-3), (cbn, got, ref)


def test_cost_volume_and_warp_match_oracle():
    g = torch.Generator().manual_seed(4)
    c1, c2 = bf(torch.randn(2, 12, 20, 37, generator=g)), bf(torch.randn(2, 12, 20, 37, generator=g))
    fl = torch.randn(2, 12, 20, 2, generator=g) * 3
    w = dense_image_warp(c2.cuda(), fl.cuda()).cpu()
    wr = OW.dense_image_warp(c2, fl)
    assert float((w - wr).abs().max()) <= 2 ** -7 * float(wr.abs().max()) + 1e-3
    cv = cost_volume(c1.cuda(), c2.cuda()).cpu()
    cr = OW.cost_volume(c1, c2)
    assert cv.shape == (2, 12, 20, 81) and float((cv - cr).abs().max()) <= 2 ** -8 * float(cr.abs().max()) + 2e-3


def test_train_op_matches_oracle_adam():
    g = torch.Generator().manual_seed(5)
    n = 1000
    p0, gr = torch.randn(n, generator=g), torch.randn(n, generator=g) * 0.5
    p, m, v = p0.clone().cuda(), torch.zeros(n).cuda(), torch.zeros(n).cuda()
    st = torch.zeros(1, dtype=torch.int64, device='cuda')
    opt = OL.TFAdam(1e-4, 0.9, 0.999, 1e-8)
    ref = {'w': p0.clone()}
    for _ in range(3):
        train_op(p, gr.cuda(), m, v, st, gradient_clip_value=0.2, can_change=False)
        opt.apply(ref, ['w'], [gr.clamp(-0.2, 0.2)])
    torch.cuda.synchronize()
    assert float((p.cpu() - ref['w']).abs().max()) < 1e-6
