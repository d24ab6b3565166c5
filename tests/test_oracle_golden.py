"""The fp32 oracle against the committed fp64 golden fixtures (tests/golden/make_golden.py)."""
import os
import numpy as np
import torch

from oracle import params as OP, pwcnet as PW, losses as OL

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
t = lambda a: torch.from_numpy(np.asarray(a)).float()


def test_losses_fixture():
    z = np.load(os.path.join(G, 'cis_losses_32x48.npz'))
    p = OP.make_params(seed=int(z['seed']), jitter=float(z['jitter']))
    L = OL.adversarial_losses(t(z['image']), t(z['flow']), p)
    assert float((L['masks'] - t(z['mask'])).abs().max()) < 1e-5
    for k in ('pred', 'pred_c', 'pred_i'):
        assert float((L[k] - t(z[k])).abs().max()) < 1e-4
    assert abs(float(L['recover']) - float(z['recover'])) < 1e-5
    assert abs(float(L['generator']) - float(z['generator'])) < 1e-5


def test_pwc_fixture():
    z = np.load(os.path.join(G, 'pwc_64x64.npz'))
    p = OP.make_params(seed=int(z['seed']), jitter=float(z['jitter']))
    fl, pyr, c1, _ = PW.predict_from_img_pairs(t(z['img1']), t(z['img2']), p, return_pyr=True)
    assert float((fl - t(z['flow'])).abs().max()) < 2e-4
    assert float((pyr[0] - t(z['flow6'])).abs().max()) < 1e-4
    assert float((c1[3] - t(z['c1_3'])).abs().max()) < 1e-4


def test_warp_costvol_fixture():
    z = np.load(os.path.join(G, 'warp_costvol_6x7.npz'))
    wr = PW.dense_image_warp(t(z['c2']), t(z['flow']))
    assert float((wr - t(z['warp'])).abs().max()) < 1e-5
    assert float((PW.cost_volume(t(z['c1']), wr) - t(z['cv'])).abs().max()) < 1e-5
