"""Known-answer tests pinning the TF-1.13 op semantics of SURVEY.md Appendix A in the oracle (hand-computed values).
The reference ships no tests or fixtures (tests/.gitkeep), so these KATs ARE the pin of the oracle ("parity unpinned"
at the reference level, see oracle/__init__.py)."""
import math
import numpy as np
import torch

from oracle import tf_ops as T, pwcnet as PW, losses as OL, metrics as OM, params as OP


def test_same_padding_is_asymmetric():
    # App. A.2: extra padding goes bottom/right
    assert T.same_pad(8, 3, 2) == (0, 1)
    assert T.same_pad(8, 5, 2) == (1, 2)
    assert T.same_pad(8, 7, 2) == (2, 3)
    assert T.same_pad(8, 4, 1) == (1, 2)
    assert T.same_pad(8, 3, 1) == (1, 1)
    assert T.same_pad(8, 3, 1, 4) == (4, 4)
    assert T.same_pad(7, 3, 2) == (1, 1)      # odd input: out=4, total=2


def test_conv_same_stride2_picks_right_taps():
    x = torch.arange(16.0).reshape(1, 4, 4, 1)
    w = torch.zeros(3, 3, 1, 1)
    w[0, 0] = 1.0                             # top-left tap only; pad (0,1) => out[i,j] = x[2i, 2j]
    y = T.conv2d_same(x, w, 2)
    assert y.reshape(2, 2).tolist() == [[0.0, 2.0], [8.0, 10.0]]
    w = torch.zeros(4, 4, 1, 1)
    w[0, 0] = 1.0                             # k4 s1 pad (1,2): out[i,j] = x[i-1, j-1]
    y = T.conv2d_same(x, w, 1).reshape(4, 4)
    assert y[0].tolist() == [0, 0, 0, 0] and y[1].tolist() == [0, 0, 1, 2]


def test_legacy_bilinear_x2():
    x = torch.tensor([0.0, 10.0, 20.0]).reshape(1, 1, 3, 1)
    y = T.resize_bilinear_legacy(x, 1, 6).reshape(-1)
    # even dst = exact pixel, odd = midpoint, last replicated (App. A.6)
    assert y.tolist() == [0.0, 5.0, 10.0, 15.0, 20.0, 20.0]
    assert T.resize_bilinear_legacy(x, 1, 3) is x


def test_legacy_bilinear_downsample_is_two_tap():
    x = torch.arange(6.0).reshape(1, 1, 6, 1)
    y = T.resize_bilinear_legacy(x, 1, 4).reshape(-1)     # scale 1.5: src 0,1.5,3,4.5
    assert torch.allclose(y, torch.tensor([0.0, 1.5, 3.0, 4.5]))


def test_nn_align_corners_drifts():
    x = torch.arange(64.0).reshape(1, 64, 1, 1).expand(1, 64, 2, 1)
    y = T.resize_nn_align_corners(x, 128, 2)[0, :, 0, 0]
    # App. A.5: src = round(d * 63/127): 0,1->0 ; 2,3->1 ; ... ; 126,127->63 ; but NOT plain duplication in the middle
    assert y[0] == 0 and y[1] == 0 and y[2] == 1 and y[3] == 1
    assert y[126] == 63 and y[127] == 63
    assert y[64] == round(64 * 63 / 127) == 32 and y[65] == 32 and y[63] == 31
    ref = [min(int(math.floor(d * (63.0 / 127.0) + 0.5)), 63) for d in range(128)]
    assert y.tolist() == [float(v) for v in ref]


def test_nn_legacy_floor():
    x = torch.arange(4.0).reshape(1, 4, 1, 1)
    assert T.resize_nn_legacy(x, 8, 1).reshape(-1).tolist() == [0, 0, 1, 1, 2, 2, 3, 3]
    assert T.resize_nn_legacy(x, 3, 1).reshape(-1).tolist() == [0, 1, 2]


def test_bn_inference_identity_stats():
    x = torch.tensor([[[[2.0]]]])
    y = T.batch_norm_inference(x, torch.tensor([3.0]), torch.tensor([0.5]))
    assert abs(float(y) - (3.0 * 2.0 / math.sqrt(1.001) + 0.5)) < 1e-6


def test_dense_image_warp_semantics():
    img = torch.arange(12.0).reshape(1, 3, 4, 1)
    flow = torch.zeros(1, 3, 4, 2)
    assert torch.equal(PW.dense_image_warp(img, flow), img)
    flow[..., 1] = 1.0                         # channel 1 displaces columns: out[j,i] = img[j, i-1], edge replicated
    out = PW.dense_image_warp(img, flow)[0, :, :, 0]
    assert out[0].tolist() == [0.0, 0.0, 1.0, 2.0]
    flow.zero_()
    flow[..., 0] = -0.5                        # channel 0 displaces rows: query y = j + 0.5
    out = PW.dense_image_warp(img, flow)[0, :, :, 0]
    assert out[0].tolist() == [2.0, 3.0, 4.0, 5.0]
    assert out[2].tolist() == [8.0, 9.0, 10.0, 11.0]      # clamped: floor<=size-2, alpha<=1 => last row replicated
    flow[..., 0] = 100.0                       # far out of range: row 0 everywhere
    out = PW.dense_image_warp(img, flow)[0, :, :, 0]
    assert out[2].tolist() == [0.0, 1.0, 2.0, 3.0]


def test_cost_volume_channel_order_and_mean():
    c1 = torch.zeros(1, 9, 9, 4)
    c2 = torch.zeros(1, 9, 9, 4)
    c1[0, 4, 4] = torch.tensor([1.0, 2.0, 3.0, 4.0])
    c2[0, 6, 3] = torch.tensor([1.0, 1.0, 1.0, 1.0])     # displacement (+2 rows, -1 col)
    cv = PW.cost_volume(c1, c2)
    ch = 9 * (2 + 4) + (-1 + 4)
    assert abs(float(cv[0, 4, 4, ch]) - 10.0 / 4.0) < 1e-6  # mean over C, not sum
    assert float(cv.abs().sum()) == float(cv[0, 4, 4, ch])
    c2[0, 6, 3] = -1.0
    assert abs(float(PW.cost_volume(c1, c2)[0, 4, 4, ch]) + 0.1 * 2.5) < 1e-6  # leaky 0.1


def test_conv_transpose_matches_definition():
    x = torch.zeros(1, 2, 2, 1)
    x[0, 0, 0, 0] = 1.0
    w = torch.arange(16.0).reshape(4, 4, 1, 1)
    y = T.conv2d_transpose_k4s2(x, w)[0, :, :, 0]          # out[2i-1+ky, 2j-1+kx] += x[i,j] w[ky,kx]
    assert y.shape == (4, 4)
    assert y[0, 0] == w[1, 1, 0, 0] and y[2, 2] == w[3, 3, 0, 0] and y[0, 1] == w[1, 2, 0, 0]


def test_preprocess_flow_population_variance():
    f = torch.tensor([1.0, 3.0]).reshape(1, 1, 2, 1).repeat(1, 1, 1, 2)
    y = OL.preprocess_flow_batch(f)
    assert torch.allclose(y[0, 0, :, 0], torch.tensor([-1.0, 1.0]))


def test_charbonnier():
    gt = torch.tensor([3.0, 0.0]).reshape(1, 1, 1, 2)
    pr = torch.zeros(1, 1, 1, 2)
    m = torch.full((1, 1, 1, 1), 0.5)
    v = float(OL.charbonnier_loss(gt, pr, m)[0])
    assert abs(v - 0.5 * (math.sqrt(9 + 1e-6) + 1e-3)) < 1e-6


def test_tf_adam_first_steps_and_shared_step():
    p = {'a': torch.tensor([1.0]), 'b': torch.tensor([1.0])}
    opt = OL.TFAdam(lr=1e-4, beta1=0.9)
    opt.apply(p, ['a'], [torch.tensor([0.2])])
    lr1 = 1e-4 * math.sqrt(1 - 0.999) / (1 - 0.9)
    exp = 1.0 - lr1 * (0.1 * 0.2) / (math.sqrt(0.001 * 0.04) + 1e-8)
    assert abs(float(p["a"]) - exp) < 1e-7
    opt.apply(p, ['b'], [torch.tensor([0.2])])             # t = 2 although 'b' is updated for the first time (shared powers)
    lr2 = 1e-4 * math.sqrt(1 - 0.999 ** 2) / (1 - 0.9 ** 2)
    exp_b = 1.0 - lr2 * (0.1 * 0.2) / (math.sqrt(0.001 * 0.04) + 1e-8)
    assert abs(float(p["b"]) - exp_b) < 1e-7


def test_clip_and_noise_branch():
    g = [torch.tensor([0.5, -0.5, 0.1])]
    c, n = OL.clip_or_noise(g, 0.2, can_change=True)
    assert not n and torch.allclose(c[0], torch.tensor([0.2, -0.2, 0.1]))
    c, n = OL.clip_or_noise([torch.full((1000,), 1e-7)], 0.2, can_change=True, gen=torch.Generator().manual_seed(0))
    assert n and float(c[0].min()) >= 0.0 and float(c[0].max()) <= 0.2 and abs(float(c[0].mean()) - 0.1) < 0.02
    c, n = OL.clip_or_noise([torch.full((10,), 1e-7)], 0.2, can_change=False)
    assert not n


def test_step_schedule():
    kinds = ['R' if OL.is_recover_step(s) else 'G' for s in range(1, 9)]
    assert kinds == ['G', 'G', 'G', 'R', 'G', 'G', 'G', 'R']   # adversarial_learner.py:386-389 with step starting at 1


def test_iou_and_border_disambiguation():
    pm = torch.zeros(1, 10, 10, 1)
    pm[0, 3:7, 3:7] = 1.0
    gt = torch.zeros(1, 10, 10, 1)
    gt[0, 3:7, 3:5] = 1.0
    assert abs(float(OM.compute_all_IoU(pm, gt)[0]) - 8.0 / 16.0) < 1e-6
    inv = 1.0 - pm                              # border-hugging mask => complement is used
    assert abs(float(OM.compute_all_IoU(inv, gt)[0]) - 0.5) < 1e-6
    assert abs(float(OM.compute_boundary_score_tf(torch.ones(1, 10, 10, 1))[0]) - 1.0) < 1e-6
    iou, ann = OM.compute_IoU(np.zeros((4, 4), bool), np.zeros((4, 4), np.float32))
    assert iou == 1.0 and ann.sum() == 0          # empty/empty: reference returns 1 (arity fixed)
    assert abs(OM.compute_mae(np.ones((2, 2)), np.zeros((2, 2))) - 1.0) < 1e-6


def test_param_counts_match_reference():
    p = OP.make_params()
    assert OP.count(p, 'MaskNet/') == 1451062 and OP.count(p, 'FlownetS/') == 3388610 and OP.count(p, 'pwcnet/') == 14079050
