"""Plumbing of the function-level API (models/functional.py) without a GPU: sub-graph plans are BUILT on CPU tensors and
inspected, kernel calls of the thin wrappers are recorded by a stub.  The numerical GPU checks live in test_functional_api_gpu.py."""
import collections

import pytest
import torch

from unsupervised_detection_b200.models import functional as F
from unsupervised_detection_b200.models import nets, utils  # noqa: F401
from unsupervised_detection_b200 import params_init


def _names(plan):
    return [op[2] for op in plan.ops if op[0] is not None]


def test_cpu_tensors_are_rejected_loudly():
    x = torch.zeros(1, 8, 8, 3)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        F.generator_net(x, torch.zeros(1, 8, 8, 2))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        F.charbonnier_loss(torch.zeros(1, 4, 4, 2), torch.zeros(1, 4, 4, 2), torch.ones(1, 4, 4, 1))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        F.dense_image_warp(x, torch.zeros(1, 8, 8, 2))


def test_generator_subgraph_plan():
    r = F._GeneratorRunner(2, 32, 48, 'cpu', 'MaskNet')
    n = _names(r.bld.fwd)
    assert n[0] == 'cis_pack_generator_input' and n.count('cis_conv_igemm') == 17 and n.count('cis_upsample_nn2x') == 2
    assert r.stats.tolist() == [[0.0, 0.0, 32.0 * 48, 32.0 * 48]] * 2          # identity normalisation: mean 0, variance 1
    assert r.mask.shape == (2, 32, 48, 1) and r.gen_in.pitch == 8
    assert [e[0] for e in r.store.entries][:4] == ['MaskNet/conv1/kernel', 'MaskNet/conv1/bias', 'MaskNet/conv1/gamma', 'MaskNet/conv1/beta']
    assert _names(r.pack).count('cis_bn_fold') == 17
    r.store.load(params_init.init_generator())                                # the registry / checkpoint names fit the sub-graph's store


def test_recover_subgraph_plan():
    r = F._RecoverRunner(1, 64, 96, 'cpu', 'FlownetS', 0.25)
    n = _names(r.bld.fwd)
    assert n[:2] == ['cis_pack_f32_to_bf16', 'cis_pack_f32_to_bf16'] and n[-1] == 'cis_resize_bilinear_f32'
    assert n.count('cis_conv_igemm') == 32                                     # 9 + 9 encoder convs, 14 decoder convs: one call, no batching x3
    packs = [op for op in r.bld.fwd.ops if op[2] == 'cis_pack_f32_to_bf16']
    assert packs[0][1][2] == 3 and packs[1][1][2] == 4                         # image: 3 channels; [flow_masked, ones, 1-mask]: 4
    assert r.flow1.shape == (1, 32, 48, 2) and r.pred.shape == (1, 64, 96, 2)
    last = r.bld.fwd.ops[-1][1]
    assert last[1:5] == (1, 32, 48, 2) and last[6:9] == (64, 96, 1.0)          # legacy-bilinear x2 back to the input size, scale 1
    r.store.load(params_init.init_recover())


def test_pwc_subgraph_plan():
    r = F._PWCRunner(1, 64, 128, 'cpu', 'pwcnet')
    n = collections.Counter(_names(r.bld.fwd))
    assert n['cis_warp_costvol'] == 5 and n['cis_pack_f32_to_bf16'] == 2 and n['cis_conv_igemm'] > 100
    assert r.flow.shape == (1, 64, 128, 2)
    with pytest.raises(ValueError, match='multiples of 64'):
        F.predict_from_img_pairs(torch.zeros(1, 100, 128, 3, device='meta').to('cpu') if False else _FakeCuda(1, 100, 128, 3), _FakeCuda(1, 100, 128, 3))


class _FakeCuda(object):
    """Just enough of a tensor for the argument checks that run before any kernel."""
    is_cuda = True
    device = 'cuda:0'

    def __init__(self, *shape):
        self.shape = shape


def test_thin_wrappers_issue_the_expected_kernel_calls(monkeypatch):
    calls = []
    monkeypatch.setattr(F._lib, 'call', lambda name, *a: calls.append((name, a)))
    monkeypatch.setattr(F, '_check_cuda', lambda *t: None)
    monkeypatch.setattr(F, '_stream', lambda: 0)
    gt, pr = torch.randn(2, 6, 5, 2), torch.randn(2, 6, 5, 2)
    out = F.charbonnier_loss(gt, pr, torch.ones(2, 6, 5, 1), cbn=0.5)
    assert out.shape == (2,) and calls[-1][0] == 'cis_charbonnier_sum' and calls[-1][1][3:8] == (2, 30, 2, 1, 0.5)
    F.charbonnier_loss(gt, pr, torch.ones(2, 6, 5, 2), cbn=1.0)
    assert calls[-1][1][3:8] == (2, 30, 2, 2, 1.0)
    with pytest.raises(ValueError):
        F.charbonnier_loss(gt, pr, torch.ones(2, 6, 5, 3))
    c1, wp = torch.randn(1, 6, 10, 196), torch.randn(1, 6, 10, 196)
    cv = F.cost_volume(c1, wp)
    name, a = calls[-1]
    assert name == 'cis_warp_costvol' and cv.shape == (1, 6, 10, 81)
    assert a[1] == 200 and a[4] == 200 and a[6] is None and a[8:12] == (1, 6, 10, 196) and a[13:15] == (88, 0)   # pitch padded to 8, no flow
    with pytest.raises(NotImplementedError):
        F.cost_volume(c1, wp, search_range=3)
    w = F.dense_image_warp(torch.randn(2, 4, 4, 5), torch.zeros(2, 4, 4, 2))
    name, a = calls[-1]
    assert name == 'cis_dense_image_warp' and w.shape == (2, 4, 4, 5) and a[1] == 8 and a[4:9] == (1.0, 2, 4, 4, 5)
    p, g, m, v = (torch.zeros(100) for _ in range(4))
    st = torch.zeros(1, dtype=torch.int64)
    F.train_op(p, g, m, v, st, gradient_clip_value=0.2, can_change=False)
    assert calls[-1][0] == 'cis_clip_adam' and calls[-1][1][4:7] == (100, 1.0, 0.2) and calls[-1][1][13] == 0
    seg = torch.tensor([[0, 60], [60, 100]])
    F.train_op(p, g, m, v, st, gradient_clip_value=0.2, can_change=True, segments=seg)
    assert [c[0] for c in calls[-2:]] == ['cis_grad_avg_abs', 'cis_clip_adam'] and calls[-1][1][13] == 1 and calls[-2][1][2] == 2
    with pytest.raises(ValueError):
        F.train_op(p, g, m, v, st, can_change=True)


def test_reference_module_paths_expose_the_functions():
    from unsupervised_detection_b200.models.nets import generator_net, recover_net
    from unsupervised_detection_b200.models.utils.loss_utils import charbonnier_loss, train_op
    from unsupervised_detection_b200.models.PWCNet.core_costvol import cost_volume
    from unsupervised_detection_b200.models.PWCNet.core_warp import dense_image_warp
    from unsupervised_detection_b200.models.PWCNet.model_pwcnet import ModelPWCNet
    assert callable(generator_net) and callable(recover_net) and charbonnier_loss is F.charbonnier_loss
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        generator_net(torch.zeros(1, 8, 8, 3), torch.zeros(1, 8, 8, 2), 'MaskNet/')      # delegates to models/functional.py
    assert train_op is F.train_op and cost_volume is F.cost_volume and dense_image_warp is F.dense_image_warp
    assert callable(ModelPWCNet.predict_from_img_pairs)
