"""The reference-facing surface on the GPU: AdversarialLearner.step()/inference() through host buffers, and
size-independent properties at BASELINE's full size (256x448, batch 4)."""
import os
import numpy as np
import pytest
import torch

from unsupervised_detection_b200.common_flags import Config
from unsupervised_detection_b200.models.adversarial_learner import AdversarialLearner

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def learner():
    L = AdversarialLearner()
    L.config = Config(img_height=256, img_width=448, batch_size=4, dataset='SYNTHETIC', flow_ckpt='synthetic', summary_freq=2)
    L.build_train_graph()
    return L


def test_flow_ckpt_is_mandatory():
    L = AdversarialLearner()
    L.config = Config(img_height=64, img_width=64, batch_size=1, dataset='SYNTHETIC', flow_ckpt='')
    with pytest.raises(IOError):
        L.build_train_graph()                       # adversarial_learner.py:343


def test_step_api_full_size(learner):
    L = learner
    assert L.graph.param_count() == 18918722       # "Number of params" (adversarial_learner.py:338)
    batch = L.reader.batch(4)
    outs = [L.step(batch) for _ in range(4)]
    assert [o['train_op'] for o in outs] == ['G', 'G', 'G', 'R']
    assert outs[-1]['global_step'] == 1
    assert 'loss_recover' in outs[1] and np.isfinite(outs[1]['loss_recover']) and np.isfinite(outs[1]['loss_generator'])
    m = L.graph.mask
    assert float(m.min()) > 0.0 and float(m.max()) < 1.0 and torch.isfinite(m).all()
    assert all(torch.isfinite(v).all() for v in L.graph.export_params().values())


def test_forward_is_deterministic_and_batch_independent(learner):
    """Samples are independent in the forward (per-sample flow normalisation, BN without batch statistics): swapping two
    samples of the batch swaps their masks bit-exactly; repeated runs are bit-identical."""
    L = learner
    g = L.graph
    b = L.reader.batch(4)
    L.feed(b[0], b[1])
    g.forward()
    m1 = g.mask.clone()
    g.forward()
    assert torch.equal(m1, g.mask)
    perm = [1, 0, 2, 3]
    L.feed(b[0][perm], b[1][perm])
    g.forward()
    assert torch.equal(m1[perm], g.mask)


def test_data_parallel_split_equals_full_batch(learner):
    """local-mean gradients of the batch halves summed == global-batch gradient (emulated on one GPU by running
    the two halves through a batch-2 graph with global_batch=4)."""
    from unsupervised_detection_b200.step_graph import CISGraph
    L = learner
    g4 = L.graph
    b = L.reader.batch(4)
    L.feed(b[0], b[1])
    g4.forward()
    g4.bwd['R'].run()
    full = g4.rec_store.grad.clone()
    g2 = CISGraph(256, 448, 2, global_batch=4)
    g2.load_params({k: v for k, v in g4.export_params().items()})
    acc = torch.zeros_like(g2.rec_store.grad)
    for h in range(2):
        g2.img1.copy_(b[0][2 * h:2 * h + 2])
        g2.img2.copy_(b[1][2 * h:2 * h + 2])
        g2.forward()
        g2.bwd['R'].run()
        acc += g2.rec_store.grad
    torch.cuda.synchronize()
    cos = float(torch.dot(acc, full) / (acc.norm() * full.norm()))
    assert cos > 0.9999 and abs(float(acc.norm() / full.norm()) - 1) < 1e-3


def test_inference_keys(learner):
    L = AdversarialLearner()
    L.setup_inference(Config(img_height=128, img_width=224, batch_size=1, dataset='SYNTHETIC'), aug_test=False)
    L.restore('synthetic')
    r = L.inference(None)
    assert set(r) == {'gen_masks', 'pred_flow', 'input_image', 'gt_flow', 'gt_masks', 'img_fname'}    # adversarial_learner.py:617-619
    assert r['gen_masks'].shape == (1, 128, 224, 1) and r['pred_flow'].shape == (1, 128, 224, 2)
    from unsupervised_detection_b200.models.utils.general_utils import compute_IoU
    iou, ann = compute_IoU(r['gt_masks'][0], r['gen_masks'][0])
    assert 0.0 <= iou <= 1.0


def test_cli_eval_scripts_write_the_reference_dumps(tmp_path):
    """test_generator.py / test_generator_ensemble.py end to end on synthetic frames: per-category report, PNG overlays and .mat
    files with the keys the offline post-processing reads (generate_soft_score_from_buffer.py:45-92)."""
    import scipy.io as sio
    import test_generator as TG
    import test_generator_ensemble as TE
    from unsupervised_detection_b200.common_flags import FLAGS
    out1, out2 = str(tmp_path / 'single'), str(tmp_path / 'ens')
    FLAGS(['prog', '--dataset=SYNTHETIC', '--ckpt_file=synthetic', '--img_height=128', '--img_width=224', '--batch_size=2',
           '--generate_visualization', '--test_save_dir=' + out1])
    TG._test_masks()
    cats = os.listdir(out1)
    assert cats and any(f.endswith('.png') for f in os.listdir(os.path.join(out1, cats[0])))
    mat = [f for f in os.listdir(os.path.join(out1, cats[0])) if f.endswith('.mat')][0]
    m = sio.loadmat(os.path.join(out1, cats[0], mat))
    assert {'flow', 'img1', 'pred_mask', 'gt_mask'} <= set(m) and m['flow'].shape == (128, 224, 2)
    FLAGS(['prog', '--dataset=SYNTHETIC', '--ckpt_file=synthetic', '--img_height=128', '--img_width=224', '--batch_size=1',
           '--generate_visualization', '--test_save_dir=' + out2])
    TE._test_masks()
    cats = os.listdir(out2)
    mat = [f for f in os.listdir(os.path.join(out2, cats[0])) if f.endswith('.mat')][0]
    m = sio.loadmat(os.path.join(out2, cats[0], mat))
    for c in (85, 90, 95, 100):
        assert 'img_1_%03d' % c in m and 'pred_mask_%03d' % c in m and 'gt_mask_%03d' % c in m
    assert m['pred_mask_100'].shape[:2] == (128, 224)


def test_checkpoint_save_resume_and_restore_roundtrip(learner, tmp_path):
    """adversarial_learner.py:300-310,345-360 / test_generator.py:45-58 through the tf.train.Saver V2 bundle files: save ->
    (flow_ckpt | recover_ckpt | resume_train | ckpt_file) give back bit-identical fp32 parameters and the global step."""
    from unsupervised_detection_b200 import checkpoint as ck
    L = learner
    d = str(tmp_path / 'exp')
    L.save(None, d, 7)
    ref = {k: v.cpu() for k, v in L.graph.export_params().items()}
    prefix = os.path.join(d, 'model-7')
    assert ck.is_bundle(prefix) and ck.latest_checkpoint(d) == prefix
    # (1) resume_train picks the latest bundle in checkpoint_dir; PWC-Net comes from flow_ckpt given in the .data-* spelling
    os.remove(prefix + '.pt')
    R = AdversarialLearner()
    R.config = Config(img_height=64, img_width=96, batch_size=1, dataset='SYNTHETIC', flow_ckpt=prefix + '.data-00000-of-00001',
                      resume_train=True, checkpoint_dir=d, full_model_ckpt='')
    R.build_train_graph()
    got = R.graph.export_params()
    assert R.global_step == L.global_step
    assert set(got) == set(ref) and all(torch.equal(got[k].cpu(), ref[k]) for k in ref)
    # (2) recover_ckpt restores only FlownetS
    R2 = AdversarialLearner()
    R2.config = Config(img_height=64, img_width=96, batch_size=1, dataset='SYNTHETIC', flow_ckpt=prefix, recover_ckpt=prefix)
    R2.build_train_graph()
    got = R2.graph.export_params()
    assert all(torch.equal(got[k].cpu(), ref[k]) for k in ref if not k.startswith('MaskNet/'))
    # (3) inference restore from the same bundle reproduces the training graph's mask on the same frames
    T = AdversarialLearner()
    T.setup_inference(Config(img_height=256, img_width=448, batch_size=4, dataset='SYNTHETIC'), aug_test=False)
    T.restore(prefix + '.index')
    b = L.reader.batch(4)
    L.feed(b[0], b[1])
    L.graph.forward()
    r = T.inference(None, batch=b)
    torch.cuda.synchronize()
    # same parameters, same kernels: the masks agree (bf16 round-off headroom only; wrong weights would differ by O(0.1))
    assert np.abs(r['gen_masks'] - L.graph.mask.cpu().numpy()).max() < 2e-3


def test_step_summaries_written_like_collect_summaries(learner, tmp_path):
    """adversarial_learner.py:260-298,391-403: on a summary step the event file gets the 8 loss scalars, 6 images and one clipped-
    gradient histogram per recover and generator variable (both nets, although only one train op runs)."""
    from unsupervised_detection_b200.summary import read_events
    L = learner
    L.config.checkpoint_dir = str(tmp_path / 'logs')
    w = L.collect_summaries()
    before = {k: v.clone() for k, v in L.graph.export_params().items()}
    batch = L.reader.batch(4)
    res = None
    for _ in range(2):                              # summary_freq = 2: exactly one of the two steps is a summary step
        res = L.step(batch, summarize=True)
    w.close()
    L.summary_writer = None
    L.config.checkpoint_dir = ''
    ev = read_events(w.path)
    assert len(ev) == 2 and ev[1]['step'] == res['global_step']
    vals = ev[1]['values']
    scal = {v['tag']: v['simple_value'] for v in vals if 'simple_value' in v}
    assert set(scal) == {'generator', 'recover', 'red_rate', 'red_rate_compl', 'reconstruction_loss', 'reconstruction_compl_loss',
                         'denominator_red_rate', 'denominator_red_rate_compl'}
    assert all(np.isfinite(x) for x in scal.values()) and scal['denominator_red_rate'] > L.config.epsilon
    assert abs(scal['generator'] - (scal['red_rate'] + scal['red_rate_compl'])) < 1e-4
    imgs = [v for v in vals if 'image' in v]
    assert [v['tag'] for v in imgs] == ['input_image/image', 'next_image/image', 'masked_flow/image', 'PWC_Flow/image', 'Rec_flow/image',
                                        'Rec_flow_compl/image']
    assert (imgs[0]['image']['height'], imgs[0]['image']['width']) == (256, 448)
    assert (imgs[1]['image']['height'], imgs[1]['image']['width']) == (384, 640)
    hist = [v for v in vals if 'histo' in v]
    nR, nG = len(L.graph.rec_store.entries), len(L.graph.gen_store.entries)
    assert len(hist) == nR + nG
    assert all(h['tag'].startswith('FlownetS//') and h['tag'].endswith('/gradients') for h in hist[:nR])
    assert all(h['tag'].startswith('MaskNet//') for h in hist[nR:]) and hist[nR]['tag'] == 'MaskNet//conv1/kernel/gradients'
    for h in (hist[0], hist[nR], hist[-1]):
        hh = h['histo']
        assert hh['num'] > 0 and -0.2001 <= hh['min'] <= hh['max'] <= 0.2001 and sum(hh['bucket']) == hh['num']
    assert any(h['histo']['max'] > 0 for h in hist[:nR]) and any(h['histo']['max'] > 0 for h in hist[nR:])
    # the summary pre-pass must not have changed what the two steps train: both nets' parameters still finite, one net updated per step
    after = L.graph.export_params()
    assert all(torch.isfinite(v).all() for v in after.values())
    assert any(not torch.equal(after[k], before[k]) for k in before if not k.startswith('pwcnet/'))
