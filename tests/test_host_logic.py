"""Host-side logic of the product package (no GPU): flags, metric, parameter layout, tap tables / channel maps, launch
lists, C-ABI surface."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from oracle import metrics as OM, params as OP
from unsupervised_detection_b200 import _lib, engine as E
from unsupervised_detection_b200.common_flags import FLAGS, FLAG_NAMES, Config
from unsupervised_detection_b200.models.utils import general_utils as GU

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_flag_surface_matches_reference():
    # common_flags.py:5-55 : 31 flags, names + defaults
    assert len(FLAG_NAMES) == 31
    exp = dict(img_width=384, img_height=192, batch_size=16, beta1=0.9, flow_normalizer=80.0, max_epochs=40, num_samples_train=5000,
               train_crop=0.9, max_temporal_len=2, min_temporal_len=1, cbn=0.5, epsilon=75.0, iters_rec=1, iters_gen=3, num_threads=6,
               resume_train=False, train_partition='trainval', dataset='DAVIS2016', summary_freq=30, save_freq=5, test_crop=0.9,
               test_temporal_shift=1, test_partition='val', generate_visualization=False)
    for k, v in exp.items():
        assert FLAGS[k].default == v, k
    c = Config(img_height=256, img_width=448)
    assert (c.img_height, c.img_width, c.iters_gen) == (256, 448, 3)
    with pytest.raises(AttributeError):
        Config(not_a_flag=1)


def test_cabi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, 'include', 'cis_b200.h')).read()
    declared = sorted(set(re.findall(r'\b(cis_[a-z0-9_]+)\s*\(', hdr)) - {'cis_stream_t'})
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(declared) == sorted(_lib.EXPORTS)
    assert lib.cis_version() >= 100


def test_ctypes_structs_match_the_c_header(tmp_path):
    import shutil
    import subprocess
    if shutil.which('gcc') is None:
        pytest.skip('no gcc')
    src = tmp_path / 'sz.c'
    src.write_text('#include "%s"\n#include <stdio.h>\n#include <stddef.h>\nint main(){printf("%%zu %%zu %%zu %%zu %%zu %%zu\\n", sizeof(CisConv), '
                   'sizeof(CisWgrad), offsetof(CisConv, src), offsetof(CisConv, mode), offsetof(CisConv, ex), offsetof(CisWgrad, splits));return 0;}\n'
                   % os.path.join(ROOT, 'include', 'cis_b200.h'))
    exe = tmp_path / 'sz'
    subprocess.check_call(['gcc', str(src), '-o', str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    exp = [C.sizeof(_lib.CisConv), C.sizeof(_lib.CisWgrad), _lib.CisConv.src.offset, _lib.CisConv.mode.offset, _lib.CisConv.ex.offset,
           _lib.CisWgrad.splits.offset]
    assert got == exp


def test_bad_descriptor_is_rejected_without_a_gpu():
    d = _lib.CisConv()
    rc = _lib.load().cis_conv_igemm(C.byref(d), None)
    assert rc == 1 and b'bad descriptor' in _lib.load().cis_last_error()
    with pytest.raises(RuntimeError):
        _lib.check(rc, 'cis_conv_igemm')


def test_metric_matches_oracle():
    g = torch.Generator().manual_seed(0)
    pm = torch.rand(3, 12, 16, 1, generator=g)
    gt = (torch.rand(3, 12, 16, 1, generator=g) > 0.5).float()
    a = GU.compute_all_IoU(pm.numpy(), gt.numpy())
    b = OM.compute_all_IoU(pm, gt).numpy()
    assert np.allclose(a, b, atol=1e-6)
    pm[0] = 0.9                                   # border-hugging => complemented
    assert np.allclose(GU.compute_all_IoU(pm.numpy(), gt.numpy()), OM.compute_all_IoU(pm, gt).numpy(), atol=1e-6)
    iou, ann = GU.compute_IoU(np.zeros((4, 4)), np.zeros((4, 4), np.float32))
    assert iou == 1.0


def test_same_pad_and_taps():
    st = E.ParamStore('cpu')
    L = E.ConvLayer(st, 'l', 3, 8, 8, stride=2)
    taps, pt, pl = L.fwd_taps(8, 8)
    assert (pt, pl) == (0, 0) and taps[0] == (0, 0) and taps[-1] == (2, 2)
    L7 = E.ConvLayer(st, 'l7', 7, 8, 8, stride=2)
    assert L7.fwd_taps(8, 8)[1:] == (2, 2)
    L4 = E.ConvLayer(st, 'l4', 4, 8, 8)
    t4, pt, _ = L4.fwd_taps(8, 8)
    assert pt == 1 and t4[0] == (-1, -1) and t4[-1] == (2, 2)
    Ld = E.ConvLayer(st, 'ld', 3, 8, 8, dil=4)
    assert Ld.fwd_taps(16, 16)[0][0] == (-4, -4)


@pytest.mark.parametrize('k,s', [(3, 2), (5, 2), (7, 2), (3, 1), (4, 1)])
def test_dgrad_parity_decomposition_covers_every_tap_once(k, s):
    st = E.ParamStore('cpu')
    L = E.ConvLayer(st, 'l', k, 8, 16, stride=s, tag='R')
    st.finalize(True)
    L.setup_fwd(list(range(8)))
    L.setup_dgrad(16, 20)
    assert len(L.dgrad_packs) == s * s
    assert sum(len(p['taps']) for p in L.dgrad_packs) == k * k
    # every (tap, co) position maps to a distinct HWIO offset
    seen = set()
    for p in L.dgrad_packs:
        km = p['kmap'].numpy()
        for v in km[km >= 0]:
            assert int(v) not in seen
            seen.add(int(v))
    assert len(seen) == k * k * 16


def test_virtual_concat_channel_map():
    st = E.ParamStore('cpu')
    L = E.ConvLayer(st, 'l', 3, 34, 16, tag='R')
    st.finalize(True)
    B = E.Builder('cpu')
    a = B.new_act(1, 4, 4, 32, dep={'R'})
    b = B.new_act(1, 4, 4, 2, dep={'R'})          # 2 real channels padded to 8
    B.conv(L, [a, b])
    assert L.in_chanmap == list(range(32)) + [32, 33] + [-1] * 6
    km = L.fwd_kmap.numpy()
    assert L.K_pad == 384 and km[0] == 0 and km[32] == 32 * 16 and km[34] == -1 and km[40] == 34 * 16


def test_param_store_layout_and_counts():
    from unsupervised_detection_b200.step_graph import CISGraph
    g = CISGraph(32, 48, 1, device='cpu', with_pwc=True, pwc_hw=(64, 64))
    assert g.gen_store.real_count() == 1451062      # adversarial_learner.py:338 "Number of params" split by scope
    assert g.rec_store.real_count() == 3388610
    assert g.pwc_store.real_count() == 14079050
    p = OP.make_params(seed=5)
    g.load_params(p)
    ex = g.export_params()
    assert set(ex) == set(p)
    for k in ('MaskNet/conv1/kernel', 'FlownetS/flow1/weights', 'pwcnet/upsample/up_feat3/kernel', 'MaskNet/conv17/gamma'):
        assert torch.equal(ex[k], p[k])
    # launch lists exist for both step kinds and the generator-step backward touches only 2B of the 3B recover batch
    assert g.bwd['R'].count() > 100 and g.bwd['G'].count() > 100
    assert g.rec_in.gen_rows == 2 and g.rec.flow1.gen_rows == 2


def test_step_schedule_and_global_step():
    from unsupervised_detection_b200.models.adversarial_learner import AdversarialLearner

    class FakeGraph(object):
        def __init__(self):
            self.modes = []

        def train_step(self, mode, allreduce=None, use_graph=False):
            self.modes.append(mode)

        def losses(self, full=False, reduce=None):
            return dict(generator=0.0, recover=0.0)
    L = AdversarialLearner()
    L.config = Config(summary_freq=4)
    L.graph = FakeGraph()
    L.feed = lambda a, b: None
    L._allreduce = lambda: None
    L.world, L.rank = 1, 0
    out = [L.step(batch=(None, None)) for _ in range(8)]
    assert L.graph.modes == ['G', 'G', 'G', 'R', 'G', 'G', 'G', 'R']          # adversarial_learner.py:386-389
    assert [o['global_step'] for o in out] == [0, 0, 0, 1, 1, 1, 1, 2]        # :382-384
    assert 'loss_generator' in out[3] and 'loss_generator' not in out[2]      # :391-394 summary_freq


def test_learner_error_conventions():
    from unsupervised_detection_b200.models.adversarial_learner import AdversarialLearner
    L = AdversarialLearner()
    L.config = Config(dataset='NOPE')
    L.rank = 0
    with pytest.raises(IOError):
        L.load_training_data()
    L.config = Config(dataset='DAVIS2016', root_dir='/nonexistent')
    with pytest.raises(IOError):
        L.load_training_data()


def test_train_loop_control_flow_on_cpu(capsys):
    """AdversarialLearner.train (adversarial_learner.py:376-420) with the GPU parts stubbed: every batch is consumed once and in order,
    the NEXT batch is handed to step() for the overlapped host->device copy, epochs end after num_samples_train/batch_size steps and
    training stops after max_epochs."""
    from unsupervised_detection_b200.common_flags import Config
    from unsupervised_detection_b200.models.adversarial_learner import AdversarialLearner

    class Reader(object):
        def __init__(self):
            self.n = 0

        def batch(self, b):
            self.n += 1
            return ('img1_%d' % self.n, 'img2_%d' % self.n, None, [])

    class Graph(object):
        def param_count(self):
            return 123

    class Stub(AdversarialLearner):
        def build_train_graph(self):
            self.rank, self.world, self.local_batch = 0, 1, 2
            self.reader, self.graph = Reader(), Graph()
            self.train_steps_per_epoch = 3
            self.calls, self.epochs = [], []

        def step(self, batch=None, fetch_losses=None, use_graph=True, next_batch=None, summarize=False):
            self.calls.append((batch[0], next_batch[0], summarize))
            return {'global_step': len(self.calls) // 4, 'train_op': 'G', 'loss_generator': 1.0, 'loss_recover': 2.0}

        def epoch_end_callback(self, sess, sv, epoch_num):
            self.epochs.append((epoch_num, len(self.calls)))

    L = Stub()
    L.train(Config(max_epochs=2, summary_freq=2, checkpoint_dir=''))
    assert [c[0] for c in L.calls] == ['img1_%d' % i for i in range(1, 7)]            # 2 epochs x 3 steps, each batch used once, in order
    assert [c[1] for c in L.calls] == ['img1_%d' % i for i in range(2, 8)]            # step k is given batch k+1 to prefetch
    assert all(c[2] for c in L.calls) and L.epochs == [(1, 3), (2, 6)]
    out = capsys.readouterr().out
    assert 'Number of params: 123' in out and 'Training completed successfully' in out and out.count('loss_generator') == 3


def test_epoch_end_callback_validation_and_saving_on_cpu(tmp_path):
    """adversarial_learner.py:422-448: validation IoU = sum over batches / (steps * batch_size); model.best on improvement, model-<epoch>
    every save_freq epochs; the IoU goes to the event file."""
    import numpy as np
    import torch
    from unsupervised_detection_b200.common_flags import Config
    from unsupervised_detection_b200.models.adversarial_learner import AdversarialLearner
    from unsupervised_detection_b200.summary import SummaryWriter, read_events

    gt = torch.zeros(2, 16, 24, 1)
    gt[:, 4:12, 6:18] = 1.0

    class Graph(object):
        def __init__(self):
            self.mask = None
            self.quality = 0

        def forward(self):
            m = torch.zeros(2, 8, 12, 1)
            if self.quality == 1:                      # half of the object
                m[:, 2:6, 3:6] = 0.9
            elif self.quality == 2:                    # the whole object (ground truth is NN-resized to the mask size)
                m[:, 2:6, 3:9] = 0.9
            self.mask = m

    class ValReader(object):
        def batch(self, b):
            return torch.zeros(2, 384, 640, 3), torch.zeros(2, 384, 640, 3), gt, ['a', 'b']

    class Stub(AdversarialLearner):
        def feed(self, img1, img2):
            self.fed = getattr(self, 'fed', 0) + 1

        def save(self, sess, checkpoint_dir, step):
            self.saved.append(step)

    L = Stub()
    L.rank, L.world, L.local_batch, L.device = 0, 1, 2, 'cpu'
    L.config = Config(batch_size=2, save_freq=2, checkpoint_dir=str(tmp_path))
    L.graph, L.val_reader, L.reader = Graph(), ValReader(), None
    L.val_steps_per_epoch, L.min_val_iou, L.saved = 3, -1.0e12, []
    L.summary_writer = SummaryWriter(str(tmp_path))
    for epoch, q in ((1, 1), (2, 0), (3, 2), (4, 2)):
        L.graph.quality = q
        L.epoch_end_callback(None, None, epoch)
    assert L.fed == 12
    # epoch 1: first result is the best so far; epoch 2: worse, but save_freq; epoch 3: better; epoch 4: equal -> only save_freq
    assert L.saved == ['best', 2, 'best', 4]
    L.summary_writer.close()
    ious = [e['values'][0]['simple_value'] for e in read_events(L.summary_writer.path)[1:]]
    assert [e['step'] for e in read_events(L.summary_writer.path)[1:]] == [1, 2, 3, 4]
    assert abs(ious[0] - 0.5) < 1e-6 and ious[1] == 0.0 and abs(ious[2] - 1.0) < 1e-6 and abs(ious[3] - 1.0) < 1e-6
