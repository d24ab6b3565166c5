"""Target for ncu: one generator step + one recover step of the full-size workload (256x448, batch 4, PWC-Net in loop),
launched eagerly (no CUDA graph) between cudaProfilerStart/Stop.  Use with `ncu --profile-from-start off ...`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unsupervised_detection_b200.common_flags import Config
from unsupervised_detection_b200.models.adversarial_learner import AdversarialLearner

B = int(os.environ.get('CIS_BATCH', '4'))
L = AdversarialLearner()
L.config = Config(img_height=256, img_width=448, batch_size=B, dataset='SYNTHETIC', flow_ckpt='synthetic', summary_freq=10 ** 9)
L.build_train_graph()
batch = L.reader.batch(B)
L.feed(batch[0], batch[1])
g = L.graph
for m in ('G', 'R'):
    g.train_step(m)
torch.cuda.synchronize()
if os.environ.get('CIS_OPS_JSON'):
    # conv ops in issue order (what ncu will see, kernel by kernel) with their algorithmic FLOPs: lets tools/ncu_table.py join the launch
    # list with the layer descriptions
    import json
    rows = []
    for m in os.environ.get('CIS_MODES', 'GR'):
        for pname, plan in (('fwd', g.fwd), ('bwd' + m, g.bwd[m])):
            for fn, a, name, fl, lane in plan.ops:
                if name == 'cis_conv_igemm':
                    d = a[0]._obj
                    info = '%s BN%d nt%d MT%d N%d %dx%d taps%d ch%d dil%d sp%d' % ('halo' if d.halo else 'gen', d.BN, d.n_tiles, d.MT, d.N, d.OH, d.OW,
                                                                               d.ntaps, sum(d.src[k].chunks for k in range(d.nsrc)) * 8, d.dil, d.splits)
                elif name == 'cis_conv_wgrad':
                    d = a[0]._obj
                    info = 'wgrad tma%d N%d %dx%d taps%d cout%d K%d sp%d' % (d.tma, d.N, d.OH, d.OW, d.ntaps, d.Cout, d.K_pad, d.splits)
                else:
                    continue
                rows.append(dict(step=m, plan=pname, op=name, flops=fl, info=info))
    json.dump(rows, open(os.environ['CIS_OPS_JSON'], 'w'))
torch.cuda.profiler.start()
for m in os.environ.get('CIS_MODES', 'GR'):
    g.train_step(m)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('done', g.losses())
