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
torch.cuda.profiler.start()
for m in os.environ.get('CIS_MODES', 'GR'):
    g.train_step(m)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('done', g.losses())
