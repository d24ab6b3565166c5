"""Developer tool: per-tile clock trace of the persistent halo kernel (CTA 0).  CIS_LIB_NAME=libcis_b200_trace.so python tools/trace_persist.py
Per tile: wait for a free accumulator stage | wait for the first halo | MMA issue until the tile's last commit | epilogue done (relative)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault('CIS_LIB_NAME', 'libcis_b200_trace.so')
import torch  # noqa: E402
from unsupervised_detection_b200 import _lib  # noqa: E402
from unsupervised_detection_b200.common_flags import Config  # noqa: E402
from unsupervised_detection_b200.models.adversarial_learner import AdversarialLearner  # noqa: E402

CAP = 8 + 5 * 64
lib = _lib.load()
lib.cis_trace_set.argtypes = [C.c_void_p, C.c_int]
L = AdversarialLearner()
L.config = Config(img_height=256, img_width=448, batch_size=4, dataset='SYNTHETIC', flow_ckpt='synthetic', summary_freq=10 ** 9)
L.build_train_graph()
b = L.reader.batch(4)
L.feed(b[0], b[1])
g = L.graph
for m in 'GR':
    g.train_step(m)
torch.cuda.synchronize()
st = torch.cuda.current_stream().cuda_stream
buf = torch.zeros(CAP, dtype=torch.int64, device='cuda')
seen = set()
for pname, plan in (('fwd', g.fwd), ('bwdG', g.bwd['G']), ('bwdR', g.bwd['R'])):
    for fn, a, name, fl, lane in plan.ops:
        if name != 'cis_conv_igemm':
            continue
        d = a[0]._obj
        ch = sum(d.src[k].chunks for k in range(d.nsrc)) * 8
        info = 'BN%d MT%d N%d %dx%d taps%d ch%d' % (d.BN, d.MT, d.N, d.OH, d.OW, d.ntaps, ch)
        if not d.halo or info in seen:
            continue
        seen.add(info)
        buf.zero_()
        lib.cis_trace_set(buf.data_ptr(), CAP)
        fn(*a, st)
        torch.cuda.synchronize()
        lib.cis_trace_set(None, 0)
        t = buf.tolist()
        tiles = [(t[8 + 5 * i], t[9 + 5 * i], t[10 + 5 * i], t[11 + 5 * i], t[12 + 5 * i]) for i in range(64) if t[11 + 5 * i]]
        if not t[0] or not t[4] or len(tiles) < 1:
            continue          # not the persistent kernel
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn(*a, st)
        e1.record()
        torch.cuda.synchronize()
        print('%-5s %6.1f us  %s  tiles/CTA %d' % (pname, e0.elapsed_time(e1) * 200, info, len(tiles)))
        print('      first tile starts %d clk after CTA start' % (tiles[0][0] - t[0]))
        for i, (a0, a1, a2, a3, a4) in enumerate(tiles[:8]):
            print('      tile %d: acc-stage wait %5d | halo wait %5d | issue %5d | epilogue done +%5d after commit | tile period %5d' %
                  (i, a1 - a0, a2 - a1, a3 - a2, (a4 - a3) if a4 else -1, (tiles[i + 1][0] - a0) if i + 1 < len(tiles) else 0))
