"""Developer tool: per-step clock trace of the MMA-issue loop of selected conv launches (CTA 0 only).

  make -C unsupervised_detection_b200/csrc trace
  CIS_LIB_NAME=libcis_b200_trace.so python tools/trace_conv.py [max_ops]

For every distinct conv launch shape of the step (slowest first, measured warm) it prints: kernel time, the clocks from CTA start to
the first operand, the per-step (wait, issue) clocks of the MMA warp, accumulator-ready and end-of-kernel stamps.  This is what
tells a latency-bound pipeline (long waits per step) from an issue-bound one."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault('CIS_LIB_NAME', 'libcis_b200_trace.so')
import torch  # noqa: E402
from unsupervised_detection_b200 import _lib  # noqa: E402
from unsupervised_detection_b200.common_flags import Config  # noqa: E402
from unsupervised_detection_b200.models.adversarial_learner import AdversarialLearner  # noqa: E402

MAX_OPS = int(sys.argv[1]) if len(sys.argv) > 1 else 40
CAP = 8 + 2 * 512
lib = _lib.load()
lib.cis_trace_set.argtypes = [C.c_void_p, C.c_int]
lib.cis_trace_set.restype = C.c_int

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
seen, ops = set(), []
for pname, plan in (('fwd', g.fwd), ('bwdG', g.bwd['G']), ('bwdR', g.bwd['R'])):
    for fn, a, name, fl, lane in plan.ops:
        if name != 'cis_conv_igemm':
            continue
        d = a[0]._obj
        ch = sum(d.src[k].chunks for k in range(d.nsrc)) * 8
        info = '%s BN%d nt%d MT%d N%d %dx%d taps%d ch%d dil%d sp%d' % ('halo' if d.halo else 'gen', d.BN, d.n_tiles, d.MT, d.N, d.OH, d.OW, d.ntaps,
                                                                     ch, d.dil, d.splits)
        if info in seen:
            continue
        seen.add(info)
        ops.append((pname, fn, a, info, fl))


def timed(fn, a, reps=5):
    fn(*a, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn(*a, st)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


rows = [(timed(fn, a), pname, fn, a, info, fl) for pname, fn, a, info, fl in ops]
rows.sort(key=lambda r: -r[0])
print('%d distinct conv launch shapes; tracing the %d slowest (eager launches, warm L2)' % (len(rows), min(MAX_OPS, len(rows))))
for us, pname, fn, a, info, fl in rows[:MAX_OPS]:
    buf.zero_()
    lib.cis_trace_set(buf.data_ptr(), CAP)
    fn(*a, st)
    torch.cuda.synchronize()
    lib.cis_trace_set(None, 0)
    t = buf.tolist()
    t0 = t[0]
    steps = [(t[8 + 2 * i], t[9 + 2 * i]) for i in range((CAP - 8) // 2) if t[8 + 2 * i]]
    if not t0 or not steps:
        print('%-5s %7.1f us  %s  (no trace: persistent kernel or CTA 0 idle)' % (pname, us, info))
        continue
    waits = [steps[0][0] - t0] + [steps[i][0] - steps[i - 1][1] for i in range(1, len(steps))]
    issue = [b_ - a_ for a_, b_ in steps]
    sw = sorted(waits[1:]) or [0]
    print('%-5s %7.1f us %6.1f GF  %s' % (pname, us, fl / 1e9, info))
    print('      steps %d | first operand after %d clk (halo ready %d) | wait/step: median %d mean %d max %d | issue/step: mean %d | '
          'loop %d clk | accum-ready +%d | epilogue+exit %d | total CTA %d clk' %
          (len(steps), waits[0], (t[1] - t0) if t[1] else -1, sw[len(sw) // 2], sum(waits[1:]) // max(1, len(waits) - 1), sw[-1],
           sum(issue) // len(issue), steps[-1][1] - steps[0][0], (t[2] - steps[-1][1]) if t[2] else -1, (t[3] - t[2]) if t[3] and t[2] else -1,
           (t[3] - t0) if t[3] else -1))
    if len(steps) <= 40:
        print('      waits: ' + ' '.join(str(w) for w in waits))
