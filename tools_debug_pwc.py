"""Scratch (GPU): run the PWC forward op by op and report when watched buffers first blow up."""
import os, sys, ctypes as C
import torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import params as OP
from unsupervised_detection_b200.step_graph import CISGraph
from unsupervised_detection_b200 import _lib
import torch.nn.functional as F
gen = torch.Generator().manual_seed(0)
p = OP.make_params(seed=1, jitter=0.1)
g2 = CISGraph(64, 96, 1, with_pwc=True, pwc_hw=(128, 192), train=False)
g2.load_params(p)
lo = torch.randn(1, 3, 8, 12, generator=gen)
img1 = (F.interpolate(lo, size=(128, 192), mode='bicubic') * 0.25).permute(0, 2, 3, 1).contiguous().clamp(-0.5, 0.5)
img2 = torch.roll(img1, shifts=(1, 2), dims=(1, 2))
g2.img1.copy_(img1); g2.img2.copy_(img2)
g2._ensure_pwc()
torch.cuda.synchronize()
watch = {}
for l in range(1, 7):
    watch['c1_%d' % l] = g2.pwc.c1[l]
    watch['c2_%d' % l] = g2.pwc.c2[l]
st = torch.cuda.current_stream().cuda_stream
seen = set()
for i, (fn, args, name) in enumerate(g2.fwd.ops):
    if fn is None:
        args()
    else:
        rc = fn(*args, st)
        assert rc == 0, (name, _lib.load().cis_last_error())
    torch.cuda.synchronize()
    info = ''
    if name == 'cis_conv_igemm':
        d = args[0]._obj
        info = 'N%d H%d W%d OH%d OW%d s%d taps%d nsrc%d src0(p%d,c%d,ch%d) K%d BN%d nt%d out(p%d,c%d,ch%d) act%d' % (
            d.N, d.H, d.W, d.OH, d.OW, d.sh, d.ntaps, d.nsrc, d.src[0].pitch, d.src[0].c_off, d.src[0].chunks, d.K_pad, d.BN, d.n_tiles,
            d.out_pitch, d.out_coff, d.out_ch, d.act)
    bad = []
    for k, a in watch.items():
        v = a.float()
        m = float(v.abs().max())
        if (m > 50 or m != m) and k not in seen:
            bad.append((k, m))
            seen.add(k)
    if i < 45 or bad:
        print(i, name, info, 'BAD' if bad else '', bad)
    if len(seen) > 3:
        break
