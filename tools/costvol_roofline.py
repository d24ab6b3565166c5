"""cis_warp_costvol on a batch >= 64 shape (its HBM roofline is only measurable on a large batched shape):
PWC-Net level 2, 96x160x32 features, batch 64.  Prints algorithmic bytes / CUDA-event time against the measured HBM peak.
Run it under `ncu --set full -k regex:warp_costvol` for the DRAM-traffic capture."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unsupervised_detection_b200 import _lib

B, h, w, C = int(os.environ.get('CV_BATCH', '64')), 96, 160, 32
lib = _lib.load()
g = torch.Generator(device='cuda').manual_seed(0)
c1 = torch.randn(B, h, w, C, device='cuda', generator=g).to(torch.bfloat16)
c2 = torch.randn(B, h, w, C, device='cuda', generator=g).to(torch.bfloat16)
flow = torch.randn(B, h, w, 2, device='cuda', generator=g) * 0.3
out = torch.zeros(B, h, w, 88, dtype=torch.bfloat16, device='cuda')
flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device='cuda')
st = torch.cuda.current_stream().cuda_stream
run = lambda: _lib.check(lib.cis_warp_costvol(c1.data_ptr(), C, 0, c2.data_ptr(), C, 0, flow.data_ptr(), 5.0, B, h, w, C, out.data_ptr(), 88, 0, st), 'costvol')
ms = []
for i in range(8):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record()
    torch.cuda.synchronize()
    if i >= 3:
        ms.append(e0.elapsed_time(e1))
ms.sort()
t = ms[len(ms) // 2]
alg = B * h * w * (2 * C * 2 + 8 + 81 * 2)            # read c1, c2 (bf16) + flow (2 fp32), write 81 bf16 channels
try:
    pk = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['hbm_gbs']
except Exception:
    pk = 6650.0
print(json.dumps(dict(kernel='cis::warp_costvol_kernel', shape=[B, h, w, C], ms=t, algorithmic_bytes=alg, achieved_gbs=alg / t / 1e6,
                      peak_gbs=pk, frac_measured=alg / t / 1e6 / pk, frac_nominal_8tbs=alg / t / 1e6 / 8000.0)))
