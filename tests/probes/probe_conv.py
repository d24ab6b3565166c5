"""Scratch probe (GPU): prints the conv-engine errors per case to gpurun_out/probe_conv.txt."""
import sys, os, json, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from test_conv_engine_gpu import CASES
from convref import run_conv_case
os.makedirs('gpurun_out', exist_ok=True)
with open('gpurun_out/probe_conv.txt', 'w') as f:
    for c in CASES:
        try:
            r = run_conv_case(**c)
            f.write(json.dumps(dict(case={k: v for k, v in c.items()}, res=r)) + '\n')
        except Exception as e:
            f.write('EXC %s %s\n' % (c, traceback.format_exc()))
        f.flush()
print(open('gpurun_out/probe_conv.txt').read())
