"""Join an ncu launch list (csv: gpu__time_duration.sum + sm__pipe_tensor_cycles_active per kernel) of tools/ncu_step.py with the conv-op
list it dumped (CIS_OPS_JSON) and print (a) every kernel family's share of the step, (b) the FLOP-weighted tensor-pipe utilisation per
layer class and (c) the top launches.  Usage: python tools/ncu_table.py launches.csv ops.json"""
import collections
import csv
import json
import sys

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.DictReader(lines)
per = collections.OrderedDict()
for r in rd:
    k = r['ID']
    d = per.setdefault(k, dict(name=r['Kernel Name'], grid=r.get('Grid Size', ''), m={}))
    try:
        d['m'][r['Metric Name']] = float(r['Metric Value'].replace(',', ''))
    except ValueError:
        pass
    d['unit_' + r['Metric Name']] = r['Metric Unit']
L = list(per.values())
for d in L:
    t = d['m'].get('gpu__time_duration.sum', 0.0)
    u = d.get('unit_gpu__time_duration.sum', 'ns')
    d['us'] = t / 1e3 if u in ('ns', 'nsecond') else t if u in ('us', 'usecond') else t * 1e3
    d['tensor'] = next((v for k, v in d['m'].items() if 'pipe_tensor' in k), 0.0)
tot = sum(d['us'] for d in L)
fam = collections.defaultdict(lambda: [0, 0.0])
for d in L:
    n = d['name'].split('(')[0].replace('void ', '').replace('cis::', '')
    n = n.split('<')[0]
    fam[n][0] += 1
    fam[n][1] += d['us']
print('%d kernel launches, %.1f us summed device time (serialised by ncu, cold caches: compare SHARES)' % (len(L), tot))
for n, (c, t) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
    print('  %-34s %5d launches %9.1f us  %5.1f %%' % (n, c, t, 100 * t / tot))
ops = json.load(open(sys.argv[2]))
CONV = ('conv_halo_kernel', 'conv_halo_persist_kernel', 'conv_igemm_kernel', 'conv_wgrad_kernel', 'conv_wgrad_halo_kernel')
K = [d for d in L if any(c in d['name'] for c in CONV)]
if len(K) != len(ops):
    print('WARNING: %d conv kernels vs %d conv ops -- positional join skipped' % (len(K), len(ops)))
    sys.exit(0)
cls = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for d, o in zip(K, ops):
    d['flops'], d['info'] = o['flops'], o['info']
    kind = 'wgrad' if o['op'] == 'cis_conv_wgrad' else ('persist' if 'persist' in d['name'] else o['info'].split()[0])
    bn = o['info'].split()[1] if o['op'] == 'cis_conv_igemm' else ''
    c = cls[(kind, bn)]
    c[0] += 1; c[1] += d['us']; c[2] += o['flops']; c[3] += o['flops'] * d['tensor']
print('\nlayer class            launches      us     GFLOP   TF/s  FLOP-weighted tensor-pipe %% (sm__pipe_tensor_cycles_active, elapsed)')
tf, tw, tu = 0.0, 0.0, 0.0
for (kind, bn), (c, us, fl, w) in sorted(cls.items(), key=lambda kv: -kv[1][1]):
    print('  %-8s %-6s %9d %9.1f %9.1f %6.0f  %5.1f' % (kind, bn, c, us, fl / 1e9, fl / us / 1e6 if us else 0, w / fl if fl else 0))
    tf += fl; tw += w; tu += us
print('  %-15s %9d %9.1f %9.1f %6.0f  %5.1f' % ('all conv', len(K), tu, tf / 1e9, tf / tu / 1e6, tw / tf))
print('\ntop 12 conv launches by FLOPs:')
for d in sorted(K, key=lambda d: -d['flops'])[:12]:
    print('  %7.1f us %6.1f GF %6.0f TF/s tensor %5.1f %%  %s' % (d['us'], d['flops'] / 1e9, d['flops'] / d['us'] / 1e6, d['tensor'], d['info']))
