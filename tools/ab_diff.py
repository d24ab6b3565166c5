"""Compare two per-launch timing dumps of tools/time_ops.py (TIME_OPS_JSON=...), layer by layer.
Usage: python tools/ab_diff.py base.json variant.json [min_us_delta]
Launch lists are matched positionally per plan (same network, same order); a layer that changed kernel shows both infos."""
import json
import sys

a, b = json.load(open(sys.argv[1])), json.load(open(sys.argv[2]))
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0


def per_step(rows):
    return sum(r['w'] * r['us'] for r in rows) / 4.0


print('per-step sum: base %.1f us, variant %.1f us (%+.1f)' % (per_step(a), per_step(b), per_step(b) - per_step(a)))
if len(a) != len(b):
    print('launch lists differ in length (%d vs %d): positional matching only up to the shorter one' % (len(a), len(b)))
rows = []
for x, y in zip(a, b):
    if x['plan'] != y['plan'] or x['op'] != y['op']:
        continue
    d = (y['us'] - x['us']) * x['w'] / 4.0
    if abs(d) >= thr:
        rows.append((d, x, y))
for d, x, y in sorted(rows, key=lambda r: r[0]):
    tag = x['info'] if x['info'] == y['info'] else '%s  ->  %s' % (x['info'], y['info'])
    print('%+8.1f us/step  %-5s x%d %-22s %7.1f -> %7.1f us  %s' % (d, x['plan'], x['w'], x['op'], x['us'], y['us'], tag))
