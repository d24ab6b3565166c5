"""Print the headline numbers of one bench.py JSON line (value, ms/step, e2e, by-kind)."""
import json
import sys

try:
    d = json.loads(open(sys.argv[1]).readline() or '{}')
    k = (d.get('config') or {}).get('ms_per_step_by_kind') or {}
    print('%.1f pairs/s  %.3f ms/step  e2e %.1f  R %.3f ms  G %.3f ms' % (d['value'], d['ms_per_step'], d['e2e']['value'], k.get('recover', 0),
                                                                       k.get('generator', 0)))
except Exception as e:
    print('no result (%s)' % e)
