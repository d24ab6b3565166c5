import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d['roofline']
print('value %.1f pairs/s  %.3f ms/step  e2e %.1f  dominant %.0f TF/s (frac %.3f, %.1f us)  conv family %.1f TF/s  conv_ms %.3f' % (
    d['value'], d['ms_per_step'], d['e2e']['value'], r['achieved'], r['frac'], r['us_per_launch'], r['conv_family']['achieved'], r['conv_family']['ms_per_step']))
