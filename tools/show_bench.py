import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('value %.1f pairs/s  %.3f ms/step  e2e %.1f  conv %.1f TF/s  conv_ms %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['conv_ms_per_step']))
