import json,sys
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); print("%.4e steps/s  ms/step %.3f  e2e %.4e  meanR %.1f frac %.3f" % (d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['mean_rules_in_timed_region'], d['roofline']['frac']))
