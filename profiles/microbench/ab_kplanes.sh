#!/bin/bash
# A/B of ZIGP_OZ_KPLANES (K not stored as fp64: the gradient kernel recomputes it) at N = 2^22, same box, back to back.
for rep in 1 2; do
for kp in 0 1; do
  ZIGP_OZ_KPLANES=$kp python bench.py --points 4194304 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']
print('kplanes=$kp ms_per_step %.1f frac %.4f' % (d['ms_per_step'], r['frac']), {k: round(v,1) for k,v in r['class_ms'].items() if v > 1}, 'elbo', repr(d['config']['elbo']), d['clocks']['sm_mhz'])
"
done
done
