#!/bin/bash
# A/B of the balanced one-wave launch of kgrad_planes_kernel (ZIGP_KG_BALANCED) at N = 2^22, same box, back to back.
# (The recorded run, profiles/ab_coll_balanced_r02.txt, also switched a collector-buffer variant of the contraction's issue loop,
# ZIGP_OZ_COLL, that was removed afterwards -- see the header of that file.)
for rep in 1 2; do
for bal in 1 0; do
  ZIGP_KG_BALANCED=$bal python bench.py --points 4194304 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']
print('balanced=$bal ms_per_step %.1f frac %.4f' % (d['ms_per_step'], r['frac']), {k: round(v,1) for k,v in r['class_ms'].items() if v > 1}, 'elbo', repr(d['config']['elbo']), d['clocks'])
"
done
done
