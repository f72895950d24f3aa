#!/bin/bash
# A/B of the fused Kronecker factor kernels at N = 2^26, 64 x 64 grid, same box, back to back:
# ZIGP_KRON_FUSED = 2 (default: factor chains + the Ms x Mt kernels), 1 (factor chains only), 0 (unfused launches); second field: ZIGP_KF_CTAS.
for rep in 1 2; do
for cfg in "2 3" "1 3" ${KRON_AB_UNFUSED:+"0 3"}; do
  set -- $cfg
  ZIGP_KRON_FUSED=$1 ZIGP_KF_CTAS=$2 python bench.py --workload kron --points 67108864 --chunk 2097152 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']
print('fused=$1 ctas=$2 ms_per_step %.1f Mpts/s %.2f' % (d['ms_per_step'], d['value'] / 1e6), {k: round(v,1) for k,v in r['class_ms'].items() if v > 1}, 'launches', r['class_launches']['gemm_tn'], 'elbo', repr(d['config']['elbo']), d['clocks']['sm_mhz'])
"
done
done
