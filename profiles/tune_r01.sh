#!/bin/bash
# engine tuning sweep (dev size N=2^20, M=1024, D=3): CTA tile / CTAs per SM / slab depth
P="zero-inflated-gp_b200"
for v in "$@"; do
    echo "== $v"
    ZIGP_LIB=$PWD/$P/libzigp_$v.so python bench.py --points 1048576 --steps 2 --warmup 3 --no-cpu-baseline --chunk 65536 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('  value %.3f Mpts/s  ms/step %.1f  gemm frac %.3f  class_ms %s'%(d['value']/1e6,d['ms_per_step'],d['roofline']['frac'],d['roofline']['class_ms']))"
done
