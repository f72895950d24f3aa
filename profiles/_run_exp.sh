timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
ZIGP_LIB=profiles/microbench/libzigp_exp9.so timeout 200 python tests/oz_trace.py 0 2>/dev/null
ZIGP_LIB=profiles/microbench/libzigp_exp9.so timeout 200 python tests/oz_trace.py 1 2>/dev/null
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_full_v7.json 2>/dev/null; python -c "
import json
d=json.loads(open('gpurun_out/bench_full_v7.json').read().strip().splitlines()[-1])
print('full: value %.4g e2e %.4g ms %.1f frac %.3f clocks %s'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['roofline']['frac'],d['clocks']))"
