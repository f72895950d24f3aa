timeout 600 python -m pytest tests/test_gpu_int8_engine.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
run() {
env $1 timeout 200 python bench.py --points 1048576 --steps 3 --warmup 3 --no-cpu-baseline --chunk 65536 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
c=d[\"roofline\"][\"class_ms\"]
print(\"$1 ms/step %.1f elbo %.15e gemm %.1f syrk %.1f slice %.1f kgen %.1f kgrad %.1f clocks %s\"%(d[\"ms_per_step\"],d[\"config\"][\"elbo\"],c[\"gemm_tn\"],c[\"syrk\"],c[\"slice\"],c[\"kgen\"],c[\"kgrad\"],d[\"clocks\"][\"sm_mhz\"]))"
}
run "ZIGP_X=slicecols2"
