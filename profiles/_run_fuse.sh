echo "== parity fuse=2"
ZIGP_OZ_FUSE=2 timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_int8_engine.py tests/test_variants.py -x -q -m gpu 2>&1 | tail -2
run() {
env $1 timeout 200 python bench.py --points 1048576 --steps 3 --warmup 3 --no-cpu-baseline --chunk 65536 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
c=d[\"roofline\"][\"class_ms\"]
print(\"$1 value %.3f Mpts/s  ms/step %.1f gemm %.1f slice %.1f kgen %.1f syrk %.1f kgrad %.1f\"%(d[\"value\"]/1e6,d[\"ms_per_step\"],c[\"gemm_tn\"],c[\"slice\"],c[\"kgen\"],c[\"syrk\"],c[\"kgrad\"]))"
}
run "ZIGP_OZ_FUSE=0"
run "ZIGP_OZ_FUSE=1"
run "ZIGP_OZ_FUSE=2"
run "ZIGP_OZ_FUSE=1 ZIGP_OZ_SLDBG=1"
run "ZIGP_OZ_FUSE=1 ZIGP_OZ_SLDBG=2"
