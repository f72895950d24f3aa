set -x
mkdir -p gpurun_out/final
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/final/bench_full.json 2> gpurun_out/final/bench_full.err; tail -c 600 gpurun_out/final/bench_full.json
timeout 300 python bench.py --impl reference > gpurun_out/final/bench_reference.json 2>gpurun_out/final/bench_reference.err; tail -c 400 gpurun_out/final/bench_reference.json
timeout 400 python bench.py --precision fp32 --no-cpu-baseline > gpurun_out/final/bench_fp32.json 2>gpurun_out/final/bench_fp32.err; tail -c 300 gpurun_out/final/bench_fp32.json
timeout 400 python bench.py --workload kron --points 67108864 --no-cpu-baseline > gpurun_out/final/bench_kron_2p26.json 2>gpurun_out/final/bench_kron.err; tail -c 300 gpurun_out/final/bench_kron_2p26.json
# profiler passes (after the same command exited 0 above without ncu)
timeout 300 python bench.py --points 131072 --steps 1 --warmup 3 --no-cpu-baseline --chunk 65536 > /dev/null && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/final/launches.csv python bench.py --points 131072 --steps 1 --warmup 3 --no-cpu-baseline --chunk 65536 > gpurun_out/final/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ozaki_gemm -s 8 -c 8 -o gpurun_out/final/int8_gemm -f python bench.py --points 131072 --steps 1 --warmup 3 --no-cpu-baseline --chunk 65536 > gpurun_out/final/ncu_full.log 2>&1
ls -la gpurun_out/final
