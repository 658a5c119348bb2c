#!/bin/bash
# One GPU session: parity tests, bench lines, ncu launch list + full capture of the dominant kernel. Outputs -> gpurun_out/
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== bench mid"; timeout 900 python bench.py --workload mid_1M_2x101_5k --steps 3 --warmup 3 > gpurun_out/bench_mid.json 2> gpurun_out/bench_mid.err; tail -c 3000 gpurun_out/bench_mid.json; tail -3 gpurun_out/bench_mid.err
if [ "${1:-}" != "quick" ]; then
echo "== bench cfg2 (default)"; timeout 1800 python bench.py > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; tail -c 3000 gpurun_out/bench_cfg2.json; tail -3 gpurun_out/bench_cfg2.err
fi
echo "== ncu launch list (mid)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_mid.csv python bench.py --workload mid_1M_2x101_5k --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300
echo "== ncu full capture of the cascade kernel"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:classify_fn -c 1 -o gpurun_out/prof_classify -f python bench.py --workload mid_1M_2x101_5k --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-300
ls -la gpurun_out | head -30
