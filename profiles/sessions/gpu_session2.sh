#!/bin/bash
# GPU session: parity tests, re-alignment tuning sweep, bench lines, launch list, ncu captures. Outputs -> gpurun_out/
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== sweep (mid)"
for cfg in "4096 1024" "4096 4096" "2048 2048" "8192 1024"; do
  set -- $cfg
  ARB_MISMAP_BUDGET=$1 ARB_MISMAP_LANES=$2 timeout 600 python bench.py --workload mid_1M_2x101_5k --steps 2 --warmup 1 --no-cpu-baseline 2> /dev/null > gpurun_out/sweep_$1_$2.json
  python - <<PY
import json
d=json.load(open("gpurun_out/sweep_$1_$2.json")); r=d["roofline"]
print("budget $1 lanes $2", {k: round(v,1) for k,v in r["device_ms"].items() if k.startswith("mism")}, r.get("mismapper_heavy_items"), "e2e s", round(d["e2e"]["seconds_per_step"],2), "out", d["e2e"]["output_seconds"])
PY
done
echo "== bench mid"; timeout 900 python bench.py --workload mid_1M_2x101_5k --steps 3 --warmup 3 > gpurun_out/bench_mid.json 2> gpurun_out/bench_mid.err; tail -c 3500 gpurun_out/bench_mid.json
echo "== bench cfg2 (default)"; timeout 1800 python bench.py > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; tail -c 3500 gpurun_out/bench_cfg2.json
echo "== ncu launch list (mid)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_mid.csv python bench.py --workload mid_1M_2x101_5k --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -1 gpurun_out/ncu_bench.log | cut -c1-200
echo "== ncu full capture: cascade + re-alignment kernels"
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:cascade_head_fn|cascade_sequences_fn|mismap_item_fn|mismap_heavy_fn' -c 4 -o gpurun_out/prof_r01b -f python bench.py --workload mid_1M_2x101_5k --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-200
ls -la gpurun_out | head -40
