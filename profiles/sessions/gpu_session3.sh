#!/bin/bash
# GPU session: parity tests, bench mid, re-alignment task sweep, cfg2 single step with progress, ingest laps, launch list + ncu captures. Outputs -> gpurun_out/
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== bench mid"; timeout 900 python bench.py --workload mid_1M_2x101_5k --steps 3 --warmup 3 > gpurun_out/bench_mid.json 2> gpurun_out/bench_mid.err; grep "^\[bench\]" gpurun_out/bench_mid.err | tail -2; tail -c 1800 gpurun_out/bench_mid.json
echo "== sweep (mid): budget lanes spawn task_lanes"
for cfg in "4096 256 2048 32" "4096 1024 2048 32" "4096 256 512 32" "4096 256 2048 128" "2048 512 1024 64" "4096 1024 0 1"; do
  set -- $cfg
  ARB_MISMAP_BUDGET=$1 ARB_MISMAP_LANES=$2 ARB_MISMAP_SPAWN=$3 ARB_MISMAP_TASK_LANES=$4 timeout 600 python bench.py --workload mid_1M_2x101_5k --steps 2 --warmup 1 --no-cpu-baseline 2> /dev/null > gpurun_out/sweep2_$1_$2_$3_$4.json
  python - <<PY
import json
d=json.load(open("gpurun_out/sweep2_$1_$2_$3_$4.json")); r=d["roofline"]
print("budget $1 lanes $2 spawn $3 task_lanes $4", {k: round(v,1) for k,v in r["device_ms"].items() if k.startswith("mism")}, "heavy", r.get("mismapper_heavy_items"), "tasks", r.get("mismapper_tasks"), "rounds", r.get("mismapper_rounds"), "e2e s", round(d["e2e"]["seconds_per_step"],2))
PY
done
echo "== cfg2 one step"; timeout 1500 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_cfg2_1step.json 2> gpurun_out/bench_cfg2_1step.err; grep "^\[bench\]" gpurun_out/bench_cfg2_1step.err; tail -c 2500 gpurun_out/bench_cfg2_1step.json
echo "== ingest laps (cfg2, 32 threads)"
ARB_TRACE=1 timeout 600 python - > gpurun_out/ingest_laps_cfg2.txt 2>&1 <<'PY'
import sys, time, glob, os
sys.path.insert(0, ".")
from arriba_b200 import lib
import bench
prefix = bench.ensure_world("cfg2_10M_2x101_50k")
p = lib.Pipeline(prefix + ".bam", prefix + ".gtf", prefix + ".fa", threads=32)
for s in range(lib.STEP_UPLOAD + 1):
    t0 = time.time(); p.step(s); print(lib.STEP_NAMES[s], round(time.time() - t0, 2), flush=True)
PY
grep -v "^WARNING" gpurun_out/ingest_laps_cfg2.txt | tail -16
echo "== ncu launch list (mid)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_mid.csv python bench.py --workload mid_1M_2x101_5k --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -1 gpurun_out/ncu_bench.log | cut -c1-200
echo "== ncu full capture: cascade + re-alignment kernels"
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:cascade_head_fn|cascade_sequences_fn|mismap_item_fn|mismap_heavy_fn|mismap_task_fn|homolog_pairs_fn' -c 7 -o gpurun_out/prof_r01d -f python bench.py --workload mid_1M_2x101_5k --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-200
ls -la gpurun_out | head -50
