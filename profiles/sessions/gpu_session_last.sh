#!/bin/bash
# Last session of the round: parity tests on the final tree, then the default workload (cfg2) with fewer timed steps.
# Outputs -> gpurun_out/
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== bench cfg2 (3 warm-up + 2 timed steps)"; ARB_TRACE=1 timeout 215 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg2_last.json 2> gpurun_out/bench_cfg2_last.err; grep "^\[bench\]" gpurun_out/bench_cfg2_last.err | tail -6
grep -v "^Filtering\|WARNING" gpurun_out/bench_cfg2_last.err | grep "^\[laps\]\|^\[ingest\]" | tail -60 > gpurun_out/host_stage_laps_cfg2_last.txt
tail -c 600 gpurun_out/bench_cfg2_last.json
