#!/bin/bash
# Short validation: parity tests and one bench line on the small workload. Outputs -> gpurun_out/
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== bench mid"; timeout 300 python bench.py --workload mid_1M_2x101_5k --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mid_check.json 2> gpurun_out/bench_mid_check.err; grep "^\[bench\]" gpurun_out/bench_mid_check.err | tail -1; tail -c 1500 gpurun_out/bench_mid_check.json
