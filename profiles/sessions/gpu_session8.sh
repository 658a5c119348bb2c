#!/bin/bash
# GPU session: parity tests, cfg2 one step with stage laps, default bench (both arms). Outputs -> gpurun_out/
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== cfg2, one step, laps"
ARB_TRACE=1 timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2> gpurun_out/laps_cfg2.err > gpurun_out/laps_cfg2.json; grep "^\[bench\]" gpurun_out/laps_cfg2.err | tail -1
grep "^\[laps\]\|^\[ingest\]" gpurun_out/laps_cfg2.err | tail -24
echo "== default bench, reference arm then ours"
timeout 900 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 300 gpurun_out/bench_reference.json
timeout 1500 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; grep "^\[bench\]" gpurun_out/bench_default.err | tail -3; tail -c 3200 gpurun_out/bench_default.json
echo "== bench mid"; timeout 600 python bench.py --workload mid_1M_2x101_5k > gpurun_out/bench_mid.json 2> gpurun_out/bench_mid.err; grep "^\[bench\]" gpurun_out/bench_mid.err | tail -1
