#!/bin/bash
# Final session of the round: smoke, parity tests, default bench (both arms), bench mid, launch list. Outputs -> gpurun_out/
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== default bench, reference arm then ours"
timeout 900 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 300 gpurun_out/bench_reference.json
ARB_TRACE=1 timeout 1500 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; grep "^\[bench\]" gpurun_out/bench_default.err | tail -3; tail -c 3400 gpurun_out/bench_default.json
grep "^\[laps\]\|^\[ingest\]" gpurun_out/bench_default.err | tail -36 > gpurun_out/host_stage_laps_cfg2.txt
echo "== bench mid"; timeout 600 python bench.py --workload mid_1M_2x101_5k > gpurun_out/bench_mid.json 2> gpurun_out/bench_mid.err; grep "^\[bench\]" gpurun_out/bench_mid.err | tail -1
echo "== ncu launch list (mid)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_mid.csv python bench.py --workload mid_1M_2x101_5k --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -1 gpurun_out/ncu_bench.log | cut -c1-200
