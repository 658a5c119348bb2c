#!/bin/bash
# GPU session: smoke, parity tests, default bench (both arms), ingest thread scaling, launch list + ncu captures of the cascade at cfg2. Outputs -> gpurun_out/
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== default bench, reference arm then ours"
timeout 900 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 400 gpurun_out/bench_reference.json
timeout 1500 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; grep "^\[bench\]" gpurun_out/bench_default.err | tail -3; tail -c 2600 gpurun_out/bench_default.json
echo "== bench mid"; timeout 600 python bench.py --workload mid_1M_2x101_5k > gpurun_out/bench_mid.json 2> gpurun_out/bench_mid.err; grep "^\[bench\]" gpurun_out/bench_mid.err | tail -1
echo "== ingest laps (cfg2) by thread count"
for th in 16 32 64 128; do
ARB_TRACE=1 timeout 300 python - $th > gpurun_out/ingest_laps_cfg2_$th.txt 2>&1 <<'PY'
import sys, time
sys.path.insert(0, ".")
from arriba_b200 import lib
import bench
prefix = bench.ensure_world("cfg2_10M_2x101_50k")
p = lib.Pipeline(prefix + ".bam", prefix + ".gtf", prefix + ".fa", threads=int(sys.argv[1]))
for s in range(lib.STEP_ANNOTATE + 1):
    t0 = time.time(); p.step(s); print(lib.STEP_NAMES[s], round(time.time() - t0, 2), flush=True)
PY
echo "threads $th"; grep "ingest\]\|^ingest\|^annotate" gpurun_out/ingest_laps_cfg2_$th.txt | tr '\n' ';' | cut -c1-900; echo
done
echo "== ncu launch list (mid)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_mid.csv python bench.py --workload mid_1M_2x101_5k --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -1 gpurun_out/ncu_bench.log | cut -c1-200
echo "== ncu full capture of the cascade kernels on the default workload (traffic)"
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:cascade_head_fn|cascade_sequences_fn' -c 2 -o gpurun_out/prof_cascade_cfg2 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full_cfg2.log 2>&1
tail -2 gpurun_out/ncu_full_cfg2.log | cut -c1-200
ls -la gpurun_out | head -50
