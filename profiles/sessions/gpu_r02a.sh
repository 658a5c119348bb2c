#!/bin/bash
# r02a: first session of round 2 -- GPU tests on the round-2 tree, default bench line with the md5 parity check of the full-size CUDA run.
set -u
mkdir -p gpurun_out/r02a
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > gpurun_out/r02a/gpu.txt
cat /sys/fs/cgroup/cpu.max > gpurun_out/r02a/cpu_max.txt; nproc >> gpurun_out/r02a/cpu_max.txt
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r02a/pytest_gpu.txt
echo "== bench cfg2 (1 warm-up + 2 timed steps, md5 parity)"; ARB_TRACE=1 timeout 900 python bench.py --steps 2 --warmup 1 > gpurun_out/r02a/bench_cfg2.json 2> gpurun_out/r02a/bench_cfg2.err; echo "rc=$?"
tail -c 1500 gpurun_out/r02a/bench_cfg2.json; tail -5 gpurun_out/r02a/bench_cfg2.err
cp /tmp/arb_bench/cfg2_10M_2x101_50k/out_rank0/library_stderr.log gpurun_out/r02a/library_stderr_cfg2.log 2>/dev/null; grep -v "WARNING" gpurun_out/r02a/library_stderr_cfg2.log | grep "^\[laps\]\|^\[ingest\]" | tail -70 > gpurun_out/r02a/host_stage_laps_cfg2.txt; rm -f gpurun_out/r02a/library_stderr_cfg2.log
