#!/bin/bash
# r02p: consensus stage with page-locked result buffers and finer laps; how fast a gigabyte of rows reaches the page cache on the box (tools/writebench.cpp)
set -u
D=gpurun_out/r02p; mkdir -p $D
echo "== pytest -m gpu -k e2e"; timeout 600 python -m pytest tests -m gpu -x -q -k e2e 2>&1 | tail -3 | tee $D/pytest_gpu_e2e.txt
echo "== bench cfg2 N=1"; ARB_TRACE=1 timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $D/bench_cfg2.json 2> $D/bench_cfg2.err; echo "rc=$?"; tail -3 $D/bench_cfg2.err
grep "^\[laps\]\|^\[ingest\]" /tmp/arb_bench/cfg2_10M_2x101_50k/out_rank0/library_stderr.log | tail -90 > $D/host_stage_laps_cfg2.txt
grep "output" $D/host_stage_laps_cfg2.txt | tail -12
echo "== writebench"; build/writebench /tmp/arb_bench 760 16 | tee $D/writebench.txt; build/writebench /tmp/arb_bench 760 32 | tail -10 | tee -a $D/writebench.txt
nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null; df -h /tmp | tail -1
