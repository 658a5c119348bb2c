#!/bin/bash
# r02q: writer through write() instead of a shared mapping, row order of the fusions file with precomputed keys; cfg5 line
set -u
D=gpurun_out/r02q; mkdir -p $D
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $D/pytest_gpu.txt
echo "== bench cfg2 N=1"; ARB_TRACE=1 timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > $D/bench_cfg2.json 2> $D/bench_cfg2.err; echo "rc=$?"; tail -4 $D/bench_cfg2.err
grep "^\[laps\]\|^\[ingest\]" /tmp/arb_bench/cfg2_10M_2x101_50k/out_rank0/library_stderr.log | tail -100 > $D/host_stage_laps_cfg2.txt
grep "output" $D/host_stage_laps_cfg2.txt | tail -11
echo "== bench cfg2 N=1, shared mapping"; ARB_WRITER_MMAP=1 ARB_TRACE=1 timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $D/bench_cfg2_mmap.json 2> $D/bench_cfg2_mmap.err; tail -1 $D/bench_cfg2_mmap.err
grep "output.*written" /tmp/arb_bench/cfg2_10M_2x101_50k/out_rank0/library_stderr.log | tail -2
echo "== bench cfg5 N=1"; timeout 1200 python bench.py --workload cfg5_10M_mismapper --steps 2 --warmup 1 --no-cpu-baseline > $D/bench_cfg5.json 2> $D/bench_cfg5.err; echo "rc=$?"; tail -2 $D/bench_cfg5.err
python - <<'P'
import json
for w in ('cfg2','cfg5'):
    try:
        l=json.loads(open('gpurun_out/r02q/bench_%s.json' % w).read().strip().splitlines()[-1])
        print(w, 'e2e', round(l['e2e']['seconds_per_step'],3), 'value', round(l['value']), 'parity', l['parity_md5_ok'], 'out', l['e2e']['output_seconds'], 'ingest', l['e2e']['host_seconds']['ingest'])
        print(' ', {k: round(v,1) for k,v in l['roofline']['device_ms'].items()})
    except Exception as e: print(w, 'failed', e)
P
