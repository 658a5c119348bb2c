#!/bin/bash
# r02t: select_best on the device, discarded block by four pwrite threads, one-walk pass B, pass 1 without spills
set -u
D=gpurun_out/r02t; mkdir -p $D
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $D/pytest_gpu.txt
echo "== bench cfg2 N=1"; ARB_TRACE=1 timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > $D/bench_cfg2.json 2> $D/bench_cfg2.err; echo "rc=$?"; tail -3 $D/bench_cfg2.err
grep "^\[laps\]\|^\[ingest\]" /tmp/arb_bench/cfg2_10M_2x101_50k/out_rank0/library_stderr.log | tail -100 > $D/host_stage_laps_cfg2.txt
grep "output" $D/host_stage_laps_cfg2.txt | tail -11
python - <<'P'
import json
l=json.loads(open('gpurun_out/r02t/bench_cfg2.json').read().strip().splitlines()[-1])
print('e2e', round(l['e2e']['seconds_per_step'],3), 'value', round(l['value']), 'parity', l['parity_md5_ok'], 'out', l['e2e']['output_seconds'], l['e2e']['host_seconds'])
print(sorted(l['e2e']['event_seconds'].items(), key=lambda kv: -kv[1])[:12])
P
