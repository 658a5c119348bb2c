#!/bin/bash
# r02u: pass B in one walk (variable stretches), record loop with look-ahead of the fragment table walk (A/B), spliced / confidence loops on all threads
set -u
D=gpurun_out/r02u; mkdir -p $D
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $D/pytest_gpu.txt
for LA in 1 0; do
  echo "== bench cfg2 N=1, look-ahead $LA"; ARB_INGEST_LOOKAHEAD=$LA ARB_TRACE=1 timeout 900 python bench.py --steps 3 --warmup 1 --no-cpu-baseline > $D/bench_cfg2_la$LA.json 2> $D/bench_cfg2_la$LA.err; echo "rc=$?"; grep "^\[bench\]" $D/bench_cfg2_la$LA.err | tail -3
  grep "^\[laps\]\|^\[ingest\]" /tmp/arb_bench/cfg2_10M_2x101_50k/out_rank0/library_stderr.log | tail -100 > $D/host_stage_laps_cfg2_la$LA.txt
  grep "of which\|inflate + scan" $D/host_stage_laps_cfg2_la$LA.txt | tail -4
done
python - <<'P'
import json
for la in (1, 0):
    l=json.loads(open('gpurun_out/r02u/bench_cfg2_la%d.json' % la).read().strip().splitlines()[-1])
    print('look-ahead', la, 'e2e', round(l['e2e']['seconds_per_step'],3), 'value', round(l['value']), 'parity', l['parity_md5_ok'], 'out', l['e2e']['output_seconds'], 'ingest', l['e2e']['host_seconds']['ingest'], 'find_fusions', round(l['roofline']['device_ms']['find_fusions_total'],1))
    print(' ', sorted(l['e2e']['event_seconds'].items(), key=lambda kv: -kv[1])[:8])
P
