#!/bin/bash
# r02v: huge pages for the ingest tables (A/B), pass B in one walk with stretches up to 8 GB
set -u
D=gpurun_out/r02v; mkdir -p $D
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $D/pytest_gpu.txt; cat /sys/kernel/mm/transparent_hugepage/enabled
for LA in 1 0; do export ARB_HUGE_PAGES=$LA
  echo "== bench cfg2 N=1, huge pages $LA"; ARB_TRACE=1 timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $D/bench_cfg2_la$LA.json 2> $D/bench_cfg2_la$LA.err; echo "rc=$?"; grep "^\[bench\]" $D/bench_cfg2_la$LA.err | tail -3
  grep "^\[laps\]\|^\[ingest\]" /tmp/arb_bench/cfg2_10M_2x101_50k/out_rank0/library_stderr.log | tail -100 > $D/host_stage_laps_cfg2_la$LA.txt
  grep "of which\|inflate + scan\|normalise\|SoA\|sort by" $D/host_stage_laps_cfg2_la$LA.txt | tail -10
done
python - <<'P'
import json
for la in (1, 0):
    l=json.loads(open('gpurun_out/r02v/bench_cfg2_la%d.json' % la).read().strip().splitlines()[-1])
    print('huge pages', la, 'e2e', round(l['e2e']['seconds_per_step'],3), 'value', round(l['value']), 'parity', l['parity_md5_ok'], 'out', l['e2e']['output_seconds'], 'ingest', l['e2e']['host_seconds']['ingest'], 'find_fusions', round(l['roofline']['device_ms']['find_fusions_total'],1))
    print(' ', sorted(l['e2e']['event_seconds'].items(), key=lambda kv: -kv[1])[:8])
P
