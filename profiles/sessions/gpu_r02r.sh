#!/bin/bash
# r02r: three predicates + e-value tallies on the device, writev; launch list of one cfg2 step; ncu --set full (with source) of the consensus kernel and of pass 1 of the re-alignment
set -u
D=gpurun_out/r02r; mkdir -p $D
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $D/pytest_gpu.txt
echo "== bench cfg2 N=1"; ARB_TRACE=1 timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > $D/bench_cfg2.json 2> $D/bench_cfg2.err; echo "rc=$?"; tail -3 $D/bench_cfg2.err
grep "^\[laps\]\|^\[ingest\]" /tmp/arb_bench/cfg2_10M_2x101_50k/out_rank0/library_stderr.log | tail -100 > $D/host_stage_laps_cfg2.txt
grep "output" $D/host_stage_laps_cfg2.txt | tail -11
python - <<'P'
import json
l=json.loads(open('gpurun_out/r02r/bench_cfg2.json').read().strip().splitlines()[-1])
print('e2e', round(l['e2e']['seconds_per_step'],3), 'value', round(l['value']), 'parity', l['parity_md5_ok'], 'out', l['e2e']['output_seconds'], l['e2e']['host_seconds'])
print(sorted(l['e2e']['event_seconds'].items(), key=lambda kv: -kv[1])[:12])
P
echo "== ncu launch list (cfg2, one step)"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $D/launches_cfg2.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-parity > $D/ncu_launches.log 2>&1; echo "rc=$?"
python tools/ncu_summary.py launches $D/launches_cfg2.csv > $D/launches_cfg2.txt; head -24 $D/launches_cfg2.txt; rm -f $D/launches_cfg2.csv
echo "== ncu full: consensus + re-alignment pass 1"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_consensus|k_mismap_items" -c 2 -o $D/prof_cons_mismap python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-parity > $D/ncu_bench.log 2>&1; echo "ncu rc=$?"
ls -la $D
