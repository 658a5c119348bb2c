#!/bin/bash
# r02k: iteration order on the device, packed comparisons in the re-alignment's extensions
set -u
D=gpurun_out/r02k; mkdir -p $D
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $D/pytest_gpu.txt
echo "== bench cfg2 N=1"; ARB_TRACE=1 timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > $D/bench_cfg2.json 2> $D/bench_cfg2.err; echo "rc=$?"; tail -3 $D/bench_cfg2.err
grep "^\[laps\]\|^\[ingest\]" /tmp/arb_bench/cfg2_10M_2x101_50k/out_rank0/library_stderr.log | tail -80 > $D/host_stage_laps_cfg2.txt
python - <<'P'
import json
l=json.loads(open('gpurun_out/r02k/bench_cfg2.json').read().strip().splitlines()[-1])
print('e2e', l['e2e']['seconds_per_step'], 'parity', l['parity_md5_ok'], 'host', l['e2e']['host_seconds'], 'out', l['e2e']['output_seconds'])
print({k:v for k,v in l['e2e']['event_seconds'].items() if v>=0.03})
print(l['roofline']['device_ms'])
P
echo "== ncu launch list (cfg2, one step)"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $D/launches_cfg2.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-parity > $D/ncu_launches.log 2>&1; echo "rc=$?"
python tools/ncu_summary.py launches $D/launches_cfg2.csv > $D/launches_cfg2.txt; head -24 $D/launches_cfg2.txt
echo "== ncu full: re-alignment pass 1, pass B, annotation pass 1"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_mismap_items|k_walk_b|annotate_pass1" -c 4 -o $D/prof_kernels python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-parity > $D/ncu_bench.log 2>&1; echo "ncu rc=$?"
rm -f $D/launches_cfg2.csv
