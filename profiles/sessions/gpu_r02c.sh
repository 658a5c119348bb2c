#!/bin/bash
# r02c: per-warp sequence kernel with the exact parallel 3-mer count, lane-group re-alignment pass 1, page-locked columns + early asynchronous upload
set -u
D=gpurun_out/r02c; mkdir -p $D
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $D/pytest_gpu.txt
echo "== bench cfg2"; ARB_TRACE=1 timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > $D/bench_cfg2.json 2> $D/bench_cfg2.err; echo "rc=$?"; tail -3 $D/bench_cfg2.err
python - <<'P'
import json
l=json.loads(open('gpurun_out/r02c/bench_cfg2.json').read().strip().splitlines()[-1])
print('e2e', l['e2e']['seconds_per_step'], 'parity', l['parity_md5_ok'], 'host', l['e2e']['host_seconds'])
for k in l['roofline']['kernels']: print(k['kernel'][:40], round(k['kernel_ms'],3), 'ms', round(k['frac'],4))
print(l['roofline']['device_ms']); print('heavy', l['roofline']['mismapper_heavy_items'], 'tasks', l['roofline']['mismapper_tasks'])
P
grep "^\[laps\]\|^\[ingest\]" /tmp/arb_bench/cfg2_10M_2x101_50k/out_rank0/library_stderr.log | tail -75 > $D/host_stage_laps_cfg2.txt
echo "== ncu launch list (cfg2, one step)"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $D/launches_cfg2.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-parity > $D/ncu_launches.log 2>&1; echo "rc=$?"
python tools/ncu_summary.py launches $D/launches_cfg2.csv > $D/launches_cfg2.txt; head -30 $D/launches_cfg2.txt
echo "== ncu full: sequences + mismap"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_cascade_sequences|k_mismap_items" -c 2 -o $D/prof_seq_mismap python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-parity > $D/ncu_bench.log 2>&1; echo "ncu rc=$?"
rm -f $D/launches_cfg2.csv.tmp; ls -la $D
