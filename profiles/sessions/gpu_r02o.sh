#!/bin/bash
# r02o: pileups + consensus of the fusions rows on the device (consensus.cu)
set -u
D=gpurun_out/r02o; mkdir -p $D
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $D/pytest_gpu.txt
echo "== bench cfg2 N=1"; ARB_TRACE=1 timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > $D/bench_cfg2.json 2> $D/bench_cfg2.err; echo "rc=$?"; tail -3 $D/bench_cfg2.err
grep "^\[laps\]\|^\[ingest\]" /tmp/arb_bench/cfg2_10M_2x101_50k/out_rank0/library_stderr.log | tail -90 > $D/host_stage_laps_cfg2.txt
python - <<'P'
import json
l=json.loads(open('gpurun_out/r02o/bench_cfg2.json').read().strip().splitlines()[-1])
print('e2e', l['e2e']['seconds_per_step'], 'parity', l['parity_md5_ok'], 'host', l['e2e']['host_seconds'], 'out', l['e2e']['output_seconds'])
print({k:v for k,v in l['e2e']['event_seconds'].items() if v>=0.03})
print(l['roofline']['device_ms'])
for k in l['roofline']['kernels']: print(k['kernel'][:44], round(k['kernel_ms'],2), 'ms frac', round(k['frac'],4), 'alg', k['algorithmic_bytes_per_launch'], 'traffic', k['traffic'])
P
grep "output" $D/host_stage_laps_cfg2.txt | tail -16
