#!/bin/bash
# r02b: lane-group sequence kernel (8 lanes per fragment, bulk-copy staging): GPU parity tests, bench cfg2, ncu of both cascade kernels on cfg2
set -u
D=gpurun_out/r02b; mkdir -p $D
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $D/pytest_gpu.txt
echo "== bench cfg2"; timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $D/bench_cfg2.json 2> $D/bench_cfg2.err; echo "rc=$?"
python - <<'P'
import json
l=json.loads(open('gpurun_out/r02b/bench_cfg2.json').read().strip().splitlines()[-1])
print('e2e', l['e2e']['seconds_per_step'], 'parity', l['parity_md5_ok'])
for k in l['roofline']['kernels']: print(k['kernel'][:40], round(k['kernel_ms'],3), 'ms', round(k['frac'],4))
print(l['roofline']['device_ms'])
P
echo "== ncu cascade kernels (cfg2, one step)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_cascade_sequences|cascade_head_fn" -c 2 -o $D/prof_cascade python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-parity > $D/ncu_bench.log 2>&1; echo "ncu rc=$?"
ls -la $D
