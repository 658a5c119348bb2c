#!/bin/bash
# r02d: warp-per-candidate pass B of find_fusions
set -u
D=gpurun_out/r02d; mkdir -p $D
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $D/pytest_gpu.txt
echo "== bench cfg2"; ARB_TRACE=1 timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > $D/bench_cfg2.json 2> $D/bench_cfg2.err; echo "rc=$?"; tail -3 $D/bench_cfg2.err
python - <<'P'
import json
l=json.loads(open('gpurun_out/r02d/bench_cfg2.json').read().strip().splitlines()[-1])
print('e2e', l['e2e']['seconds_per_step'], 'parity', l['parity_md5_ok'], 'host', l['e2e']['host_seconds'])
print(l['roofline']['device_ms'])
P
