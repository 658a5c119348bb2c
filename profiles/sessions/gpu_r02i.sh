#!/bin/bash
# r02i (2 GPUs): block table of the k-mer index (lanes per item swept), in_vitro / spliced support / multimappers on the device, N=2 bench
set -u
D=gpurun_out/r02i; mkdir -p $D
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $D/pytest_gpu.txt
for L in 8 16 32; do
  echo "== bench cfg2 N=1 lanes=$L"; ARB_MISMAP_GROUP_LANES=$L timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $D/bench_cfg2_lanes$L.json 2> $D/bench_cfg2_lanes$L.err; echo "rc=$?"
  python - $L <<'P'
import json,sys
L=sys.argv[1]
try:
    l=json.loads(open('gpurun_out/r02i/bench_cfg2_lanes%s.json'%L).read().strip().splitlines()[-1])
    d=l['roofline']['device_ms']; print('lanes',L,'e2e',round(l['e2e']['seconds_per_step'],2),'parity',l['parity_md5_ok'],'pass1',round(d['mismappers_pass1'],1),'pass2',round(d['mismappers_pass2'],1),'heavy',l['roofline']['mismapper_heavy_items'],'tasks',l['roofline']['mismapper_tasks'])
except Exception as e: print('failed',e); print(open('gpurun_out/r02i/bench_cfg2_lanes%s.err'%L).read()[-1500:])
P
done
echo "== bench cfg2 N=2"; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 2 --warmup 1 > $D/bench_cfg2_n2.json 2> $D/bench_cfg2_n2.err; echo "rc=$?"; grep -v "NCCL INFO" $D/bench_cfg2_n2.err | tail -8 | cut -c1-400
python - <<'P'
import json
try:
    l=json.loads(open('gpurun_out/r02i/bench_cfg2_n2.json').read().strip().splitlines()[-1])
    print(2, 'value', l['value'], 'e2e', l['e2e']['seconds_per_step'], 'parity', l['parity_md5_ok'], 'host', l['e2e']['host_seconds'], l.get('secondary_mode'))
    print(l['roofline']['device_ms'])
except Exception as e: print('N=2 failed', e)
P
