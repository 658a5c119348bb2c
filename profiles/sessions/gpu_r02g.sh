#!/bin/bash
# r02g (2 GPUs): NCCL test of the multi-GPU protocol, bench cfg2 at N=1 and N=2 on the same box
set -u
D=gpurun_out/r02g; mkdir -p $D
nvidia-smi --query-gpu=index,name --format=csv > $D/gpus.txt; nproc >> $D/gpus.txt; cat /sys/fs/cgroup/cpu.max >> $D/gpus.txt
echo "== pytest sharded (nccl)"; timeout 900 python -m pytest tests/test_sharded.py -m gpu -x -q 2>&1 | tail -5 | tee $D/pytest_sharded.txt
echo "== bench cfg2 N=1"; timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $D/bench_cfg2_n1.json 2> $D/bench_cfg2_n1.err; echo "rc=$?"; tail -3 $D/bench_cfg2_n1.err
echo "== bench cfg2 N=2"; NCCL_DEBUG=INFO timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 2 --warmup 1 > $D/bench_cfg2_n2.json 2> $D/bench_cfg2_n2.err; echo "rc=$?"; grep -v "NCCL INFO" $D/bench_cfg2_n2.err | tail -12; grep -c "NCCL INFO" $D/bench_cfg2_n2.err; grep "nranks" $D/bench_cfg2_n2.err | head -3
python - <<'P'
import json
for n in (1,2):
    try:
        l=json.loads(open('gpurun_out/r02g/bench_cfg2_n%d.json'%n).read().strip().splitlines()[-1])
        print(n, 'value', l['value'], 'e2e', l['e2e']['seconds_per_step'], 'parity', l['parity_md5_ok'], 'host', l['e2e']['host_seconds'], l.get('secondary_mode'))
        print(l['roofline']['device_ms'])
    except Exception as e: print(n, 'failed', e)
P
