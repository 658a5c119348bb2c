#!/bin/bash
# r02s (4 GPUs): GPU tests incl. the NCCL variant of the sharded run; bench cfg2 at N=4 and N=2 (ONE sample over the ranks), N=1 on the same box
set -u
D=gpurun_out/r02s; mkdir -p $D
nvidia-smi -L | head -8; cat /sys/fs/cgroup/cpu.max
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $D/pytest_gpu.txt
for N in 4 2; do
  echo "== bench cfg2 N=$N"
  ARB_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29510+N)) bench.py --gpus $N --steps 2 --warmup 1 --no-secondary > $D/bench_cfg2_n$N.json 2> $D/bench_cfg2_n$N.err; echo "rc=$?"
  grep "^\[bench\]" $D/bench_cfg2_n$N.err | tail -3
  grep "^\[laps\] sharded" /tmp/arb_bench/cfg2_10M_2x101_50k/out_rank0/library_stderr.log | tail -12 > $D/sharded_laps_n$N.txt; cat $D/sharded_laps_n$N.txt
  tail -c 1500 $D/bench_cfg2_n$N.err | grep -i "nccl\|error\|Traceback" | head -5
done
echo "== bench cfg2 N=1"; timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $D/bench_cfg2_n1.json 2> $D/bench_cfg2_n1.err; echo "rc=$?"; grep "^\[bench\]" $D/bench_cfg2_n1.err | tail -2
python - <<'P'
import json
for n in (1, 2, 4):
    try:
        l=json.loads(open('gpurun_out/r02s/bench_cfg2_n%d.json' % n).read().strip().splitlines()[-1])
        print('N=%d' % n, 'e2e', round(l['e2e']['seconds_per_step'],3), 'value', round(l['value']), 'parity', l.get('parity_md5_ok'), 'scaling', l['scaling'], 'threads', l['config'].get('host_threads_rank0_sharded'), 'cpus', l['config'].get('host_cpus_usable'))
    except Exception as e: print('N=%d failed' % n, e)
P
