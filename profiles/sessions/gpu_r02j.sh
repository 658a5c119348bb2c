#!/bin/bash
# r02j (2 GPUs): event stages on the device + genome upload with the reference; per-phase times of the sharded run; ncu of the re-alignment kernel
set -u
D=gpurun_out/r02j; mkdir -p $D
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $D/pytest_gpu.txt
echo "== bench cfg2 N=1"; ARB_TRACE=1 timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $D/bench_cfg2.json 2> $D/bench_cfg2.err; echo "rc=$?"; tail -3 $D/bench_cfg2.err
grep "^\[laps\]\|^\[ingest\]" /tmp/arb_bench/cfg2_10M_2x101_50k/out_rank0/library_stderr.log | tail -80 > $D/host_stage_laps_cfg2.txt
python - <<'P'
import json
l=json.loads(open('gpurun_out/r02j/bench_cfg2.json').read().strip().splitlines()[-1])
print('e2e', l['e2e']['seconds_per_step'], 'parity', l['parity_md5_ok'], 'host', l['e2e']['host_seconds'], 'out', l['e2e']['output_seconds'])
print({k:v for k,v in l['e2e']['event_seconds'].items() if v>=0.03})
print(l['roofline']['device_ms'])
P
echo "== bench cfg2 N=2 (trace)"; ARB_TRACE=1 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 2 --warmup 1 --no-secondary > $D/bench_cfg2_n2.json 2> $D/bench_cfg2_n2.err; echo "rc=$?"; grep "bench\]" $D/bench_cfg2_n2.err | tail -3 | cut -c1-300
grep "laps\] sharded" /tmp/arb_bench/cfg2_10M_2x101_50k/out_rank0/library_stderr.log | tail -12 | tee $D/sharded_laps_n2.txt
python - <<'P'
import json
try:
    l=json.loads(open('gpurun_out/r02j/bench_cfg2_n2.json').read().strip().splitlines()[-1])
    print(2, 'value', l['value'], 'e2e', l['e2e']['seconds_per_step'], 'parity', l['parity_md5_ok'])
except Exception as e: print('N=2 failed', e)
P
echo "== ncu full: re-alignment pass 1"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_mismap_items" -c 1 -o $D/prof_mismap python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-parity > $D/ncu_bench.log 2>&1; echo "ncu rc=$?"
ls -la $D | head -20
