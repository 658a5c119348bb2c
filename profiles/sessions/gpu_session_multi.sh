#!/bin/bash
# N-GPU session (gpurun --gpus N): NCCL parity test of the sharded run, bench in both multi-GPU modes. Outputs -> gpurun_out/
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_multi.txt 2>&1
echo "== sharded parity on $N GPUs (NCCL)"; timeout 900 python -m pytest tests/test_sharded.py -m gpu -x -q 2>&1 | tail -4
for mode in samples sharded; do
  echo "== bench mid, $N GPUs, mode $mode"
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --workload mid_1M_2x101_5k --steps 2 --warmup 1 --mode $mode --no-cpu-baseline > gpurun_out/bench_mid_${N}gpu_$mode.json 2> gpurun_out/bench_mid_${N}gpu_$mode.err
  grep "^\[bench\]" gpurun_out/bench_mid_${N}gpu_$mode.err | tail -2; tail -c 1500 gpurun_out/bench_mid_${N}gpu_$mode.json; tail -3 gpurun_out/bench_mid_${N}gpu_$mode.err | cut -c1-300
done
echo "== single GPU reference point (same box)"
timeout 600 python bench.py --workload mid_1M_2x101_5k --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_mid_1gpu_samebox.json 2> /dev/null; tail -c 600 gpurun_out/bench_mid_1gpu_samebox.json
