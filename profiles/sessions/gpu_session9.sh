#!/bin/bash
# GPU session: how do the host stages scale with the thread count on this box? Outputs -> gpurun_out/
set -u
mkdir -p gpurun_out
{ echo "nproc $(nproc)"; echo "cpu.max $(cat /sys/fs/cgroup/cpu.max 2>/dev/null)"; cat /sys/fs/cgroup/cpu.stat 2>/dev/null | head -6; lscpu | grep -E "Model name|Socket|Core|Thread|NUMA|MHz" ; free -g | head -2; } > gpurun_out/host_info.txt 2>&1
cat gpurun_out/host_info.txt
for th in 16 32 64 96 128; do
  timeout 600 python bench.py --threads $th --steps 1 --warmup 1 --no-cpu-baseline 2> gpurun_out/threads_$th.err > gpurun_out/threads_$th.json
  echo "threads $th: $(grep '^\[bench\]' gpurun_out/threads_$th.err | tail -1)"
  python - gpurun_out/threads_$th.json <<'PY'
import json, sys
lines = [l for l in open(sys.argv[1]) if l.startswith("{")]
d = json.loads(lines[-1]); e = d["e2e"]
print("   ", {k: v for k, v in e["event_seconds"].items() if v >= 0.2}, e["ingest_split"])
PY
done
cat /sys/fs/cgroup/cpu.stat 2>/dev/null | head -6
