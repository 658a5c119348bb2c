#!/bin/bash
# GPU session: parity tests, re-alignment with the continuation registry (budget sweep mid, cfg2 one step), ingest laps. Outputs -> gpurun_out/
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
summ() { python - "$1" <<'PY'
import json, sys
lines = [l for l in open(sys.argv[1]) if l.startswith("{")]
d = json.loads(lines[-1]); r = d["roofline"]; e = d["e2e"]
print(sys.argv[1].split("/")[-1], {k: round(v, 1) for k, v in r["device_ms"].items() if k.startswith("mism") or k == "homologs"}, "heavy", r.get("mismapper_heavy_items"), "tasks", r.get("mismapper_tasks"), "rounds", r.get("mismapper_rounds"),
      "| e2e s", round(e["seconds_per_step"], 2), "ingest", e["host_seconds"]["ingest"], "output", e["output_seconds"], "value", round(d["value"]))
PY
}
echo "== budget / lanes sweep (mid): budget lanes spawn task_lanes"
for cfg in "4096 1024 0 32" "1024 1024 0 32" "512 256 0 32" "1024 256 0 64" "4096 1024 512 32"; do
  set -- $cfg
  ARB_MISMAP_BUDGET=$1 ARB_MISMAP_LANES=$2 ARB_MISMAP_SPAWN=$3 ARB_MISMAP_TASK_LANES=$4 timeout 600 python bench.py --workload mid_1M_2x101_5k --steps 2 --warmup 1 --no-cpu-baseline 2> /dev/null > gpurun_out/reg_mid_$1_$2_$3_$4.json; summ gpurun_out/reg_mid_$1_$2_$3_$4.json
done
echo "== cfg2, one step: budget lanes"
for cfg in "4096 1024" "1024 256"; do
  set -- $cfg
  ARB_MISMAP_BUDGET=$1 ARB_MISMAP_LANES=$2 timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2> gpurun_out/reg_cfg2_$1_$2.err > gpurun_out/reg_cfg2_$1_$2.json; grep "^\[bench\]" gpurun_out/reg_cfg2_$1_$2.err | tail -1; summ gpurun_out/reg_cfg2_$1_$2.json
done
echo "== ingest laps (cfg2, 64 threads)"
ARB_TRACE=1 timeout 300 python - > gpurun_out/ingest_laps_cfg2_64.txt 2>&1 <<'PY'
import sys, time
sys.path.insert(0, ".")
from arriba_b200 import lib
import bench
prefix = bench.ensure_world("cfg2_10M_2x101_50k")
p = lib.Pipeline(prefix + ".bam", prefix + ".gtf", prefix + ".fa", threads=64)
for s in range(lib.STEP_ANNOTATE + 1):
    t0 = time.time(); p.step(s); print(lib.STEP_NAMES[s], round(time.time() - t0, 2), flush=True)
PY
grep "ingest\]\|^ingest\|^annotate" gpurun_out/ingest_laps_cfg2_64.txt
ls gpurun_out | head -60
