#!/bin/bash
# GPU session: parity tests, occupancy sweep of the re-alignment kernels (mid + cfg2), default bench (both arms). Outputs -> gpurun_out/
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
summ() { python - "$1" <<'PY'
import json, sys
d = json.load(open(sys.argv[1])); r = d["roofline"]; e = d["e2e"]
print(sys.argv[1].split("/")[-1], {k: round(v, 1) for k, v in r["device_ms"].items() if k.startswith("mism") or k == "homologs"}, "heavy", r.get("mismapper_heavy_items"), "tasks", r.get("mismapper_tasks"), "rounds", r.get("mismapper_rounds"),
      "| e2e s", round(e["seconds_per_step"], 2), "ingest", e["host_seconds"]["ingest"], e["ingest_split"], "output", e["output_seconds"], "value", round(d["value"]))
PY
}
echo "== occupancy sweep (mid)"
for occ in 2 3 4; do
  ARB_MISMAP_OCC=$occ timeout 600 python bench.py --workload mid_1M_2x101_5k --steps 2 --warmup 1 --no-cpu-baseline 2> /dev/null > gpurun_out/occ_mid_$occ.json; summ gpurun_out/occ_mid_$occ.json
done
echo "== occupancy / spawn sweep (cfg2, one step)"
for cfg in "2 2048" "4 2048" "4 512"; do
  set -- $cfg
  ARB_MISMAP_OCC=$1 ARB_MISMAP_SPAWN=$2 timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2> gpurun_out/occ_cfg2_$1_$2.err > gpurun_out/occ_cfg2_$1_$2.json; grep "^\[bench\]" gpurun_out/occ_cfg2_$1_$2.err | tail -1; summ gpurun_out/occ_cfg2_$1_$2.json
done
echo "== default bench, reference arm then ours"
timeout 900 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 1200 gpurun_out/bench_reference.json
timeout 1500 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; grep "^\[bench\]" gpurun_out/bench_default.err | tail -3; tail -c 3000 gpurun_out/bench_default.json
ls -la gpurun_out | head -40
