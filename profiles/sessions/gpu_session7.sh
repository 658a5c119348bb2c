#!/bin/bash
# GPU session: parity tests, re-alignment registry on mid + cfg2 (one step each), stage laps on the 128-core host. Outputs -> gpurun_out/
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
summ() { python - "$1" <<'PY'
import json, sys
lines = [l for l in open(sys.argv[1]) if l.startswith("{")]
d = json.loads(lines[-1]); r = d["roofline"]; e = d["e2e"]
print(sys.argv[1].split("/")[-1], {k: round(v, 1) for k, v in r["device_ms"].items() if k.startswith("mism") or k == "homologs"}, "heavy", r.get("mismapper_heavy_items"), "tasks", r.get("mismapper_tasks"), "rounds", r.get("mismapper_rounds"), r.get("mismapper_registry"),
      "| e2e s", round(e["seconds_per_step"], 2), "ingest", e["host_seconds"]["ingest"], "output", e["output_seconds"], "value", round(d["value"]))
PY
}
echo "== mid"; timeout 600 python bench.py --workload mid_1M_2x101_5k --steps 2 --warmup 1 --no-cpu-baseline 2> /dev/null > gpurun_out/reg2_mid.json; summ gpurun_out/reg2_mid.json
echo "== cfg2, one step"
ARB_TRACE=1 timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2> gpurun_out/reg2_cfg2.err > gpurun_out/reg2_cfg2.json; grep "^\[bench\]" gpurun_out/reg2_cfg2.err | tail -1; summ gpurun_out/reg2_cfg2.json
grep "^\[laps\]\|^\[ingest\]" gpurun_out/reg2_cfg2.err | tail -28
echo "== cfg2, one step, budget 1024"
ARB_MISMAP_BUDGET=1024 timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2> /dev/null > gpurun_out/reg2_cfg2_b1024.json; summ gpurun_out/reg2_cfg2_b1024.json
