#!/bin/bash
# r02w: the tree at the end of round 2: smoke, GPU tests, the default bench line exactly as the driver runs it, the reference arm, the mismapper-heavy workload
set -u
D=gpurun_out/r02w; mkdir -p $D
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $D/smoke.txt
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $D/pytest_gpu.txt
echo "== python bench.py (defaults)"; ARB_TRACE=1 timeout 1200 python bench.py > $D/bench_default.json 2> $D/bench_default.err; echo "rc=$?"; grep "^\[bench\]" $D/bench_default.err | tail -3
grep "^\[laps\]\|^\[ingest\]" /tmp/arb_bench/cfg2_10M_2x101_50k/out_rank0/library_stderr.log | tail -100 > $D/host_stage_laps_cfg2.txt
echo "== python bench.py --impl reference"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $D/bench_reference.json 2> $D/bench_reference.err; echo "rc=$?"
echo "== bench cfg5"; timeout 1200 python bench.py --workload cfg5_10M_mismapper --steps 2 --warmup 1 --no-cpu-baseline > $D/bench_cfg5.json 2> $D/bench_cfg5.err; echo "rc=$?"; grep "^\[bench\]" $D/bench_cfg5.err | tail -2
python - <<'P'
import json
for w in ('default', 'cfg5', 'reference'):
    try:
        l=json.loads(open('gpurun_out/r02w/bench_%s.json' % w).read().strip().splitlines()[-1])
        print(w, 'value', round(l['value']), l['unit'], 'ms_per_step', round(l['ms_per_step']), 'parity', l.get('parity_md5_ok'), 'launches', l.get('gpu_launches'), 'clocks', l.get('clocks'))
        if w != 'reference':
            print('  e2e', l['e2e']['host_seconds'], 'out', l['e2e']['output_seconds'], 'events', round(sum(l['e2e']['event_seconds'].values()), 3))
            print('  roofline', l['roofline']['kernel'], round(l['roofline']['kernel_ms'], 1), 'ms frac', round(l['roofline']['frac'], 4), 'l1/l2 GB/s', l['roofline'].get('l1_l2_request_gbs'))
            print('  device_ms', {k: round(v, 1) for k, v in l['roofline']['device_ms'].items()})
            print('  cpu_baseline', l.get('cpu_baseline'))
    except Exception as e: print(w, 'failed', e)
P
