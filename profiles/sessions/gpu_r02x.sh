#!/bin/bash
# r02x: the L=151 workload (cfg3_15M_2x151_75k: the largest 2x151 workload whose reference run fits the development container's memory and time)
set -u
D=gpurun_out/r02x; mkdir -p $D
echo "== bench cfg3 N=1"; ARB_TRACE=1 timeout 1500 python bench.py --workload cfg3_15M_2x151_75k --steps 2 --warmup 1 --no-cpu-baseline > $D/bench_cfg3.json 2> $D/bench_cfg3.err; echo "rc=$?"; grep "^\[bench\]" $D/bench_cfg3.err | tail -3
grep "^\[laps\]\|^\[ingest\]" /tmp/arb_bench/cfg3_15M_2x151_75k/out_rank0/library_stderr.log | tail -70 > $D/host_stage_laps_cfg3.txt
python - <<'P'
import json
l=json.loads(open('gpurun_out/r02x/bench_cfg3.json').read().strip().splitlines()[-1])
print('cfg3 value', round(l['value']), 'ms_per_step', round(l['ms_per_step']), 'fragments', l['fragments_per_step'], 'parity', l['parity'])
print('  e2e', l['e2e']['host_seconds'], 'out', l['e2e']['output_seconds'], 'events', round(sum(l['e2e']['event_seconds'].values()), 3))
print('  device_ms', {k: round(v, 1) for k, v in l['roofline']['device_ms'].items()})
for k in l['roofline']['kernels']: print('  ', k['kernel'][:44], round(k['kernel_ms'],2), 'ms frac', round(k['frac'],4))
P
