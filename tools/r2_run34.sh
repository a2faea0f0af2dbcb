#!/bin/bash
# builder-run bench lines of the tree with k_sweep_q (both arms, default workload), launch list of one pass
set -u
cd /root/repo; mkdir -p gpurun_out
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_ours.json 2> gpurun_out/r02_bench_ours.err; tail -2 gpurun_out/r02_bench_ours.err
timeout 1200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; tail -2 gpurun_out/r02_bench_ref.err
python - <<'PY'
import json
o=json.loads(open('gpurun_out/r02_bench_ours.json').read().strip().splitlines()[-1]); r=json.loads(open('gpurun_out/r02_bench_ref.json').read().strip().splitlines()[-1])
print('ours', o['value'], o['iter_ms'], o['ms_per_step'], o['e2e'], o['parity_bits_equal'], o['kernel_ms'])
print('ref ', r['value'], r['iter_ms'], r['ms_per_step'], r['e2e']['ms_per_call'], r['kernel_ms'])
print('ratio iter', o['value']/r['value'], 'pass', r['ms_per_step']/o['ms_per_step'], 'e2e', r['e2e']['ms_per_call']/o['e2e']['ms_per_call'])
for k in o['roofline']['kernels']: print(k['kernel'], k['frac'], k.get('tex',{}).get('frac'), k.get('tex',{}).get('taps_executed_over_contract'), k['traffic'])
print(o['secondary']['cfg2']['iter_ms'], r['secondary']['cfg2']['iter_ms'], o['secondary']['cfg2'].get('ms_per_step'), r['secondary']['cfg2'].get('ms_per_step'), o['secondary']['cfg4']['wall_ms'])
print(o['stage_ms']); print(r['stage_ms'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_' -c 200 --csv --log-file gpurun_out/r02e_cfg3_launches.csv python tests/tools/time_ours.py cfg3 1 launches > /dev/null 2>&1; wc -l gpurun_out/r02d_cfg3_launches.csv
