#!/bin/bash
# round 2, call 9: whole GPU suite after the K3 fix + T&T fusion variants; bench cfg3 (ours) with the parity check; facade bench
set -u
cd /root/repo; mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_9_pytest.log
cat gpurun_out/r2_9_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_9_bench_ours.json 2> gpurun_out/r2_9_bench_ours.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_9_bench_ours.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','iter_ms','ms_per_step','kernel_ms','parity_bits_equal','output_crc32')})
print(d['stage_ms'])
print(d.get('secondary'))
PY
tail -3 gpurun_out/r2_9_bench_ours.err
timeout 600 python tests/tools/main_program_bench.py --out gpurun_out/r02_main_program_bench.json 2>&1 | tail -2
APD_B200_POOL=0 timeout 600 python tests/tools/main_program_bench.py --out gpurun_out/r02_main_program_bench_nopool.json 2>&1 | tail -2
