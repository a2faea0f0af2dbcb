#!/bin/bash
set -u
cd /root/repo; mkdir -p gpurun_out; rm -f gpurun_out/time_ours.jsonl
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q 2>&1 | tail -3
for c in cfg3 cfg2 cfg3s; do
timeout 300 python tests/tools/time_ours.py $c 2 sweepq4 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['case'], d['crc']['planes'], d['crc']['states'], d['iter_ms'], d['total_ms'], 'K14', d['stage_ms']['K14 classify'])"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_sweep_q' -c 1 -o gpurun_out/r02g_cfg3s_sweepq -f python tests/tools/time_ours.py cfg3s 1 ncu > gpurun_out/r2_27_ncu.log 2>&1; tail -1 gpurun_out/r2_27_ncu.log | cut -c1-100
