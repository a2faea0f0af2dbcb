#!/bin/bash
cd /root/repo; mkdir -p gpurun_out; rm -f gpurun_out/time_ours.jsonl
timeout 900 python -m pytest tests/test_parity_gpu.py -q 2>&1 | tail -3
timeout 300 python tests/tools/sweep_ab.py 64 48 2 2>&1 | grep "states differ"
for c in cfg3 cfg2 cfg3s; do
timeout 300 python tests/tools/time_ours.py $c 2 sweepq5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['case'], d['crc']['planes'], d['crc']['states'], d['iter_ms'], d['total_ms'], 'K14', d['stage_ms']['K14 classify'])"
done
