#!/bin/bash
cd /root/repo; mkdir -p gpurun_out; rm -f gpurun_out/time_ours.jsonl
for c in cfg3 cfg2; do
APD_SQ_BLOCKS=5 timeout 300 python tests/tools/time_ours.py $c 2 sweepq5_b5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('blocks 5', d['case'], d['crc']['planes'], d['crc']['states'], d['iter_ms'], d['total_ms'], 'K14', d['stage_ms']['K14 classify'])"
done
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
