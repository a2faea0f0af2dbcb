#!/bin/bash
set -u
cd /root/repo; mkdir -p gpurun_out; rm -f gpurun_out/time_ours.jsonl
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q 2>&1 | tail -3
for p in 1 0; do for c in cfg3 cfg2 cfg3s; do
APD_SQ_PIPE=$p timeout 300 python tests/tools/time_ours.py $c 2 sweepq3_p$p 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pipe $p', d['case'], d['crc']['planes'], d['crc']['states'], d['iter_ms'], d['total_ms'], 'K14', d['stage_ms']['K14 classify'])"
done; done
APD_SQ_PIPE=1 APD_SQ_BLOCKS=3 timeout 300 python tests/tools/time_ours.py cfg3 2 sweepq3_p1_b3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pipe 1 blocks 3', d['case'], d['crc']['planes'], d['crc']['states'], d['iter_ms'], d['total_ms'], 'K14', d['stage_ms']['K14 classify'])"
