#!/bin/bash
set -u
cd /root/repo; mkdir -p gpurun_out; rm -f gpurun_out/time_ours.jsonl
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q 2>&1 | tail -2
for s in 0 1; do for c in cfg3 cfg3s; do
APD_WQ_SORT=$s timeout 300 python tests/tools/time_ours.py $c 2 sort$s 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['case'], d['tag'], d['crc']['planes'], d['iter_ms'], d['total_ms'], [d['stage_ms'][k] for k in d['stage_ms'] if 'weak' in k])"
done; done
