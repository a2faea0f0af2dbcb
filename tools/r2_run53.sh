#!/bin/bash
cd /root/repo; mkdir -p gpurun_out; rm -f gpurun_out/time_ours.jsonl
timeout 150 python -m pytest tests/test_parity_gpu.py -x -q -k "golden or live" 2>&1 | tail -2
for c in cfg3 cfg3s; do
timeout 150 python tests/tools/time_ours.py $c 2 weakgeom2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); s=d['stage_ms']; print(d['case'], d['crc'], 'iter', d['iter_ms'], 'total', d['total_ms'], 'weak', [round(v,2) for k,v in s.items() if 'weak' in k])"
done
