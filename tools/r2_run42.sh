#!/bin/bash
cd /root/repo; mkdir -p gpurun_out; rm -f gpurun_out/time_ours.jsonl
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q 2>&1 | tail -2
for c in cfg3 cfg2 cfg3s; do
timeout 300 python tests/tools/time_ours.py $c 3 strongroll 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); s=d['stage_ms']; print(d['case'], d['crc']['planes'], d['crc']['states'], 'iter', d['iter_ms'], 'total', d['total_ms'], 'strong', [round(v,2) for k,v in s.items() if 'strong' in k], 'K5', s['K5 init'])"
done
