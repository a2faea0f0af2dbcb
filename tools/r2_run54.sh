#!/bin/bash
cd /root/repo; mkdir -p gpurun_out; rm -f gpurun_out/time_ours.jsonl
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 150 python tests/tools/time_ours.py cfg3 2 final 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); s=d['stage_ms']; print(d['case'], d['crc'], 'iter', d['iter_ms'], 'total', d['total_ms'])"
