#!/bin/bash
set -u
cd /root/repo; mkdir -p gpurun_out; rm -f gpurun_out/time_ours.jsonl
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q 2>&1 | tail -2
timeout 300 python tests/tools/time_ours.py cfg3 2 k3skip 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['crc']['planes'], d['iter_ms'], d['total_ms'], 'K3', d['stage_ms']['K3 anchors'])"
timeout 300 python tests/tools/time_ours.py cfg3s 2 k3skip 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['crc']['planes'], d['iter_ms'], d['total_ms'], 'K3', d['stage_ms']['K3 anchors'])"
