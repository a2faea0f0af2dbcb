#!/bin/bash
# round 2, call 13: timing + parity after hoisting the geometric term; GPU converter test
set -u
cd /root/repo; mkdir -p gpurun_out; rm -f gpurun_out/time_ours.jsonl
timeout 600 python -m pytest tests/test_colmap_gpu.py tests/test_parity_gpu.py -x -q 2>&1 | tail -3
timeout 300 python tests/tools/time_ours.py cfg3 2 geomhoist 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['crc']['planes'], d['iter_ms'], d['total_ms'], d['stage_ms']['K14 classify'], d['stage_ms']['it1 K9 weak black'], d['stage_ms']['it1 K6 strong black'])"
