#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 400 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity 2> gpurun_out/r2_58.err | tail -1 | python -c "
import sys,json; o=json.loads(sys.stdin.read()); print('bench', o['value'], o['iter_ms'], o['ms_per_step']); print(o['secondary']['cfg4']['wall_ms'], o['secondary']['cfg4']['per_rank_process_ms'], o['secondary']['cfg2']['iter_ms'])"
