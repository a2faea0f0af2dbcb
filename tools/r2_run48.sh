#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "asynchronous or golden" 2>&1 | tail -5
timeout 1200 python bench.py --steps 3 --warmup 3 2> gpurun_out/r2_48_bench.err | tail -1 | python -c "
import sys,json; o=json.loads(sys.stdin.read()); print('bench', o['value'], o['iter_ms'], o['ms_per_step'], o['parity_bits_equal'], o['e2e'])"
tail -3 gpurun_out/r2_48_bench.err
