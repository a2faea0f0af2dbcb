#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 3 --warmup 3 2> gpurun_out/r2_36_bench.err | tail -1 | python -c "
import sys,json; o=json.loads(sys.stdin.read()); print('bench', o['value'], o['iter_ms'], o['ms_per_step'], o['parity_bits_equal'], o['kernel_ms'], o['gpu_launches'])"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
