#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 300 python tests/tools/sweep_ab.py 64 48 2 2>&1 | tail -8
timeout 300 python tests/tools/sweep_ab.py 320 240 4 2>&1 | tail -8
