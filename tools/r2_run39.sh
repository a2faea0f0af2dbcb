#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 600 python tests/tools/dup_stats.py cfg3s 2>&1 | tail -12
timeout 600 python tests/tools/dup_stats.py cfg2q 2>&1 | tail -12
