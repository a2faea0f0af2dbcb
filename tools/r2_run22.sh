#!/bin/bash
set -u
cd /root/repo; mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_sweep_q' -c 1 -o gpurun_out/r02d_cfg3s_sweepq -f python tests/tools/time_ours.py cfg3s 1 ncu > gpurun_out/r2_22_ncu.log 2>&1; tail -1 gpurun_out/r2_22_ncu.log | cut -c1-100
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_sweep_q' -c 1 -o gpurun_out/r02d_mid_sweepq -f python tests/tools/time_ours.py mid 1 ncu > gpurun_out/r2_22b_ncu.log 2>&1; tail -1 gpurun_out/r2_22b_ncu.log | cut -c1-100
APD_SWEEP_IMPL=old timeout 300 python tests/tools/time_ours.py mid 2 old 2>&1 | tail -1 | cut -c 1-200
