#!/bin/bash
set -u
cd /root/repo; mkdir -p gpurun_out
APD_SQ_PIPE=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_sweep_q' -c 1 -o gpurun_out/r02f_cfg3s_sweepq -f python tests/tools/time_ours.py cfg3s 1 ncu > gpurun_out/r2_26_ncu.log 2>&1; tail -1 gpurun_out/r2_26_ncu.log | cut -c1-100
