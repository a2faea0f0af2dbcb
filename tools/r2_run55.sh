#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_strong' -s 2 -c 1 -o gpurun_out/r02h_cfg3s_strong_full -f python tests/tools/time_ours.py cfg3s 1 ncu > gpurun_out/r2_55_ncu.log 2>&1; tail -1 gpurun_out/r2_55_ncu.log | cut -c1-100
