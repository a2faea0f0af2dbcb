#!/bin/bash
# ncu launch list of one cfg3 pass with the final kernels (own kernels only: the scene generator's torch kernels are filtered out)
cd /root/repo; mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_' -c 200 --csv --log-file gpurun_out/r02e_cfg3_launches.csv python tests/tools/time_ours.py cfg3 1 launches > /dev/null 2>&1
wc -l gpurun_out/r02e_cfg3_launches.csv
