#!/bin/bash
set -u
cd /root/repo; mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_sweep_q' -c 1 -o gpurun_out/r02_cfg3_k_sweep_q -f python tests/tools/time_ours.py cfg3 1 ncu > gpurun_out/r2_33_ncu.log 2>&1; tail -1 gpurun_out/r2_33_ncu.log | cut -c1-100
timeout 1500 python tests/tools/parity_fuzz.py 200 4242 > gpurun_out/r02b_parity_fuzz.log 2>&1; tail -1 gpurun_out/r02b_parity_fuzz.log; cp gpurun_out/parity_fuzz.json gpurun_out/r02b_parity_fuzz.json
bash tools/sanitize.sh
