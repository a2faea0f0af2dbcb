#!/bin/bash
# round 2, call 14: randomised parity sweep (350 configurations, new seed) and compute-sanitizer over the final kernels
set -u
cd /root/repo; mkdir -p gpurun_out
timeout 1500 python tests/tools/parity_fuzz.py 350 777 > gpurun_out/r02_parity_fuzz.log 2>&1; echo "fuzz rc=$?"; tail -2 gpurun_out/r02_parity_fuzz.log
cp gpurun_out/parity_fuzz.json gpurun_out/r02_parity_fuzz.json
bash tools/sanitize.sh
