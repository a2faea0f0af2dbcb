#!/bin/bash
# round 2, call 2: k_weak_q v2 (packed two-slot math, lane-slot scheduling): parity + timing of the 3- and 4-block register variants
set -u
mkdir -p gpurun_out
rm -f gpurun_out/time_ours.jsonl
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q 2>&1 | tail -15 > gpurun_out/r2_2_pytest.log
cat gpurun_out/r2_2_pytest.log
for c in cfg3s cfg3; do
  APD_WQ_BLOCKS=4 timeout 300 python tests/tools/time_ours.py $c 2 quad4 2>&1 | tail -1
  APD_WQ_BLOCKS=3 timeout 300 python tests/tools/time_ours.py $c 2 quad3 2>&1 | tail -1
done
APD_WQ_BLOCKS=${NCU_BLOCKS:-4} timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_weak_q' -c 1 -o gpurun_out/r02w2_full -f \
  python tests/tools/time_ours.py cfg3s 1 ncu > gpurun_out/r2_2_ncu.log 2>&1
tail -3 gpurun_out/r2_2_ncu.log
