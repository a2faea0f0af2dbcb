#!/bin/bash
# round 2, call 1: parity of the quad-per-pixel k_weak_q + A/B timing against the first design
set -u
mkdir -p gpurun_out
rm -f gpurun_out/time_ours.jsonl
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q 2>&1 | tail -15 > gpurun_out/r2_1_pytest.log
cat gpurun_out/r2_1_pytest.log
timeout 300 python tests/tools/parity_probe.py apd > gpurun_out/r2_1_probe_apd.log 2>&1; tail -40 gpurun_out/r2_1_probe_apd.log
for c in cfg3s cfg3; do
  APD_WEAK_IMPL=old timeout 300 python tests/tools/time_ours.py $c 2 old 2>&1 | tail -1
  timeout 300 python tests/tools/time_ours.py $c 2 quad 2>&1 | tail -1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_weak_q|k_gen_anchors' -c 2 -o gpurun_out/r02w_full -f \
  python tests/tools/time_ours.py cfg3s 1 ncu > gpurun_out/r2_1_ncu.log 2>&1
tail -3 gpurun_out/r2_1_ncu.log
ls -la gpurun_out | tail -5
