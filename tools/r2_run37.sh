#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
export PYTHONPATH=.:tests
for tool in initcheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 python tools/initcheck_case.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; tail -3 gpurun_out/sanitize_$tool.log
done
timeout 1500 python tests/tools/parity_fuzz.py 300 90210 > gpurun_out/r02c_parity_fuzz.log 2>&1; tail -1 gpurun_out/r02c_parity_fuzz.log; cp gpurun_out/parity_fuzz.json gpurun_out/r02c_parity_fuzz.json
