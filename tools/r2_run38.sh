#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 900 python tests/tools/radius_fuzz.py 8 > gpurun_out/r02c_radius_fuzz.log 2>&1; tail -2 gpurun_out/r02c_radius_fuzz.log; grep -v OK gpurun_out/r02c_radius_fuzz.log | head -5
