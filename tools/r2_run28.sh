#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -q 2>&1 | tail -60 > gpurun_out/r2_28_pytest.log
tail -30 gpurun_out/r2_28_pytest.log
