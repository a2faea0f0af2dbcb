#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
cp apd_mvs_b200/libapd_b200.so /tmp/base.so
echo "== base"; timeout 600 python -m pytest tests/test_parity_gpu.py -q -k "golden or live" 2>&1 | tail -3
cp gpurun_tmp/lib_SQ_SEQ_DEC.so apd_mvs_b200/libapd_b200.so
echo "== sequential decision"; timeout 600 python -m pytest tests/test_parity_gpu.py -q -k "golden or live" 2>&1 | tail -3
cp /tmp/base.so apd_mvs_b200/libapd_b200.so
