#!/bin/bash
cd /root/repo
timeout 300 python -m pytest tests/test_parity_gpu.py -q -x -k "asynchronous or error_behaviour or golden or rerun" 2>&1 | tail -3
