#!/bin/bash
cd /root/repo
timeout 300 python -m pytest tests/test_main_program_gpu.py tests/test_facade.py -q -x 2>&1 | tail -4
