#!/bin/bash
# ncu launch list of the bench command with the final kernels (own arm, cfg3): per-kernel shares of a step
cd /root/repo; mkdir -p gpurun_out
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02e_cfg3_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-parity --no-secondary --no-cpu-baseline > gpurun_out/r2_60_bench.log 2>&1
wc -l gpurun_out/r02e_cfg3_bench_launches.csv; tail -1 gpurun_out/r2_60_bench.log | cut -c1-160
