#!/bin/bash
# round 2, call 4: whole GPU suite, new bench on cfg3s / cfg3 (both arms), the reference's main() on the pooled facade
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_4_pytest.log
cat gpurun_out/r2_4_pytest.log
timeout 600 python bench.py --workload cfg3s --steps 3 --warmup 3 --no-secondary > gpurun_out/r2_4_bench_cfg3s.json 2> gpurun_out/r2_4_bench_cfg3s.err
tail -c 1200 gpurun_out/r2_4_bench_cfg3s.json; tail -5 gpurun_out/r2_4_bench_cfg3s.err
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_4_bench_ours.json 2> gpurun_out/r2_4_bench_ours.err
tail -c 6000 gpurun_out/r2_4_bench_ours.json; tail -5 gpurun_out/r2_4_bench_ours.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_4_bench_ref.json 2> gpurun_out/r2_4_bench_ref.err
tail -c 2500 gpurun_out/r2_4_bench_ref.json; tail -5 gpurun_out/r2_4_bench_ref.err
timeout 600 python tests/tools/main_program_bench.py --out gpurun_out/r02_main_program_bench.json 2>&1 | tail -3
