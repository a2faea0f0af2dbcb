#!/bin/bash
# round 2, call 3: new bench.py (cfg3s quick, then the default cfg3 line of both arms) + full-resolution ncu captures
set -u
mkdir -p gpurun_out
timeout 600 python bench.py --workload cfg3s --steps 3 --warmup 3 --no-secondary > gpurun_out/r2_3_bench_cfg3s.json 2> gpurun_out/r2_3_bench_cfg3s.err
tail -c 1500 gpurun_out/r2_3_bench_cfg3s.json; tail -5 gpurun_out/r2_3_bench_cfg3s.err
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_3_bench_ours.json 2> gpurun_out/r2_3_bench_ours.err
tail -c 3000 gpurun_out/r2_3_bench_ours.json; tail -5 gpurun_out/r2_3_bench_ours.err
cp profiles/contract_masks.json gpurun_out/contract_masks.json 2>/dev/null
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_3_bench_ref.json 2> gpurun_out/r2_3_bench_ref.err
tail -c 1500 gpurun_out/r2_3_bench_ref.json; tail -5 gpurun_out/r2_3_bench_ref.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_weak_q|k_gen_anchors|k_sweep' -c 3 -o gpurun_out/r02c3_full -f \
  python tests/tools/time_ours.py cfg3 1 ncu > gpurun_out/r2_3_ncu.log 2>&1
tail -3 gpurun_out/r2_3_ncu.log
