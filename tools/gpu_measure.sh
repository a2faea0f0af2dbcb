#!/bin/bash
# One gpurun call: both bench arms, the ncu launch list of the bench command and one --set full capture of the
# dominant kernels. Outputs land in gpurun_out/ (scratch); summaries are copied to profiles/ by tools/summarise_ncu.py.
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
tail -c 600 gpurun_out/bench_ref.json; echo; tail -c 2500 gpurun_out/bench_ours.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_' -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launch_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_strong|k_sweep|k_init_planes' -s 6 -c 4 \
    -o gpurun_out/${TAG}_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/full_bench.log 2>&1
ls -la gpurun_out | tail -8
