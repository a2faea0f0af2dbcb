#!/bin/bash
# round 2, call 11: k_strong register variants; ncu evidence of the final kernels (quarter-size --set full, full-size counters, launch list)
set -u
cd /root/repo; mkdir -p gpurun_out; rm -f gpurun_out/time_ours.jsonl
for v in 0 1 2 3; do APD_STRONG_VARIANT=$v timeout 300 python tests/tools/time_ours.py cfg2 2 strong$v 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['tag'], d['iter_ms'], d['stage_ms']['it1 K6 strong black'], d['crc']['planes'])"; done
M=dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_requests_pipe_tex_mem_texture.sum,l1tex__t_output_wavefronts_pipe_tex_mem_texture.sum,l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,launch__registers_per_thread,gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active
for k in k_strong k_weak_q k_sweep; do
  timeout 900 ncu --section SourceCounters --metrics $M --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/r02_cfg3_$k -f python tests/tools/time_ours.py cfg3 1 ncu > gpurun_out/r2_11_ncu_$k.log 2>&1
  tail -1 gpurun_out/r2_11_ncu_$k.log | cut -c1-120
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_strong|k_weak_q|k_sweep|k_gen_anchors' -s 1 -c 6 -o gpurun_out/r02_cfg3s_full -f python tests/tools/time_ours.py cfg3s 1 ncu > gpurun_out/r2_11_ncu_full.log 2>&1
tail -1 gpurun_out/r2_11_ncu_full.log | cut -c1-120
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_|apd::' -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-secondary --no-parity --no-cpu-baseline > gpurun_out/r2_11_launch_bench.log 2>&1
tail -c 300 gpurun_out/r2_11_launch_bench.log
