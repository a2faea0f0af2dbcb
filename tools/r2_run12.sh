#!/bin/bash
# round 2, call 12: k_sweep / K3 ncu captures (full-size counters, quarter-size --set full), whole GPU suite on the final tree
set -u
cd /root/repo; mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
M=dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_requests_pipe_tex_mem_texture.sum,l1tex__t_output_wavefronts_pipe_tex_mem_texture.sum,l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,launch__registers_per_thread,gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active
timeout 900 ncu --section SourceCounters --metrics $M --clock-control none --import-source on -k regex:k_sweep -c 1 -o gpurun_out/r02_cfg3_k_sweep -f python tests/tools/time_ours.py cfg3 1 ncu > gpurun_out/r2_12_ncu_k_sweep.log 2>&1
tail -1 gpurun_out/r2_12_ncu_k_sweep.log | cut -c1-120
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_sweep|k_gen_anchors' -c 2 -o gpurun_out/r02b_cfg3s_full -f python tests/tools/time_ours.py cfg3s 1 ncu > gpurun_out/r2_12_ncu_full.log 2>&1
tail -1 gpurun_out/r2_12_ncu_full.log | cut -c1-120
