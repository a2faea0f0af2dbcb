#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
M=l1tex__t_requests_pipe_tex_mem_texture.sum,l1tex__t_output_wavefronts_pipe_tex_mem_texture.sum,gpu__time_duration.sum,l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__t_sector_hit_rate.pct
for v in 0 1; do
  if [ $v = 1 ]; then export APD_DBG_NOANCHOR=1; fi
  timeout 600 ncu --metrics $M --clock-control none -k regex:'k_weak_q' -c 2 --csv --log-file gpurun_out/r2_35_noanchor$v.csv python tests/tools/time_ours.py cfg3 1 dbg > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_35_noanchor$v.csv')) if len(r)>10 and r[0].isdigit()]
d={}
for r in rows: d.setdefault(r[0],{})[r[-3]]=float(r[-1].replace(',',''))
for k,m in d.items(): print('noanchor=$v', k, {a.split('.')[0][-40:]:b for a,b in m.items()}, 'wf/req', m['l1tex__t_output_wavefronts_pipe_tex_mem_texture.sum']/m['l1tex__t_requests_pipe_tex_mem_texture.sum'])
PY
done
