#!/usr/bin/env python
"""Digest of an ncu report (raw page + source page) for one kernel: pipe utilisation, stalls, texture wavefronts,
lane utilisation of the TEX instructions, instruction mix per TEX. Usage: tools/ncu_digest.py <rep> <kernel regex>"""
import collections, csv, io, re, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "l1tex__t_requests_pipe_tex_mem_texture.sum",
        "l1tex__t_output_wavefronts_pipe_tex_mem_texture.sum", "l1tex__t_sectors_pipe_tex_mem_texture.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum",
        "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmalite.sum.pct_of_peak_sustained_active"]
for r in rows[2:]:
    print("=====", r[hdr.index("Kernel Name")][:60])
    for k in keys:
        if k in hdr:
            print(f"  {k:75s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}")
    st = {}
    for i, h in enumerate(hdr):
        m = re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio", h)
        if m:
            try: st[m.group(1)] = round(float(r[i]), 3)
            except ValueError: pass
    print("  stalls/issue:", sorted(st.items(), key=lambda kv: -kv[1])[:8])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
his = [i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r]
if his:
    hi = his[0]; h = rows[hi]
    s, ie, te, po, sm = h.index("Source"), h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("Predicated-On Thread Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
    mix, pon, smp = collections.Counter(), collections.Counter(), collections.Counter()
    data = []
    for r in rows[hi + 1:]:
        if len(r) <= ie or not r[ie].isdigit(): continue
        if hi != his[0]: break
        op = r[s].split()
        if not op: continue
        o = (op[1] if op[0].startswith("@") else op[0]).split(".")[0]
        mix[o] += int(r[ie]); pon[o] += int(r[po]); smp[o] += int(r[sm] or 0)
        data.append((int(r[sm] or 0), int(r[ie]), r[s]))
    tot = sum(mix.values()); tex = max(mix["TEX"], 1)
    print(f"  warp instructions {tot}, TEX {mix['TEX']}, per TEX {tot / tex:.2f}, predicated-on lanes per TEX {pon['TEX'] / tex:.2f}, thread-level fetches {pon['TEX']}")
    print("  mix per TEX:", {k: round(v / tex, 2) for k, v in mix.most_common(18)})
    tots = sum(smp.values())
    print("  top stall sites:")
    for x in sorted(data, key=lambda x: -x[0])[:14]:
        print(f"    {100.0 * x[0] / max(tots, 1):5.1f}%  x{x[1]:<10d} {x[2][:80]}")
