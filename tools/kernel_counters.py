#!/usr/bin/env python
"""Per-launch ncu counters of the three NCC kernels -> profiles/kernel_counters.json (read by bench.py for `roofline.traffic`
and `roofline.tex`): DRAM bytes, texture quad requests, data-pipe wavefronts (raw page) and the executed THREAD-LEVEL
texture fetches = sum over TEX instructions of "Predicated-On Thread Instructions Executed" (source page).

    python tools/kernel_counters.py gpurun_out/<report>.ncu-rep <workload> [source description]
"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, workload = sys.argv[1], sys.argv[2]
desc = sys.argv[3] if len(sys.argv) > 3 else f"ncu capture {os.path.basename(rep)}"
MERGE = True        # several reports (one kernel each) may be folded into one workload entry
KERNELS = {"k_strong": "k_strong", "k_weak": "k_weak_q", "k_sweep": "k_sweep"}


def ncu(args):
    return subprocess.run(["ncu", "-i", rep] + args, capture_output=True, text=True).stdout


def to_bytes(v, unit):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


raw = list(csv.reader(io.StringIO(ncu(["--page", "raw", "--csv"]))))
hdr, units = raw[0], raw[1]
col = {h: i for i, h in enumerate(hdr)}
# source page: sections "Kernel Name",<name> / header / rows; every launch is listed twice (two views): keep every other one
sections, cur = [], None
for r in csv.reader(io.StringIO(ncu(["--page", "source", "--csv"]))):
    if len(r) >= 2 and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "tex": 0, "inst": 0}
        sections.append(cur)
    elif cur is not None and cur["hdr"] is None and "Source" in r:
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None:
        h = cur["hdr"]
        try:
            n = int(r[h.index("Instructions Executed")]); p = int(r[h.index("Predicated-On Thread Instructions Executed")])
        except (ValueError, IndexError):
            continue
        cur["inst"] += n
        op = r[h.index("Source")].split()
        if op and (op[1] if op[0].startswith("@") and len(op) > 1 else op[0]).split(".")[0] == "TEX":
            cur["tex"] += p
out = {}
for key, pat in KERNELS.items():
    rows = [r for r in raw[2:] if pat in r[col["Kernel Name"]]]
    secs = [s for s in sections if pat in s["name"]]
    if len(secs) == 2 * len(rows):
        secs = secs[::2]
    if not rows:
        continue
    def mean(name, conv=None):
        vals = []
        for r in rows:
            try:
                v = float(r[col[name]])
            except (KeyError, ValueError):
                continue
            if v != v:
                continue
            vals.append(conv(v, units[col[name]]) if conv else v)
        return sum(vals) / len(vals) if vals else None
    ent = {"launches_captured": len(rows), "ms_under_ncu": mean("gpu__time_duration.sum"), "source": desc}
    rd, wr = mean("dram__bytes_read.sum", to_bytes), mean("dram__bytes_write.sum", to_bytes)
    if rd is not None and wr is not None:
        ent["dram_bytes_per_launch"] = rd + wr
    for name, k in (("l1tex__t_requests_pipe_tex_mem_texture.sum", "tex_quad_requests_per_launch"),
                    ("l1tex__t_output_wavefronts_pipe_tex_mem_texture.sum", "tex_wavefronts_per_launch"),
                    ("l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed", "tex_data_pipe_pct"),
                    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
                    ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"), ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
                    ("launch__registers_per_thread", "registers")):
        v = mean(name)
        if v is not None:
            ent[k] = v
    if secs and any(s["tex"] for s in secs):
        ent["tex_thread_fetches_per_launch"] = sum(s["tex"] for s in secs) / len(secs)
        ent["warp_instructions_per_launch"] = sum(s["inst"] for s in secs) / len(secs)
    out[key] = ent
path = os.path.join(ROOT, "profiles", "kernel_counters.json")
allc = json.load(open(path)) if os.path.exists(path) else {}
allc.setdefault(workload, {}).update(out)
json.dump(allc, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
