#!/usr/bin/env python
"""Turn gpurun_out/<tag>_full.ncu-rep and gpurun_out/<tag>_launches.csv into the tracked summaries under profiles/.

  python tools/summarise_ncu.py r01f

writes profiles/<tag>_ncu_summary.json (per captured kernel: duration, DRAM bytes, pipe utilisation, stall mix,
instruction mix per texture instruction), profiles/<tag>_launches.csv (kernel name, launch count, total and mean
gpu__time_duration) and profiles/traffic.json (per-launch DRAM bytes of the dominant kernel, read by bench.py)."""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tex.sum.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "l1tex__t_requests_pipe_tex_mem_texture.sum", "l1tex__t_sectors_pipe_tex_mem_texture.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "gpc__cycles_elapsed.avg.per_second",
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True, check=True).stdout


def to_bytes(v, unit):
    f = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit)
    return None if f is None else float(v) * f


def main():
    tag = sys.argv[1]
    rep = os.path.join(ROOT, "gpurun_out", f"{tag}_full.ncu-rep")
    out = {"source": f"ncu --set full --clock-control none --import-source on (gpurun_out/{tag}_full.ncu-rep, not tracked)", "kernels": []}
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        k = {"name": r[hdr.index("Kernel Name")], "id": r[hdr.index("ID")]}
        for key in KEYS:
            if key in hdr:
                i = hdr.index(key)
                try:
                    k[key] = float(r[i]); k[key + ".unit"] = units[i]
                except ValueError:
                    pass
        stalls = {}
        for i, h in enumerate(hdr):
            m = re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio", h)
            if m:
                try:
                    stalls[m.group(1)] = round(float(r[i]), 3)
                except ValueError:
                    pass
        k["stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8])
        if "dram__bytes_read.sum" in k:
            rd = to_bytes(k["dram__bytes_read.sum"], k["dram__bytes_read.sum.unit"])
            wr = to_bytes(k["dram__bytes_write.sum"], k["dram__bytes_write.sum.unit"])
            k["dram_bytes_per_launch"] = None if rd is None or wr is None else rd + wr
        out["kernels"].append(k)
    # instruction mix per texture instruction from the source page
    for k in out["kernels"]:
        try:
            src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--launch-id" if False else "--kernel-id", f":::{int(k['id']) + 1}"] if False else ["-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + re.escape(k["name"].split("(")[0].split("<")[0].split()[-1])]))))
        except subprocess.CalledProcessError:
            continue
        h = next((row for row in src if "Source" in row and "Instructions Executed" in row), None)
        if h is None:
            continue
        ia, ie = h.index("Source"), h.index("Instructions Executed")
        mix = collections.Counter()
        for row in src[src.index(h) + 1:]:
            if len(row) <= ie:
                continue
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", row[ia])
            try:
                n = int(row[ie])
            except ValueError:
                continue
            if m:
                mix[m.group(2)] += n
        if mix.get("TEX"):
            tot = sum(mix.values())
            k["warp_inst_per_tex"] = round(tot / mix["TEX"], 2)
            k["inst_mix_per_tex"] = {op: round(n / mix["TEX"], 2) for op, n in mix.most_common(16)}
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "profiles", f"{tag}_ncu_summary.json"), "w"), indent=1)
    # traffic of the dominant kernel (mean over its captured launches)
    tr = [k["dram_bytes_per_launch"] for k in out["kernels"] if "k_strong" in k["name"] and k.get("dram_bytes_per_launch")]
    if tr and tag.startswith("r01"):        # round 1 file layout; round 2: tools/kernel_counters.py -> profiles/kernel_counters.json
        json.dump({"kernel": "k_strong", "workload": "cfg2", "dram_bytes_per_launch": sum(tr) / len(tr), "launches_captured": len(tr),
                   "source": f"profiles/{tag}_ncu_summary.json"}, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    # launch list
    ll = os.path.join(ROOT, "gpurun_out", f"{tag}_launches.csv")
    if os.path.exists(ll):
        txt = open(ll).read()
        txt = txt[txt.index('"ID"'):]
        rows = list(csv.reader(io.StringIO(txt)))
        h = rows[0]
        iname, ival, iunit = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
        agg = collections.OrderedDict()
        for r in rows[1:]:
            if len(r) <= ival:
                continue
            v = float(r[ival].replace(",", ""))
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iunit], 1e-6)
            a = agg.setdefault(r[iname].split("(")[0], [0, 0.0])
            a[0] += 1; a[1] += v
        total = sum(a[1] for a in agg.values())
        with open(os.path.join(ROOT, "profiles", f"{tag}_launches.csv"), "w") as f:
            f.write("kernel,launches,total_ms,mean_ms,share\n")
            for name, (n, ms) in agg.items():
                f.write(f"\"{name}\",{n},{ms:.3f},{ms / n:.4f},{ms / total:.4f}\n")
    print("ok")


if __name__ == "__main__":
    main()
