"""Turns gpurun_out/{bench,launches,prof_*}_<tag> into the tracked summaries under profiles/."""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)


def fl(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return 0.0


def raw_metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, r = rows[0], rows[1], rows[2]
    return {h: (r[i], units[i]) for i, h in enumerate(hdr)}


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]

lines = ["# ncu summaries, tag %s" % tag, ""]
traffic = {}
issue = {}
for kern in ("tile", "prep", "vfh"):
    rep = os.path.join(G, "prof_%s_%s.ncu-rep" % (kern, tag))
    if not os.path.exists(rep):
        continue
    m = raw_metrics(rep)
    lines.append("## %s kernel (`ncu --set full --clock-control none`, one launch of `bench.py --steps 2 --warmup 3`)" % kern)
    lines.append("")
    lines.append("| metric | value | unit |")
    lines.append("|---|---|---|")
    for w in WANT:
        if w in m:
            lines.append("| %s | %s | %s |" % (w, m[w][0], m[w][1]))
    lines.append("")
    if kern == "tile":
        def to_bytes(v, u):
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            return fl(v) * mult
        traffic["c4"] = to_bytes(*m["dram__bytes_read.sum"]) + to_bytes(*m["dram__bytes_write.sum"])
        issue["c4"] = {"warp_instructions_per_launch": fl(m["smsp__inst_executed.sum"][0]),
                       "source": "ncu smsp__inst_executed.sum, profiles/ncu_summary_%s.md (tile kernel, one launch)" % tag}
    # stall reasons
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr = rows[1]
    data = [r for r in rows[2:] if len(r) == len(hdr) and r[0] != "Address"]
    si = hdr.index("# Samples")
    tot = sum(fl(r[si]) for r in data) or 1.0
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = sorted(((sum(fl(r[hdr.index(h)]) for r in data), h) for h in stalls), reverse=True)[:6]
    lines.append("warp-state samples: " + ", ".join("%s %.1f%%" % (h, 100 * v / tot) for v, h in agg))
    lines.append("")
    # hottest CUDA source lines (ncu --page source --print-source sass,cuda)
    src2 = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                          capture_output=True, text=True).stdout
    cur_file, hdr2, per_line = None, None, []
    for r in csv.reader(src2.splitlines()):
        if len(r) == 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr2 = r
        elif hdr2 and len(r) == len(hdr2) and r[0] not in ("", "Line No"):
            per_line.append((cur_file, r))
    if per_line:
        s2, i2 = hdr2.index("# Samples"), hdr2.index("Instructions Executed")
        ts = sum(fl(r[s2]) for _, r in per_line) or 1.0
        ti = sum(fl(r[i2]) for _, r in per_line) or 1.0
        lines.append("hottest source lines (share of samples / of executed warp instructions):")
        lines.append("")
        lines.append("| file:line | samples | instr | source |")
        lines.append("|---|---|---|---|")
        for f, r in sorted(per_line, key=lambda fr: -fl(fr[1][s2]))[:12]:
            lines.append("| %s:%s | %.1f%% | %.1f%% | `%s` |" % (f, r[0], 100 * fl(r[s2]) / ts, 100 * fl(r[i2]) / ti,
                                                             r[1].strip()[:90].replace("|", "\\|")))
        lines.append("")
open(os.path.join(P, "ncu_summary_%s.md" % tag), "w").write("\n".join(lines))
if traffic:
    tj = os.path.join(P, "traffic.json")
    old = json.load(open(tj)) if os.path.exists(tj) else {}
    old.update(traffic)
    json.dump(old, open(tj, "w"), indent=1)

if issue:
    ij = os.path.join(P, "issue.json")
    old = json.load(open(ij)) if os.path.exists(ij) else {}
    old.update(issue)
    json.dump(old, open(ij, "w"), indent=1)

# launch list
lc = os.path.join(G, "launches_%s.csv" % tag)
if os.path.exists(lc):
    rows = [r for r in csv.reader(open(lc)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        agg.setdefault(r[ki].split("(")[0][:70], []).append(fl(r[vi]))
    tot = sum(sum(v) for v in agg.values())
    with open(os.path.join(P, "launches_%s.md" % tag), "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none), `bench.py --steps 4 --warmup 3`\n\n")
        f.write("| kernel | launches | mean us | share of listed time |\n|---|---|---|---|\n")
        for k, v in agg.items():
            f.write("| %s | %d | %.1f | %.1f%% |\n" % (k, len(v), sum(v) / len(v) / 1e3, 100 * sum(v) / tot))
    import shutil
    shutil.copy(lc, os.path.join(P, "launches_%s.csv" % tag))
bj = os.path.join(G, "bench_%s.json" % tag)
if os.path.exists(bj):
    import shutil
    shutil.copy(bj, os.path.join(P, "bench_%s.json" % tag))
print("\n".join(lines[:40]))
