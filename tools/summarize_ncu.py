"""Turns gpurun_out/{launches.csv, *.ncu-rep} into a markdown summary under profiles/.
usage: python tools/summarize_ncu.py <launches.csv> <report.ncu-rep> <out.md> [title]"""
import collections
import csv
import subprocess
import sys

launches, rep, out = sys.argv[1:4]
title = sys.argv[4] if len(sys.argv) > 4 else "ncu summary"
lines = ["# " + title, ""]

rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    v *= {"us": 1e-3, "ns": 1e-6, "ms": 1.0, "s": 1e3}.get(r[ui], 1.0)
    agg.setdefault(r[ki].split("(")[0][-70:], []).append(v)
tot = sum(sum(v) for v in agg.values())
lines += ["## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: compare shares)", "",
          "| kernel | launches | total ms | share | avg ms |", "|---|---:|---:|---:|---:|"]
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:14]:
    lines.append("| `%s` | %d | %.3f | %.1f%% | %.4f |" % (k, len(v), sum(v), 100 * sum(v) / tot, sum(v) / len(v)))
lines.append("")

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, units = rr[0], rr[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_static", "launch__grid_size",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__cycles_active.avg", "sm__cycles_elapsed.max", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_atom.sum"]
lines += ["## `ncu --set full` (per launch)", ""]
for r in rr[2:]:
    lines += ["### `%s`" % r[h.index("Kernel Name")].split("(")[0][-70:], "", "| metric | value | unit |", "|---|---:|---|"]
    for w in want:
        if w in h:
            lines.append("| %s | %s | %s |" % (w, r[h.index(w)], units[h.index(w)]))
    stalls = sorted(((h[i], float(r[i])) for i in range(len(h))
                     if "issue_stalled" in h[i] and h[i].endswith("per_warp_active.pct") and r[i]), key=lambda x: -x[1])[:6]
    if stalls:
        lines += ["", "top stall reasons (% of active warps): " +
                  ", ".join("%s %.1f" % (k.split("issue_stalled_")[1].replace("_per_warp_active.pct", ""), v) for k, v in stalls)]
    lines.append("")
open(out, "w").write("\n".join(lines) + "\n")
print("wrote", out)
