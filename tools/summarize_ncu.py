#!/usr/bin/env python
"""Summarise ncu CSV exports for profiles/: (1) a gpu__time_duration launch list -> per-kernel totals of one steady frame,
(2) a --set full raw page -> per-launch duration, DRAM bytes, tensor-pipe and L2 figures."""
import csv, sys, json, collections, re

def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("mega::", "").replace("void ", "")[:70]

def launch_list(path, first_marker="stem_prep"):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    k, v, g = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    names = [short(r[k]) for r in rows]
    starts = [i for i, n in enumerate(names) if first_marker in n]
    if len(starts) >= 2:
        rows, names = rows[starts[-2]:starts[-1]], names[starts[-2]:starts[-1]]
    tot = collections.OrderedDict()
    for r, n in zip(rows, names):
        t = tot.setdefault(n, [0, 0.0]); t[0] += 1; t[1] += float(r[v].replace(",", "")) / 1e3
    total = sum(t[1] for t in tot.values())
    out = ["| kernel | launches | us (serialised, cold) | share |", "|---|---:|---:|---:|"]
    for n, (c, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{n}` | {c} | {us:.1f} | {us / total:.3f} |")
    out.append(f"| total | {sum(t[0] for t in tot.values())} | {total:.1f} | 1 |")
    return "\n".join(out)

def full_raw(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    want = {"Kernel Name": "kernel", "gpu__time_duration.sum": "ns", "dram__bytes_read.sum": "dram_rd", "dram__bytes_write.sum": "dram_wr",
            "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
            "sm__inst_executed_pipe_uniform.sum": None, "launch__registers_per_thread": "regs",
            "lts__t_sector_hit_rate.pct": "l2_hit", "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct", "launch__grid_size": "grid"}
    idx = {want[h]: i for i, h in enumerate(hdr) if h in want and want[h]}
    units = rows[1]
    out = []
    for r in rows[2:]:
        d = {k: r[i] for k, i in idx.items()}
        d["kernel"] = short(d["kernel"])
        for k in ("dram_rd", "dram_wr"):
            if k in d:
                d[k] = f"{d[k]} {units[idx[k]]}"
        out.append(d)
    return out

if __name__ == "__main__":
    if sys.argv[1] == "list":
        print(launch_list(sys.argv[2]))
    else:
        print(json.dumps(full_raw(sys.argv[2]), indent=1))
