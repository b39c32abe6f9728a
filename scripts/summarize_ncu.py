#!/usr/bin/env python
"""Turns ncu outputs brought back in gpurun_out/ into the markdown summaries kept under profiles/.

  launch list : ncu --metrics gpu__time_duration.sum --csv --log-file X.csv <cmd>
  full capture: ncu --set full -o X <cmd>      (read here with `ncu -i X.ncu-rep --page raw --csv`)
"""
import collections
import csv
import subprocess
import sys

FULL = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def launch_list(path, anchor, skip=("spin_kernel",)):
    rows = list(csv.DictReader(l for l in open(path) if not l.startswith("==")))
    names = [r["Kernel Name"] for r in rows]
    idx = [i for i, n in enumerate(names) if anchor in n]
    a, b = idx[-2], idx[-1]                     # one whole eager step between two forward launches
    step = [r for r in rows[a:b] if not any(s in r["Kernel Name"] for s in skip)]
    agg = collections.OrderedDict()
    for r in step:
        k = r["Kernel Name"].split("(")[0][-70:]
        c = agg.setdefault(k, [0, 0.0, r["Grid Size"], r["Block Size"]])
        c[0] += 1
        c[1] += float(r["Metric Value"]) / 1e3
    total = sum(v[1] for v in agg.values())
    mine = sum(v[1] for k, v in agg.items() if "rk::" in k)
    print(f"launches in the step: {len(step)}, summed device time {total:.1f} us (cold cache, serialised by ncu); "
          f"librank_b200 kernels: {mine:.1f} us = {100 * mine / total:.1f} % of the step\n")
    print("| us | share | launches | grid | block | kernel |\n|---:|---:|---:|---|---|---|")
    for k, (n, t, g, bs) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
        print(f"| {t:.2f} | {100 * t / total:.1f} % | {n} | {g} | {bs} | `{k}` |")


def full_capture(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print(f"\n**`{r[col['Kernel Name']].split('(')[0]}`**\n\n| metric | value |\n|---|---|")
        for m in FULL:
            if m in col:
                print(f"| {m} | {r[col[m]]} {units[col[m]]} |")
        stalls = sorted(((float(r[i]), h.split("issue_stalled_")[1].split("_per_")[0]) for i, h in enumerate(hdr)
                         if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")
                         and "not_issued" not in h), reverse=True)[:5]
        print("| top stalls (warps per issue) | " + ", ".join(f"{n} {v:.2f}" for v, n in stalls) + " |")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launch_list(sys.argv[2], sys.argv[3])
    else:
        full_capture(sys.argv[2])
