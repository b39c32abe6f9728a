#!/usr/bin/env python
"""Attribute an ncu source-page export (SASS rows) to CUDA source lines through nvdisasm's line table.
    ncu -i X.ncu-rep --page source --csv --kernel-name regex:K > k.csv
    cuobjdump -xelf all librank_b200.so; nvdisasm -g -c file.cubin > file.sass
    python scripts/ncu_by_line.py k.csv file.sass <mangled-substring> source.cu [top]
Prints, per source line: stall samples, warp instructions executed, shared-memory wavefronts (excess)."""
import csv
import re
import sys
from collections import defaultdict


def main(csv_path, sass_path, kernel, src_path, top=40):
    rows = list(csv.reader(open(csv_path)))
    hdr = rows[1]
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    # the export repeats the table once per view; keep the first copy
    first = data[0][0]
    for i in range(1, len(data)):
        if data[i][0] == first:
            data = data[:i]
            break
    ix = {h: i for i, h in enumerate(hdr)}
    num = lambda r, k: float(r[ix[k]]) if r[ix[k]] not in ("", "-") else 0.0
    lines, cur, inside = [], ("", 0), False
    for ln in open(sass_path):
        if ln.startswith(".text."):
            inside = kernel in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "(.*)", line (\d+)', ln)
        if m:
            cur = (m.group(1), int(m.group(2)))
        elif re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            lines.append(cur)
    if abs(len(lines) - len(data)) > 1:              # ncu appends one row after the last instruction
        print("instruction count mismatch: sass %d vs ncu %d" % (len(lines), len(data)))
    n = min(len(lines), len(data))
    agg = defaultdict(lambda: [0.0, 0.0, 0.0, 0.0, defaultdict(float)])
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not" not in h]
    for i in range(n):
        a = agg[lines[i]]
        a[0] += num(data[i], "# Samples")
        a[1] += num(data[i], "Instructions Executed")
        a[2] += num(data[i], "L1 Wavefronts Shared")
        a[3] += num(data[i], "L1 Wavefronts Shared Excessive")
        for h in stall_cols:
            v = num(data[i], h)
            if v:
                a[4][h[6:]] += v
    cache = {}

    def text_of(key):
        path, line = key
        if path not in cache:
            try:
                cache[path] = open(src_path if path.endswith(src_path.split("/")[-1]) else path).read().splitlines()
            except OSError:
                cache[path] = []
        src = cache[path]
        return src[line - 1].strip()[:80] if 0 < line <= len(src) else "?"
    tot_s = sum(a[0] for a in agg.values())
    tot_i = sum(a[1] for a in agg.values())
    print("samples %d, warp instructions %d, shared wavefronts %d (excess %d)" % (
        tot_s, tot_i, sum(a[2] for a in agg.values()), sum(a[3] for a in agg.values())))
    print("%-16s %7s %6s %9s %6s %8s %8s  %s" % ("line", "samples", "%", "instr", "%", "smem wf", "excess", "source / top stalls"))
    for line, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        st = ", ".join("%s %d" % kv for kv in sorted(a[4].items(), key=lambda kv: -kv[1])[:3])
        text = text_of(line)
        line = "%s:%d" % (line[0].split("/")[-1], line[1])
        print("%-16s %7d %5.1f%% %9d %5.1f%% %8d %8d  %s   [%s]" % (
            line, a[0], 100 * a[0] / max(tot_s, 1), a[1], 100 * a[1] / max(tot_i, 1), a[2], a[3], text, st))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]) if len(sys.argv) > 5 else 40)
