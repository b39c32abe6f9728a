"""Per-kernel totals of an ncu `--metrics gpu__time_duration.sum --csv` launch list.
usage: launch_summary.py launches.csv [n_steps]   (n_steps divides the totals: per-step figures)"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    n_steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    hdr, agg = None, collections.OrderedDict()
    for r in csv.reader(open(path, errors="replace")):
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(d["Metric Unit"], 1.0)
        a = agg.setdefault(d["Kernel Name"], [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    print(f"{path}: {sum(a[0] for a in agg.values())} launches, {total / n_steps:.1f} us per step")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:45]:
        print(f"  {t / n_steps:9.1f} us/step {100 * t / total:5.1f}%  avg {t / n:8.1f} us x{n:5d}  {k[:90]}")


if __name__ == "__main__":
    main()
