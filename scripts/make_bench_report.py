#!/usr/bin/env python
"""profiles/<tag>_bench.md from the JSON lines bench.py printed (profiles/<tag>_bench/*.json).
    python scripts/make_bench_report.py r02
`default.json` = the line of `python bench.py` (primary workload + `other_workloads`), `reference.json` = the
`--impl reference` line, the others = `python bench.py --workload W`."""
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(path):
    text = open(path).read().strip().splitlines()
    text = text[-1] if text else ""
    return json.loads(text) if text.startswith("{") else None


def main(tag):
    rows = {}
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", f"{tag}_bench", "*.json"))):
        d = load(path)
        if d:
            rows[os.path.basename(path)[:-5]] = d
    ref = rows.pop("reference", None)
    default = rows.get("default")
    out = [f"# Round-{int(tag[1:])} bench results (1 x B200)", "",
           "Raw JSON lines: `profiles/%s_bench/`.  Step = zero_grad + forward + loss + backward of the whole model" % tag,
           "(hot path + unchanged torch tower), replayed from a CUDA graph, 256 MiB written between steps to evict L2.",
           "`value`: inputs resident in HBM; `e2e`: packed pinned host inputs, one H2D copy per step (prefetched on a copy",
           "stream, inside the timed brackets) + the D2H read of the loss.  Hot path = a CUDA graph of `model.hot_path`",
           "forward + backward incl. the embedding-gradient reduction, alone, L2 flushed before every replay;",
           "GB/s = algorithmic bytes (SURVEY 8d) / that time; roof = measured copy bandwidth (MEASURED_PEAKS.json).",
           "ATen = the oracle port of the reference module on the same GPU through stock ATen/cuBLAS kernels (eager; its",
           "embedding backward synchronises with the host and cannot be graph-captured).", ""]
    if default:
        d = default
        r = d["roofline"]
        cpu = d.get("cpu_baseline") or {}
        a = d.get("aten_cuda_baseline") or {}
        out += ["## The default line (`python bench.py`): %s" % d["config"]["workload"], "",
                "| ms/step | value samples/s | e2e samples/s | hot path us | alg. GB/s | frac of HBM roof | CPU port samples/s (cores) "
                "| e2e / CPU | ATen-on-CUDA eager ms | step / ATen |",
                "|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|",
                "| %.3f | %.3g | %.3g | %.0f | %.0f | %.3f | %s | %s | %s | %s |" % (
                    d["ms_per_step"], d["value"], d["e2e"]["value"], 1000 * r["hot_ms_per_step"], r["achieved"], r["frac"],
                    "%.3g (%d)" % (cpu["value"], cpu["cores"]) if cpu else "-",
                    "%.0fx" % (d["e2e"]["value"] / cpu["value"]) if cpu else "-",
                    "%.2f" % a["eager_ms_per_step"] if a else "-",
                    "%.1fx" % a["ours_over_aten_eager"] if a else "-"), "",
                "Per-call shares (eager pass behind a spin kernel, CUDA events around each ABI call; upper bounds):", ""]
        out += ["| entry point | us / step | calls |", "|---|---:|---:|"]
        for k, v in (d.get("hotpath_calls_eager_events") or {}).items():
            out.append("| `%s` | %.1f | %.0f |" % (k, 1000 * v["ms_per_step"], v["calls_per_step"]))
        out += ["", "clocks: %s" % json.dumps(d.get("clocks")), ""]
        others = d.get("other_workloads") or {}
        if others:
            out += ["## Every other workload (50 steps each, same run)", "",
                    "| workload | batch | ms/step | value samples/s | e2e samples/s | hot path us | hot share of step | frac of HBM roof "
                    "| launches/step (hot) | ATen eager ms | step / ATen |",
                    "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
            for k, v in others.items():
                if "error" in v:
                    out.append("| %s | error: %s |" % (k, v["error"][:80]))
                    continue
                out.append("| %s | %d | %.3f | %.3g | %.3g | %.0f | %.2f | %.4f | %.0f (%d) | %s | %s |" % (
                    v["workload"], v["batch"], v["ms_per_step"], v["value"], v["e2e_value"], 1000 * v["hot_ms_per_step"],
                    v["hot_share_of_step"], v["roofline_frac"], v["gpu_launches_per_step"], v["hot_launches_per_step"],
                    "%.2f" % v["aten_cuda_eager_ms"] if v.get("aten_cuda_eager_ms") else "-",
                    "%.1fx" % v["ours_over_aten_eager"] if v.get("ours_over_aten_eager") else "-"))
            out.append("")
    singles = {k: v for k, v in rows.items() if k != "default"}
    if singles:
        out += ["## Single-workload runs (`python bench.py --workload W --no-others --steps 100`)", "",
                "| workload | ms/step | value samples/s | e2e samples/s | hot path us | frac of HBM roof | CPU port samples/s (cores) | e2e / CPU |",
                "|---|---:|---:|---:|---:|---:|---:|---:|"]
        for k, d in singles.items():
            r, cpu = d["roofline"], d.get("cpu_baseline")
            out.append("| %s | %.3f | %.3g | %.3g | %.0f | %.4f | %s | %s |" % (
                d["config"]["workload"], d["ms_per_step"], d["value"], d["e2e"]["value"], 1000 * r["hot_ms_per_step"], r["frac"],
                "%.3g (%d)" % (cpu["value"], cpu["cores"]) if cpu else "-",
                "%.0fx" % (d["e2e"]["value"] / cpu["value"]) if cpu else "-"))
        out.append("")
    if ref:
        out += ["## Reference arm (`bench.py --impl reference`)", "",
                "%s: %.3g samples/s on %d host cores (%s)." % (ref["config"]["workload"], ref["value"], ref["cpu_baseline"]["cores"],
                                                             ref["cpu_baseline"]["sample"]), ""]
    open(os.path.join(ROOT, "profiles", f"{tag}_bench.md"), "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r02")
