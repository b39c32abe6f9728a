#!/usr/bin/env python
"""profiles/<tag>_bench.md from the JSON lines bench.py printed (profiles/<tag>_bench/*.json).
    python scripts/make_bench_report.py r01
"""
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORDER = ["dcn", "deepfm", "fwfm", "afm", "afm_fp32", "din", "din_softmax", "din_tc", "din_softmax_tc", "bst", "deepcrossing"]


def main(tag):
    rows = {}
    for path in glob.glob(os.path.join(ROOT, "profiles", f"{tag}_bench", "*.json")):
        text = open(path).read().strip()
        if text.startswith("{"):
            rows[os.path.basename(path)[:-5]] = json.loads(text)
    ref = rows.pop("reference_dcn", None)
    keys = [k for k in ORDER if k in rows] + sorted(k for k in rows if k not in ORDER)
    out = [f"# Round-{int(tag[1:])} bench results (1 x B200, `python bench.py --workload W --steps 30 --warmup 5`)", "",
           "Raw JSON lines: `profiles/%s_bench/`.  Step = zero_grad + forward + loss + backward of the whole model" % tag,
           "(hot path + unchanged torch tower), replayed from a CUDA graph, 256 MiB written between steps to evict L2.",
           "`value`: inputs resident in HBM; `e2e`: packed pinned host inputs, ONE H2D copy + the D2H read of the loss",
           "inside the timed region.  CPU port = the oracle restatement of the reference model on the box's host cores",
           "(fwd+loss+bwd, same batch).  Hot path = summed device time of the librank_b200 calls of a step (spin-queued",
           "eager pass, CUDA events); GB/s = algorithmic bytes (SURVEY 8d) / that time; roof = measured copy bandwidth.", "",
           "| workload | batch | ms/step | value samples/s | e2e samples/s | CPU port samples/s (cores) | e2e / CPU | hot path us "
           "| hot path us (warm L2) | alg. GB/s | frac of HBM roof | librank launches/step |",
           "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
    for k in keys:
        d = rows[k]
        r, cpu = d["roofline"], d.get("cpu_baseline")
        out.append("| %s | %d | %.3f | %.3g | %.3g | %s | %s | %.0f | %.0f | %.0f | %.3f | %d |" % (
            d["config"]["workload"], d["config"].get("batch_per_gpu", d["config"].get("batch", 0)), d["ms_per_step"],
            d["value"], d["e2e"]["value"],
            "%.3g (%d)" % (cpu["value"], cpu["cores"]) if cpu else "-",
            "%.0fx" % (d["e2e"]["value"] / cpu["value"]) if cpu else "-",
            1000 * r["hot_ms_per_step"], 1000 * r["hot_ms_per_step_warm_l2"], r["achieved"], r["frac"],
            d["gpu_launches"] / d["steps"]))
    out += ["", "## Per-call device time of the library calls of a step (us per step, L2 flushed before every step)", "",
            "forward / backward = the fused interaction kernels; tower = the fused Dice / BatchNorm layers of the DNN tower",
            "(`rk_dice_bn_*`, `rk_bn_act_*`: outside the hot-path roofline, inside the step).", "",
            "| workload | " + " | ".join(["forward", "backward", "rk_plan_build", "rk_embgrad_segment_reduce", "other hot path",
                                         "tower fwd", "tower bwd"]) + " |",
            "|---|---:|---:|---:|---:|---:|---:|---:|"]
    for k in keys:
        calls = {n: 1000 * v["ms_per_step"] for n, v in rows[k]["hotpath_calls"].items()}
        tower = lambda n: n.startswith("rk_dice_bn") or n.startswith("rk_bn_act")
        fwd = sum(v for n, v in calls.items() if n.endswith("_fwd") and not tower(n))
        bwd = sum(v for n, v in calls.items() if n.endswith("_bwd") and not tower(n))
        tf = sum(v for n, v in calls.items() if n.endswith("_fwd") and tower(n))
        tb = sum(v for n, v in calls.items() if n.endswith("_bwd") and tower(n))
        plan, seg = calls.get("rk_plan_build", 0.0), calls.get("rk_embgrad_segment_reduce", 0.0)
        other = max(0.0, sum(calls.values()) - fwd - bwd - plan - seg - tf - tb)
        out.append("| %s | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f |" % (
            rows[k]["config"]["workload"], fwd, bwd, plan, seg, other, tf, tb))
    if ref:
        out += ["", "## Reference arm (`bench.py --impl reference --workload dcn`)", "",
                "%.3g samples/s on %d host cores (%s)." % (ref["value"], ref["cpu_baseline"]["cores"], ref["cpu_baseline"]["sample"])]
    clocks = {k: rows[k].get("clocks") for k in keys}
    out += ["", "## Clocks during the timed regions", "",
            ", ".join("%s: %s MHz%s" % (k, c.get("sm_mhz"), (" " + str(c["reasons"])) if c.get("reasons") else "")
                      for k, c in clocks.items() if c)]
    open(os.path.join(ROOT, "profiles", f"{tag}_bench.md"), "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r01")
