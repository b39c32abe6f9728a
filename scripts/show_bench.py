"""Prints the interesting parts of a bench.py JSON line (gpurun_out/*.json)."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "steps", "gpu_launches")}, "e2e", round(d["e2e"]["value"]))
r = d["roofline"]
print({k: (round(r[k], 4) if isinstance(r[k], float) else r[k]) for k in r if k not in ("what", "traffic_note", "traffic_by_kernel")})
a = d.get("aten_cuda_baseline") or {}
print("aten", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in a.items() if k != "what"})
for k, v in (d.get("hotpath_calls_eager_events") or {}).items():
    print("  ", k, round(v["ms_per_step"] * 1e3, 1), "us x", v["calls_per_step"])
print(d.get("cpu_baseline"))
for k, v in (d.get("other_workloads") or {}).items():
    if "error" in v:
        print(k, v)
        continue
    print(f"{k:15s} ms {v['ms_per_step']:.3f} hot {v['hot_ms_per_step']*1e3:6.1f} us frac {v['roofline_frac']:.4f} launches {v['hot_launches_per_step']}/{v['gpu_launches_per_step']:.0f} "
          f"aten eager {v['aten_cuda_eager_ms']:.2f} graph {v['aten_cuda_graph_ms']} x{v['ours_over_aten_eager']:.1f}")
