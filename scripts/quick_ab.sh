#!/bin/bash
# One GPU call: the model parity tests of the kernels touched + one 100-step bench line per workload given.
#   bash scripts/quick_ab.sh <tag> <pytest -k expr> <workload> [<workload> ...]
tag=$1; kexpr=$2; shift 2
mkdir -p gpurun_out/$tag
timeout 420 python -m pytest tests/test_gpu_models.py tests/test_gpu_edges.py -q -m gpu -k "$kexpr" -x 2>&1 | grep -v arbiter | tail -5
for w in "$@"; do
  timeout 200 python bench.py --workload $w --no-others --no-aten --no-cpu-baseline --steps 100 2>/dev/null | tail -1 > gpurun_out/$tag/$w.json
  python - <<PY
import json
d=json.loads(open("gpurun_out/$tag/$w.json").read())
r=d["roofline"]
print("$w", "step %.3f ms  value %.3g  e2e %.3g  hot %.1f us  frac %.4f" % (d["ms_per_step"], d["value"], d["e2e"]["value"], 1000*r["hot_ms_per_step"], r["frac"]))
for k,v in (d.get("hotpath_calls_eager_events") or {}).items(): print("    %-28s %.1f us" % (k, 1000*v["ms_per_step"]))
PY
done
