#!/bin/bash
# usage: scripts/bench_calls.sh <workload> [extra bench args]  — prints ms/step and per-call device times
w=$1; shift
python bench.py --workload "$w" --steps 20 --warmup 5 --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c '
import sys, json
d = json.loads(sys.stdin.read())
print("ms_per_step", round(d["ms_per_step"], 4), "e2e", d["e2e"]["value"] if d.get("e2e") else None)
for k, v in (d.get("roofline", {}).get("hotpath_calls") or d.get("hotpath_calls") or {}).items():
    print("  %-28s %6.1f us x %g" % (k, 1000 * v["ms_per_step"], v["calls_per_step"]))
'
