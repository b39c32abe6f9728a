#!/usr/bin/env python
"""profiles/<tag>_scaling.md from gpurun_out/<tag>_scale/*.json (scripts/scale_run.sh N) — copies the JSON lines to
profiles/<tag>_scale/ as well."""
import glob
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(tag):
    src = os.path.join(ROOT, "gpurun_out", f"{tag}_scale")
    dst = os.path.join(ROOT, "profiles", f"{tag}_scale")
    os.makedirs(dst, exist_ok=True)
    rows, sharded = {}, {}
    for path in sorted(glob.glob(os.path.join(src, "*.json"))):
        text = open(path).read().strip()
        if not text.startswith("{"):
            continue
        shutil.copy(path, dst)
        d = json.loads(text.splitlines()[-1])
        name = os.path.basename(path)[:-5]
        if name.startswith("sharded"):
            sharded[d["n_gpus"]] = d
        else:
            w, n = name.rsplit("_n", 1)
            rows.setdefault(w, {})[int(n)] = d
    out = [f"# Round-{int(tag[1:])} multi-GPU runs (one box, `scripts/scale_run.sh N` under `gpurun --gpus N`)", "",
           "Weak scaling: the per-GPU batch is fixed (8192; DeepFM 1024), one process per GPU under torchrun, every rank holds",
           "all tables, the gradient all-reduce (NCCL over NVLink/NVSwitch, `ReduceOp.AVG`, one call) is captured in the step's",
           "CUDA graph.  `value` = global samples per second from the max over ranks of the device-timed step; efficiency =",
           "value(N) / (N x value(1)).", ""]
    for w, per in rows.items():
        if 1 not in per:
            continue
        base = per[1]["value"]
        out += [f"## {per[1]['config']['workload']}", "",
                "| GPUs | ms/step | value samples/s | efficiency | e2e samples/s | e2e efficiency | + ms vs 1 GPU |", "|---:|---:|---:|---:|---:|---:|---:|"]
        for n in sorted(per):
            d = per[n]
            out.append("| %d | %.3f | %.4g | %.3f | %.4g | %.3f | %+.3f |" % (
                n, d["ms_per_step"], d["value"], d["value"] / (n * base), d["e2e"]["value"],
                d["e2e"]["value"] / (n * per[1]["e2e"]["value"]), d["ms_per_step"] - per[1]["ms_per_step"]))
        out.append("")
    if sharded:
        out += ["## BST with the feedid table row-sharded over the ranks (`scripts/sharded_bst.py`, 12.5 M rows x 16 per rank)", "",
                "Step = zero_grad + forward (owner/route kernels, all-to-all of indices, local gather, all-to-all of rows, block, tower)",
                "+ loss + backward (all-to-all of the row gradients, owner-side sorted reduction into touched rows, dense all-reduce of the",
                "replicated parameters) + fused Adam on the dense parameters + RowwiseAdam on the shard's touched rows; per-GPU batch 1024.",
                "The fixed-capacity exchange has no host synchronisation, so forward + backward + all collectives replay from one CUDA graph.", "",
                "| GPUs | table rows | shard GB | ms/step | samples/s | launch | parity vs replicated (logit / shard grad / other grads) |",
                "|---:|---:|---:|---:|---:|---|---|"]
        for n in sorted(sharded):
            d = sharded[n]
            p = d["parity_vs_replicated"]
            out.append("| %d | %.3g | %.2f | %.3f | %.4g | %s | %s: %.1e / %.1e / %.1e |" % (
                n, d["table_rows"], d["shard_gb"], d["ms_per_step"], d["samples_per_s"], d.get("launch", "eager"),
                "ok" if p["ok_all_ranks"] else "FAILED", p["logit_rel_err"], p["shard_grad_rel_err"], p["other_grad_rel_err"]))
        out.append("")
    notes = os.path.join(ROOT, "profiles", f"{tag}_scaling_notes.md")
    if os.path.exists(notes):
        out += [open(notes).read().rstrip(), ""]
    path = os.path.join(ROOT, "profiles", f"{tag}_scaling.md")
    open(path, "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r02")
