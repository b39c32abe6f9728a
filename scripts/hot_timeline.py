#!/usr/bin/env python
"""Device timeline of ONE replay of a workload's hot-path CUDA graph (development aid).

    python scripts/hot_timeline.py din_tc [--full]      # on the GPU box

Builds the workload exactly as bench.py does, captures the hot-path-only graph (bench.Stepper, hot_only) or the
whole step (--full), flushes L2, replays it under torch.profiler (CUPTI kernel activities) and prints every
kernel with its start offset, duration and stream: which launches overlap, where the gaps are, what the critical
path is.  Timestamps under a profiler are not bench numbers; the shares and the ordering are what this is for."""
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402


def main(workload="din_tc", full=False):
    import bench
    import rank_b200
    from rank_b200.staging import PackedBatch
    from torch.profiler import ProfilerActivity, profile

    wl = bench.WORKLOADS[workload]()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    vocab = rank_b200.write_vocab_dir(tempfile.mkdtemp(prefix="rk_vocab_")) + "/"
    torch.manual_seed(0)
    model = wl.model(rank_b200, False, vocab).to(dev).train()
    raw = wl.make_batch(wl.batch, 1000)
    pb = PackedBatch.like(raw, dev).fill(raw)
    pb.to_device()
    torch.cuda.synchronize()
    st = bench.Stepper(model, wl, pb, True, None, hot_only=not full)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for i in range(5):
        flush.fill_(i)
        st.run(i)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for i in range(3):
            flush.fill_(i)
            torch.cuda.synchronize()
            st.run(10 + i)
            torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "Memset" not in e.name]
    evs.sort(key=lambda e: e.time_range.start)
    # split into replays at the big flush kernels
    runs, cur = [], []
    for e in evs:
        if "FillFunctor<unsigned char>" in e.name or "fill" in e.name.lower() and e.time_range.elapsed_us() > 40:
            if cur:
                runs.append(cur)
            cur = []
            continue
        cur.append(e)
    if cur:
        runs.append(cur)
    run = runs[-1]
    t0 = run[0].time_range.start
    end = max(e.time_range.end for e in run)
    print("%s %s: %d kernels, first start -> last end = %.1f us" % (workload, "step" if full else "hot path", len(run), end - t0))
    print("%9s %9s %8s  %-6s %s" % ("start us", "end us", "dur us", "stream", "kernel"))
    for e in run:
        s, d = e.time_range.start - t0, e.time_range.elapsed_us()
        stream = getattr(e, "device_resource_id", getattr(e, "device_index", 0))
        print("%9.1f %9.1f %8.1f  %-6s %s" % (s, s + d, d, stream, e.name[:90]))


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    main(args[0] if args else "din_tc", "--full" in sys.argv)
