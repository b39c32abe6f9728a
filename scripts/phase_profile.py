"""Phase breakdown of an instrumented kernel (development aid, not part of the product).

Builds a second copy of the library with -DRK_PROFILE into scripts/_prof/ (csrc/prof.cuh: clock64
deltas of thread 0 accumulated per phase), runs a workload's step a few times and prints the share
of each phase.  Only ONE instrumented kernel may run per measurement (they share the counters):
    python scripts/phase_profile.py build                 # here (nvcc)
    python scripts/phase_profile.py run din_tc fwd        # on the GPU box: DIN tensor-core forward
    python scripts/phase_profile.py run afm fwd|bwd
"""
import ctypes, subprocess, sys, tempfile
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
PROF = ROOT / "scripts" / "_prof" / "librank_b200_prof.so"
PHASES = {
    ("din_tc", "fwd"): ["setup", "group load", "wait rows", "build A1", "mma1+prefetch", "epilogue1", "mma2", "epilogue2",
                        "weights", "pooling", "assembly"],
    ("bst", "bwd"): ["setup", "A1 recompute q/k/v", "A2 attention+ffn fwd", "B ln2 bwd", "C ffn bwd", "D ln1+wo bwd",
                     "E attention bwd", "F1 dq/dk/dv outer", "F2 input grad", "F3 pos grad", "final"],
    ("bst_tc", "bwd"): ["tile setup", "A1 x -> q/k/v round trip", "A2 attention fwd", "A3 Wo, W1, W2 round trips + LN",
                        "B ln2 bwd, W2^T, W1^T + stage A", "C ln1 bwd, Wo^T + stage B", "D attention bwd",
                        "E Wq/k/v^T + stage C, dx, pos grad", "exit"],
    ("afm", "fwd"): ["setup", "wait rows", "build A1", "mma1+prefetch", "epilogue", "softmax", "pool store", "pool sum"],
    ("afm", "bwd"): ["setup", "wait rows", "build A1", "mma1+prefetch", "epilogue1+mask", "softmax+g_s", "X line",
                        "mma2+mma3", "epilogue2", "g_rows", "final"],
}

def build():
    pkg = next(ROOT.glob("*_b200"))
    PROF.parent.mkdir(exist_ok=True)
    srcs = sorted((pkg / "csrc").glob("*.cu"))
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
           "--expt-relaxed-constexpr", "-DRK_PROFILE", "-rdc=true", "-I", str(ROOT / "include"), "-I", str(pkg / "csrc"),
           "-shared", "-o", str(PROF), *map(str, srcs), "--cudart", "static"]
    subprocess.run(cmd, check=True)

def run(workload="din_tc", which="fwd"):
    import torch
    import rank_b200
    from rank_b200 import _lib, synthetic
    _lib.LIB_PATH = PROF
    lib = _lib.load()
    import bench
    wl = bench.WORKLOADS[workload]()
    dev = torch.device("cuda", 0)
    vocab = rank_b200.write_vocab_dir(tempfile.mkdtemp(prefix="rk_vocab_")) + "/"
    torch.manual_seed(0)
    model = wl.model(rank_b200, False, vocab).to(dev).train()
    batch = synthetic.to_device(wl.make_batch(wl.batch, 1000), dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    fn = lib.rk_debug_profile
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
    out = (ctypes.c_ulonglong * 16)()
    n = 10

    def step(i):
        flush.zero_()
        loss = wl.loss(model, batch)
        if which == "bwd":
            fn(out, 1)                     # drop what the (instrumented) forward recorded
            loss.backward()
            fn(out, 0)
            return [x for x in out]
        return None

    for i in range(3):
        wl.loss(model, batch).backward()
    fn(out, 1)
    acc = [0] * 16
    for i in range(n):
        got = step(i)
        if got is None:
            fn(out, 1)
            got = [x for x in out]
        acc = [a + g for a, g in zip(acc, got)]
        fn(out, 1)
    v = [x / n for x in acc]
    tot = v[11]
    names = PHASES[(workload, which)]
    print(f"{workload} {which}: per launch {v[13]:.0f} groups, {v[12]:.0f} tiles, CTA-cycles total {tot:.3e}")
    for name, x in zip(names, v):
        print(f"  {name:16s} {100 * x / tot:5.1f} %   {x / max(v[12], 1):8.0f} cyc/tile")


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build()
    else:
        run(*sys.argv[2:4])
