"""Phase breakdown of tc::din_fwd_tc_kernel (development aid, not part of the product).

Builds a second copy of the library with -DRK_DIN_PROFILE into scripts/_prof/ (clock64 deltas of
thread 0 accumulated per phase), runs the DIN forward a few times and prints the share of each phase.
    python scripts/din_tc_phase_profile.py build     # here (nvcc)
    python scripts/din_tc_phase_profile.py run       # on the GPU box
"""
import ctypes, subprocess, sys, tempfile
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
PROF = ROOT / "scripts" / "_prof" / "librank_b200_prof.so"
PHASES = ["setup", "group load", "wait rows", "build A1", "mma1+prefetch", "epilogue1", "mma2", "epilogue2",
          "weights", "pooling", "assembly", "TOTAL", "tiles", "groups"]

def build():
    pkg = next(ROOT.glob("*_b200"))
    PROF.parent.mkdir(exist_ok=True)
    srcs = sorted((pkg / "csrc").glob("*.cu"))
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
           "--expt-relaxed-constexpr", "-DRK_DIN_PROFILE", "-I", str(ROOT / "include"), "-I", str(pkg / "csrc"),
           "-shared", "-o", str(PROF), *map(str, srcs), "--cudart", "static"]
    subprocess.run(cmd, check=True)

def run():
    import torch
    import rank_b200
    from rank_b200 import _lib, synthetic
    _lib.LIB_PATH = PROF
    lib = _lib.load()
    import bench
    wl = bench.DINTensorCoreWorkload()
    dev = torch.device("cuda", 0)
    vocab = rank_b200.write_vocab_dir(tempfile.mkdtemp(prefix="rk_vocab_")) + "/"
    torch.manual_seed(0)
    model = wl.model(rank_b200, False, vocab).to(dev).train()
    batch = synthetic.to_device(wl.make_batch(8192, 1000), dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    fn = lib.rk_debug_din_profile
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
    out = (ctypes.c_ulonglong * 16)()
    for i in range(3):
        wl.loss(model, batch)
    fn(out, 1)
    n = 10
    for i in range(n):
        flush.zero_()
        with torch.no_grad():
            pass
        wl.loss(model, batch)
    fn(out, 1)
    v = [x / n for x in out]
    tot = v[11]
    print(f"per launch: {v[13]:.0f} groups, {v[12]:.0f} tiles, CTA-cycles total {tot:.3e}")
    for name, x in zip(PHASES[:11], v[:11]):
        print(f"  {name:14s} {100 * x / tot:5.1f} %   {x / max(v[12], 1):8.0f} cyc/tile")

if __name__ == "__main__":
    {"build": build, "run": run}[sys.argv[1]]()
