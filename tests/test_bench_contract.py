"""bench.py's reference arm runs on the CPU: check the one-JSON-line contract and its keys here."""
import json
import os
import subprocess
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "fwfm",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.strip().splitlines()
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    assert d["config"]["workload"] and "model" not in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["value"] > 0 and d["warmup"] >= 3


def test_tower_runs_the_modules_on_cpu():
    """No CUDA tensor, no fused kernel: run_tower is exactly the reference's layer loop."""
    import rank_b200
    from rank_b200 import tower
    torch.manual_seed(0)
    layers = nn.ModuleList([nn.Linear(6, 8), rank_b200.Dice(8), nn.BatchNorm1d(8), nn.Linear(8, 4), nn.BatchNorm1d(4),
                            nn.ReLU(), nn.Linear(4, 1)]).train()
    x = torch.randn(512, 6)
    a = tower.run_tower(layers, x)
    torch.manual_seed(0)
    b = x
    for m in layers:
        b = m(b)
    # the second pass updated the running statistics again, the outputs of a training-mode pass do not depend on them
    assert torch.equal(a, b)
