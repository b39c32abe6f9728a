"""The C-ABI library loads and exports every symbol include/rank_b200.h declares (CPU only)."""
import ctypes
import os
import re

import pytest

import rank_b200
from rank_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "rank_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rk_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    names = declared_symbols()
    assert names, "no declarations parsed"
    assert sorted(_lib.PROTOTYPES) == names


def test_library_exports_every_declared_symbol():
    assert _lib.library_path().exists(), "librank_b200.so not built (run __graft_entry__.build())"
    lib = ctypes.CDLL(str(_lib.library_path()))
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert _lib.load().rk_version() == _lib.ABI_VERSION


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_lib.RkField) == 32
    assert ctypes.sizeof(_lib.RkGradTable) == 32
    assert ctypes.sizeof(_lib.RkDirectTable) == 56       # rk_direct_table_t
    assert ctypes.sizeof(_lib.RkBstBlock) == 18 * 8 + 8  # rk_bst_block_t


def test_workspace_queries_need_no_gpu():
    lib = _lib.load()
    assert lib.rk_plan_workspace_bytes(0) > 0
    assert lib.rk_plan_workspace_bytes(8192 * 57) >= 2 * 4 * 8192 * 57


def test_no_cpu_fallback():
    """CPU tensors are rejected, not silently computed somewhere else."""
    import torch
    from rank_b200.sparse import gather_concat
    w = torch.zeros(4, 2)
    i = torch.zeros(3, dtype=torch.int64)
    with pytest.raises(rank_b200.RankB200Error):
        gather_concat([w], [i], [0])


def test_argument_counts_match_the_header():
    """Every ctypes prototype has as many arguments as the C declaration (catches binding drift)."""
    text = open(os.path.join(ROOT, "include", "rank_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    decls = dict(re.findall(r"\b(rk_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S))
    assert set(decls) == set(_lib.PROTOTYPES)
    for name, params in decls.items():
        params = " ".join(params.split())
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(_lib.PROTOTYPES[name][1]), (name, n, len(_lib.PROTOTYPES[name][1]))
