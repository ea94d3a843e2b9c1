"""CPU tests of the boundary: the C-ABI library loads here (no GPU) and exports every symbol
include/unsflow.h declares; argument validation and the no-CPU-fallback rule."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "unsflow.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(unsflow_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(uf):
    L = uf.lib()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/unsflow.h but not exported"
    assert L.unsflow_version() == 100


def test_ctypes_struct_layout_matches_header(uf):
    cabi = uf._cabi
    # sizes implied by the C declarations (LP64)
    assert C.sizeof(cabi.Vorts) == 8 + 6 * 8
    assert C.sizeof(cabi.Surf) == 8 + 3 * 8 + 8 + 8 + 6 * 8 + 3 * 8 + 12 * 8 + C.sizeof(cabi.Vorts)
    assert C.sizeof(cabi.Field) == 16 + 3 * C.sizeof(cabi.Vorts)
    assert C.sizeof(cabi.Opts) == 4 * 4 + 8 + 8 + 8 + 8 + 8


def test_argument_validation_needs_no_gpu(uf):
    L = uf.lib()
    assert L.unsflow_ind_vel(-1, None, None, None, None, 0, None, None, None, None) != 0
    assert b"negative" in L.unsflow_last_error()
    assert L.unsflow_wake_create(None, 10, 0, 0, 1, None, 0) != 0
    assert L.unsflow_ldvm_run(None, 1, None) != 0


def test_no_cpu_fallback_without_device(uf):
    """On a box without a GPU every compute entry must fail loudly, not compute on the CPU."""
    if uf.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(uf.UnsflowError, match="no usable CUDA device"):
        uf.ind_vel([uf.TwoDVort(0.0, 0.0, 1.0, 0.02)], np.array([1.0]), np.array([0.0]))
    surf = uf.TwoDSurf("FlatPlate", 0.25, uf.KinemDef(uf.ConstDef(0.1), uf.ConstDef(0.0), uf.ConstDef(1.0)), [10.0])
    with pytest.raises(uf.UnsflowError):
        uf.ldvm(surf, uf.TwoDFlowField(), 3, 0.015, write_summary=False)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "unsteadyflowsolvers.jl_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in txt.lower(), f"{f} mentions the oracle"


def test_bench_reference_arm_runs_on_cpu():
    """`bench.py --impl reference` (the reference CPU path = oracle port on the host cores) prints
    exactly one JSON line on stdout with the contract keys; tiny sample so it takes < 2 s"""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n", "20000",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["value"] > 1e7 and d["unit"] == "interactions/s"
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
