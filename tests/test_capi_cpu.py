"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/hcore_b200.h declares,
and refuses (loudly) to compute without a CUDA device -- there is no CPU fallback to fall into."""
import ctypes as C
import os
import re

import pytest
import torch

import hcorepp_b200
from hcorepp_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    names = _capi.declared_symbols()
    assert len(names) > 50
    missing = [n for n in names if not hasattr(_capi.lib, n)]
    assert not missing, missing
    for must in ("hcb_dtlr_gemm_batched", "hcb_stlr_gemm_batched", "hcb_dcompress_batched", "hcb_dtlr_matmul",
                 "hcb_dgeqrf", "hcb_dsvd", "hcb_dunmqr", "hcb_dprocess_v", "hcb_ctx_create", "hcb_malloc"):
        assert must in names


def test_struct_layouts_match_header():
    assert C.sizeof(_capi.hcb_tile) == 56
    assert _capi.hcb_tile.d_rank.offset == 24 and _capi.hcb_tile.d_data.offset == 32
    assert _capi.hcb_tile.d_state.offset == 40 and _capi.hcb_tile.fixed_rank.offset == 48
    assert C.sizeof(_capi.hcb_compress_params) == 40


def test_workspace_formula_is_monotone():
    w1 = _capi.lib.hcb_dtlr_gemm_workspace(16, 1024, 1024, 1024, 64)
    w2 = _capi.lib.hcb_dtlr_gemm_workspace(16, 1024, 1024, 1024, 128)
    w3 = _capi.lib.hcb_dtlr_gemm_workspace(32, 1024, 1024, 1024, 128)
    assert 0 < w1 < w2 < w3


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback():
    h = C.c_void_p()
    rc = _capi.lib.hcb_ctx_create(0, C.byref(h))
    assert rc == 3  # HCB_ENODEVICE
    assert b"no CPU fallback" in _capi.lib.hcb_last_error()
    with pytest.raises(_capi.HcbError):
        hcorepp_b200.RunContext(0)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under hcorepp_b200/ may import, link or execute it."""
    pkg = os.path.join(ROOT, "hcorepp_b200")
    for root in (pkg, os.path.join(ROOT, "include")):
        for dirpath, _, files in os.walk(root):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")) or f == "Makefile":
                    text = open(os.path.join(dirpath, f), errors="ignore").read()
                    assert not re.search(r"^\s*(from|import)\s+oracle", text, re.M), f
                    assert "libhcorepp_ref" not in text and "tlr_oracle" not in text and "oracle/" not in text, f
                    assert "scipy" not in text, f


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the reference's CPU path, no GPU involved) prints exactly one JSON line on stdout with
    the keys the driver reads."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not os.path.exists(os.path.join(root, "oracle", "_ref", "libhcorepp_ref.so")):
        import pytest
        pytest.skip("oracle/_ref not built")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--tiles", "2", "--nb", "256",
                          "--steps", "1", "--warmup", "1", "--cpu-budget-s", "5"], capture_output=True, text=True, timeout=600,
                         cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "tile-GEMM/s" and d["value"] > 0 and d["higher_is_better"] is True
    for key in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
