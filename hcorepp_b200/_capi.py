"""ctypes binding of libhcore_b200.so (include/hcore_b200.h).

The CUDA library is the product: if it is missing this module raises at import time -- there is no Python, NumPy or
CPU fallback behind it (and nothing here ever imports the test-only checker package).
"""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libhcore_b200.so")
HEADER = os.path.join(os.path.dirname(_HERE), "include", "hcore_b200.h")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(nvcc, sm_100a). hcorepp_b200 has no CPU fallback.")

lib = C.CDLL(LIB_PATH)

i32, i64, vp, sz = C.c_int32, C.c_int64, C.c_void_p, C.c_size_t


class hcb_tile(C.Structure):
    """Mirror of `struct hcb_tile` (include/hcore_b200.h) == operators::TileMetadata + buffer (Tile.hpp:30-52)."""
    _fields_ = [("type", i32), ("m", i32), ("n", i32), ("ld", i32), ("max_rank", i32), ("rank_bound", i32),
                ("d_rank", vp), ("d_data", vp), ("d_state", vp), ("fixed_rank", i32), ("reserved", i32)]


class hcb_compress_params(C.Structure):
    """Mirror of `struct hcb_compress_params` == operators::CompressionParameters (CompressionParameters.hpp:44-46)."""
    _fields_ = [("accuracy", C.c_double), ("use_trmm", i32), ("use_ungqr", i32), ("truncated_svd", i32),
                ("fixed_rank", i64), ("svd_type", i32), ("reserved", i32)]


TILE_DENSE, TILE_COMPRESSED = 0, 1
STATE_ORTHO_U = 1
EBOUND = 6

# ---- context / memory -------------------------------------------------------------------------------------------
lib.hcb_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
lib.hcb_ctx_create_on_stream.argtypes = [C.c_int, vp, C.POINTER(vp)]
lib.hcb_ctx_destroy.argtypes = [vp]
lib.hcb_ctx_sync.argtypes = [vp]
lib.hcb_ctx_stream.argtypes = [vp]
lib.hcb_ctx_stream.restype = vp
lib.hcb_ctx_device.argtypes = [vp]
lib.hcb_ctx_sm_count.argtypes = [vp]
lib.hcb_ctx_reserve_workspace.argtypes = [vp, sz]
lib.hcb_ctx_workspace_bytes.argtypes = [vp]
lib.hcb_ctx_workspace_bytes.restype = sz
lib.hcb_ctx_stats.argtypes = [vp, C.POINTER(C.c_uint64), C.c_int]
lib.hcb_malloc.argtypes = [vp, sz, C.POINTER(vp)]
lib.hcb_free.argtypes = [vp, vp]
lib.hcb_memcpy.argtypes = [vp, vp, vp, sz, C.c_int]
lib.hcb_memset.argtypes = [vp, vp, C.c_int, sz]
lib.hcb_ctx_phase_timing.argtypes = [vp, C.c_int]
lib.hcb_ctx_phase_times.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
lib.hcb_phase_name.argtypes = [C.c_int]
lib.hcb_phase_name.restype = C.c_char_p
N_PHASES = 9
lib.hcb_last_error.restype = C.c_char_p
lib.hcb_version.restype = C.c_char_p
lib.hcb_launch_count.restype = C.c_uint64
lib.hcb_launch_count_reset.restype = None

_PT = C.POINTER(hcb_tile)
_PP = C.POINTER(hcb_compress_params)


def _declare(p, ct):
    f = lambda name: getattr(lib, f"hcb_{p}{name}")
    f("gemm").argtypes = [vp, C.c_int, C.c_int, i64, i64, i64, ct, vp, i64, vp, i64, ct, vp, i64]
    f("multiply_by_alpha").argtypes = [vp, vp, i64, i64, i64, i64, ct]
    f("process_v").argtypes = [vp, i64, i64, C.c_int, i64, ct, vp, i64, vp, i64, vp, C.c_int]
    f("new_rank").argtypes = [vp, C.c_int, vp, i64, ct, C.POINTER(i64)]
    f("new_rank_device").argtypes = [vp, C.c_int, vp, i64, ct, vp]
    f("uvptr").argtypes = [vp, i64, i64, vp, vp]
    f("vtnew").argtypes = [vp, i64, C.c_int, i64, vp, vp, i64, i64]
    f("uvptr_conj").argtypes = [vp, i64, i64, vp]
    f("fill_identity").argtypes = [vp, i64, vp]
    f("lacpy").argtypes = [vp, C.c_int, i64, i64, vp, i64, vp, i64]
    f("laset").argtypes = [vp, C.c_int, i64, i64, ct, ct, vp, i64]
    f("geqrf").argtypes = [vp, i64, i64, vp, i64, vp]
    f("ungqr").argtypes = [vp, i64, i64, i64, vp, i64, vp]
    f("unmqr").argtypes = [vp, C.c_int, C.c_int, i64, i64, i64, vp, i64, vp, vp, i64]
    f("svd").argtypes = [vp, i64, i64, vp, i64, vp, vp, i64, vp, i64]
    f("trmm").argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, i64, i64, ct, vp, i64, vp, i64]
    f("tlr_gemm_batched").argtypes = [vp, i64, _PT, C.c_int, _PT, C.c_int, _PT, ct, ct, _PP, vp]
    f("compress_batched").argtypes = [vp, i64, C.POINTER(vp), i64, _PT, _PP, vp]
    f("tlr_matmul").argtypes = [vp, i64, i64, i64, _PT, _PT, _PT, C.POINTER(i64), i64, i64, i64, ct, ct, _PP, vp]
    f("potrf").argtypes = [vp, C.c_int, i64, vp, i64, vp]
    f("trsm").argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, i64, i64, ct, vp, i64, vp, i64]
    f("syrk").argtypes = [vp, C.c_int, C.c_int, i64, i64, ct, vp, i64, ct, vp, i64]
    f("fill_triangle").argtypes = [vp, C.c_int, i64, vp, i64, ct]
    f("symmetrize").argtypes = [vp, C.c_int, i64, vp, i64]
    f("transpose").argtypes = [vp, i64, i64, vp, i64, vp, i64]
    f("tlr_trsm_batched").argtypes = [vp, i64, _PT, C.POINTER(vp), C.POINTER(i64)]
    f("tlr_syrk_batched").argtypes = [vp, i64, _PT, C.POINTER(vp), C.POINTER(i64), ct, ct]
    f("tlr_potrf").argtypes = [vp, i64, i64, C.POINTER(vp), i64, _PT, _PP, vp, vp]
    f("tlr_matmul_panel_step").argtypes = [vp, i64, i64, vp, vp, _PT, ct, ct, _PP, vp, C.c_int]
    f("tlr_gemm_workspace").argtypes = [i64, i64, i64, i64, i64]
    f("tlr_gemm_workspace").restype = sz


_declare("d", C.c_double)
_declare("s", C.c_float)


class HcbError(RuntimeError):
    pass


def check(rc: int):
    if rc != 0:
        raise HcbError(f"libhcore_b200 error {rc}: {lib.hcb_last_error().decode()}")


def declared_symbols():
    """Every function name include/hcore_b200.h declares (used by the CPU test that the .so exports them all)."""
    text = open(HEADER).read()
    names = set(re.findall(r"\b(hcb_[a-z_0-9]+)\s*\(", text.split("#define HCB_DECLARE_KERNEL_TABLE")[0]))
    macro = text.split("#define HCB_DECLARE_KERNEL_TABLE(P, T)")[1].split("HCB_DECLARE_KERNEL_TABLE(d, double)")[0]
    for stem in re.findall(r"hcb_##P##([a-z_0-9]+)\s*\(", macro):
        names.add(f"hcb_d{stem}")
        names.add(f"hcb_s{stem}")
    return sorted(names)
