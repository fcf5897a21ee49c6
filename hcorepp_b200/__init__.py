"""hcorepp_b200 -- B200-native (sm_100a) tile low-rank GEMM behind the HCore++ interface.

Only what the TLR-GEMM hot path needs: csrc/ (hand-written CUDA kernels + the C ABI of include/hcore_b200.h), and
api.py, a thin host-side mirror of the reference's operator interface (RunContext, CompressionParameters, DenseTile,
CompressedTile, HCore.Gemm, TileMatrix, tile_matrix_multiplication) used by the parity tests and bench.py.
Importing this package requires the built CUDA library (no CPU fallback).
"""
from . import _capi  # noqa: F401  (raises ImportError when libhcore_b200.so is missing)
from .api import (RunContext, CompressionParameters, DenseTile, CompressedTile, HCore, TileMatrix,  # noqa: F401
                  tile_matrix_multiplication, gemm_batched, SymTileMatrix, tlr_cholesky)

__all__ = ["RunContext", "CompressionParameters", "DenseTile", "CompressedTile", "HCore", "TileMatrix",
           "tile_matrix_multiplication", "gemm_batched", "SymTileMatrix", "tlr_cholesky"]
