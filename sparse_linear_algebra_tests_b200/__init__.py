"""B200-native SpGEMM engine behind the reference's graph-matrix interface.

Only the hot path of imlvts/sparse-linear-algebra-tests lives here: CSR x CSR SpGEMM over
saturating unsigned path-count matrices (`B200Matrix.matmul`) with its host-side builders.
The CUDA kernels and the C ABI are in `csrc/` (libb200spgemm.so); importing this package
does not need a GPU, using it does.
"""
from . import hostgen
from ._native import B200Error, Comm, Context, DeviceCsr, ShapeMismatch, Stats, EXPORTS, LIB_PATH
from .graph_b200 import B200Matrix, default_context, set_default_context
from .einsum import InvalidSpec, einsum_sparse_driven, einsum_sparse_hash

__all__ = ["B200Matrix", "Context", "Comm", "DeviceCsr", "Stats", "B200Error", "ShapeMismatch", "hostgen",
           "default_context", "set_default_context", "EXPORTS", "LIB_PATH", "einsum_sparse_driven", "einsum_sparse_hash", "InvalidSpec"]
