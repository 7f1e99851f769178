"""ctypes front-end of the CPU oracle (oracle/spgemm_oracle.c).

TEST INFRASTRUCTURE ONLY -- may be imported from tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs, never from the product
package.  See the header of spgemm_oracle.c for what pins its parity.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "spgemm_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return _LIB_PATH


class _OCsr(C.Structure):
    _fields_ = [("rows", C.c_uint64), ("cols", C.c_uint64), ("nnz", C.c_uint64),
                ("row_ptr", C.POINTER(C.c_uint64)), ("col_idx", C.POINTER(C.c_uint32)),
                ("values", C.c_void_p), ("val_bits", C.c_int)]


_P = C.POINTER(_OCsr)
_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.oracle_free.argtypes = [_P]
        L.oracle_wrap.restype = _P
        L.oracle_wrap.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_unwrap.argtypes = [_P]
        for suf in ("u32", "u64"):
            getattr(L, f"oracle_matmul_seq_{suf}").restype = _P
            getattr(L, f"oracle_matmul_seq_{suf}").argtypes = [_P, _P]
            getattr(L, f"oracle_matmul_par_{suf}").restype = _P
            getattr(L, f"oracle_matmul_par_{suf}").argtypes = [_P, _P, C.c_int]
            getattr(L, f"oracle_add_{suf}").restype = _P
            getattr(L, f"oracle_add_{suf}").argtypes = [_P, _P]
            getattr(L, f"oracle_from_coo_{suf}").restype = _P
            getattr(L, f"oracle_from_coo_{suf}").argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p,
                                                              C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_get_u32.restype = C.c_uint32
        L.oracle_get_u32.argtypes = [_P, C.c_uint64, C.c_uint32]
        L.oracle_get_u64.restype = C.c_uint64
        L.oracle_get_u64.argtypes = [_P, C.c_uint64, C.c_uint32]
        L.oracle_widen_to_u64.restype = _P
        L.oracle_widen_to_u64.argtypes = [_P]
        L.oracle_narrow_to_u32.restype = _P
        L.oracle_narrow_to_u32.argtypes = [_P]
        L.oracle_row_products.argtypes = [_P, _P, C.c_void_p]
        L.oracle_lattice.restype = _P
        L.oracle_lattice.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.oracle_thin_stdrng.restype = _P
        L.oracle_thin_stdrng.argtypes = [_P, C.c_void_p, C.c_double]
        L.oracle_chacha12_u64.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        L.oracle_lattice_csr_xorshift.restype = _P
        L.oracle_lattice_csr_xorshift.argtypes = [C.c_uint64, C.c_double, C.c_uint64, C.c_int]
        L.oracle_rmat.restype = _P
        L.oracle_rmat.argtypes = [C.c_int, C.c_uint64, C.c_double, C.c_double, C.c_double, C.c_uint64, C.c_int]
        L.oracle_time_matmul.restype = C.c_double
        L.oracle_time_matmul.argtypes = [_P, _P, C.c_int, C.c_int, C.c_int]
        L.oracle_rcm_order.argtypes = [_P, C.c_void_p]
        L.oracle_rcm_order.restype = None
        L.oracle_permute.argtypes = [_P, C.c_void_p]
        L.oracle_permute.restype = _P
        L.oracle_bandwidth_stats.argtypes = [_P, C.c_void_p, C.c_void_p]
        L.oracle_bandwidth_stats.restype = None
        L.oracle_max_threads.restype = C.c_int
        for f, t in (("oracle_sadd_u32", C.c_uint32), ("oracle_smul_u32", C.c_uint32),
                     ("oracle_sadd_u64", C.c_uint64), ("oracle_smul_u64", C.c_uint64)):
            getattr(L, f).restype = t
            getattr(L, f).argtypes = [t, t]
        _lib = L
    return _lib


def _vdtype(bits: int):
    return np.uint32 if bits == 32 else np.uint64


@dataclass
class Csr:
    """Host CSR with the reference's field names (src/graph_csr.rs:42-53)."""
    rows: int
    cols: int
    row_ptr: np.ndarray   # uint64, rows+1
    col_idx: np.ndarray   # uint32
    values: np.ndarray    # uint32 or uint64

    @property
    def n(self) -> int:
        return self.rows

    @property
    def val_bits(self) -> int:
        return 32 if self.values.dtype == np.uint32 else 64

    def nnz(self) -> int:
        return int(self.values.shape[0])

    def get(self, r: int, c: int) -> int:
        s, e = int(self.row_ptr[r]), int(self.row_ptr[r + 1])
        i = int(np.searchsorted(self.col_idx[s:e], c))
        return int(self.values[s + i]) if i < e - s and self.col_idx[s + i] == c else 0

    def row_block(self, r0: int, r1: int) -> "Csr":
        s, e = int(self.row_ptr[r0]), int(self.row_ptr[r1])
        return Csr(r1 - r0, self.cols, (self.row_ptr[r0:r1 + 1] - self.row_ptr[r0]).astype(np.uint64),
                   self.col_idx[s:e].copy(), self.values[s:e].copy())

    def equals(self, o: "Csr") -> bool:
        return (self.rows == o.rows and self.cols == o.cols and self.values.dtype == o.values.dtype
                and np.array_equal(self.row_ptr, o.row_ptr) and np.array_equal(self.col_idx, o.col_idx)
                and np.array_equal(self.values, o.values))

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.csr_matrix((self.values.astype(np.uint64), self.col_idx.astype(np.int64),
                              self.row_ptr.astype(np.int64)), shape=(self.rows, self.cols))


class _Wrapped:
    def __init__(self, m: Csr):
        self.keep = (np.ascontiguousarray(m.row_ptr, dtype=np.uint64),
                     np.ascontiguousarray(m.col_idx, dtype=np.uint32),
                     np.ascontiguousarray(m.values))
        self.p = lib().oracle_wrap(m.rows, m.cols, m.nnz(), self.keep[0].ctypes.data, self.keep[1].ctypes.data,
                                   self.keep[2].ctypes.data, m.val_bits)

    def __enter__(self):
        return self.p

    def __exit__(self, *a):
        lib().oracle_unwrap(self.p)


def _take(p) -> Csr:
    if not p:
        raise ValueError("oracle: shape or value-width mismatch")
    m = p.contents
    rows, cols, nnz, bits = int(m.rows), int(m.cols), int(m.nnz), int(m.val_bits)
    rp = np.ctypeslib.as_array(m.row_ptr, shape=(rows + 1,)).copy()
    ci = np.ctypeslib.as_array(m.col_idx, shape=(max(nnz, 1),))[:nnz].copy()
    vp = C.cast(m.values, C.POINTER(C.c_uint32 if bits == 32 else C.c_uint64))
    vv = np.ctypeslib.as_array(vp, shape=(max(nnz, 1),))[:nnz].copy()
    lib().oracle_free(p)
    return Csr(rows, cols, rp, ci, vv)


def matmul(a: Csr, b: Csr) -> Csr:
    """CsrMatrix::matmul (sequential Gustavson)."""
    with _Wrapped(a) as pa, _Wrapped(b) as pb:
        return _take(getattr(lib(), f"oracle_matmul_seq_u{a.val_bits}")(pa, pb))


def matmul_par(a: Csr, b: Csr, nthreads: int | None = None) -> Csr:
    """CsrMatrix::matmul_par (two-pass symbolic + numeric, OpenMP for rayon)."""
    nt = nthreads or max_threads()
    with _Wrapped(a) as pa, _Wrapped(b) as pb:
        return _take(getattr(lib(), f"oracle_matmul_par_u{a.val_bits}")(pa, pb, nt))


def add(a: Csr, b: Csr) -> Csr:
    with _Wrapped(a) as pa, _Wrapped(b) as pb:
        return _take(getattr(lib(), f"oracle_add_u{a.val_bits}")(pa, pb))


def from_coo(rows: int, cols: int, r, c, v, val_bits: int = 32, saturating: bool = False) -> Csr:
    r = np.ascontiguousarray(r, dtype=np.uint32)
    c = np.ascontiguousarray(c, dtype=np.uint32)
    v = np.ascontiguousarray(v, dtype=_vdtype(val_bits))
    return _take(getattr(lib(), f"oracle_from_coo_u{val_bits}")(rows, cols, r.shape[0], r.ctypes.data, c.ctypes.data,
                                                                 v.ctypes.data, int(saturating)))


def from_edges(n: int, edges, val_bits: int = 32) -> Csr:
    e = np.asarray(edges, dtype=np.uint32).reshape(-1, 2)
    return from_coo(n, n, e[:, 0], e[:, 1], np.ones(e.shape[0], dtype=_vdtype(val_bits)), val_bits)


def from_edges_undirected(n: int, edges, val_bits: int = 32) -> Csr:
    """graph_csr.rs:137-147."""
    r, c = [], []
    for a, b in edges:
        r.append(a); c.append(b)
        if a != b:
            r.append(b); c.append(a)
    return from_coo(n, n, r, c, np.ones(len(r), dtype=_vdtype(val_bits)), val_bits)


def identity(n: int, val_bits: int = 32) -> Csr:
    return Csr(n, n, np.arange(n + 1, dtype=np.uint64), np.arange(n, dtype=np.uint32),
               np.ones(n, dtype=_vdtype(val_bits)))


def empty(n: int, val_bits: int = 32) -> Csr:
    return Csr(n, n, np.zeros(n + 1, dtype=np.uint64), np.zeros(0, dtype=np.uint32), np.zeros(0, dtype=_vdtype(val_bits)))


def widen(a: Csr) -> Csr:
    return Csr(a.rows, a.cols, a.row_ptr.copy(), a.col_idx.copy(), a.values.astype(np.uint64))


def narrow(a: Csr) -> Csr:
    return Csr(a.rows, a.cols, a.row_ptr.copy(), a.col_idx.copy(), a.values.astype(np.uint32))


def row_products(a: Csr, b: Csr) -> np.ndarray:
    out = np.zeros(a.rows, dtype=np.uint64)
    with _Wrapped(a) as pa, _Wrapped(b) as pb:
        lib().oracle_row_products(pa, pb, out.ctypes.data)
    return out


def lattice(dims, torus: bool, val_bits: int = 32) -> Csr:
    d = np.asarray(dims, dtype=np.uint64)
    return _take(lib().oracle_lattice(d.ctypes.data, d.shape[0], int(torus), val_bits))


def thin_stdrng(a: Csr, seed_bytes: bytes, density: float) -> Csr:
    assert len(seed_bytes) == 32
    sb = np.frombuffer(seed_bytes, dtype=np.uint8).copy()
    with _Wrapped(a) as pa:
        return _take(lib().oracle_thin_stdrng(pa, sb.ctypes.data, float(density)))


def chacha12_u64(seed_bytes: bytes, n: int) -> np.ndarray:
    sb = np.frombuffer(seed_bytes, dtype=np.uint8).copy()
    out = np.zeros(n, dtype=np.uint64)
    lib().oracle_chacha12_u64(sb.ctypes.data, n, out.ctypes.data)
    return out


def lattice_csr_xorshift(s: int, target_epn: float, seed: int, val_bits: int = 32) -> Csr:
    return _take(lib().oracle_lattice_csr_xorshift(s, float(target_epn), seed, val_bits))


def rmat(scale: int, edge_factor: int, a: float, b: float, c: float, seed: int, val_bits: int = 64) -> Csr:
    return _take(lib().oracle_rmat(scale, edge_factor, a, b, c, seed, val_bits))


def reference_bench_instance(side: int = 30, target_epn: float = 3.0, val_bits: int = 32) -> Csr:
    """The exact operand of bench_repeated_exponentiation (graph_magnus.rs:707-719)."""
    full = lattice([side, side, side], True, val_bits)
    density = target_epn / (full.nnz() / full.rows)
    return thin_stdrng(full, bytes([42] * 32), density)


def rcm_order(a: Csr) -> np.ndarray:
    """CsrMatrix::rcm's ordering (graph_csr.rs:663-723): perm[new] = old."""
    perm = np.zeros(a.rows, dtype=np.uint32)
    with _Wrapped(a) as pa:
        lib().oracle_rcm_order(pa, perm.ctypes.data)
    return perm


def permute(a: Csr, perm) -> Csr:
    """CsrMatrix::permute (graph_csr.rs:727-785)."""
    perm = np.ascontiguousarray(perm, dtype=np.uint32)
    with _Wrapped(a) as pa:
        return _take(lib().oracle_permute(pa, perm.ctypes.data))


def bandwidth_stats(a: Csr):
    """CsrMatrix::bandwidth_stats (graph_csr.rs:802-818): (max |r-c|, mean |r-c|)."""
    mx, avg = C.c_uint64(), C.c_double()
    with _Wrapped(a) as pa:
        lib().oracle_bandwidth_stats(pa, C.byref(mx), C.byref(avg))
    return int(mx.value), float(avg.value)


def time_matmul(a: Csr, b: Csr, par: bool, nthreads: int, iters: int) -> float:
    """Seconds per multiply, reference protocol (wall clock, alloc+free inside)."""
    with _Wrapped(a) as pa, _Wrapped(b) as pb:
        return float(lib().oracle_time_matmul(pa, pb, int(par), nthreads, iters))


def max_threads() -> int:
    return int(lib().oracle_max_threads())
