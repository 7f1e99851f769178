"""ctypes binding of libb200spgemm.so (the C ABI in include/b200_spgemm.h).

The shared library is built in-tree by `make -C sparse_linear_algebra_tests_b200/csrc`
(or `__graft_entry__.build()`).  There is no fallback of any kind: if the library is
missing, or no sm_100 device is usable, the first call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200_LIB_PATH") or os.path.join(_HERE, "csrc", "libb200spgemm.so")   # (developer override: A/B builds)

B200_OK, B200_ERR_BADARG, B200_ERR_SHAPE, B200_ERR_ALLOC, B200_ERR_CUDA, B200_ERR_FORMAT, B200_ERR_NCCL = range(7)


class B200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"b200 error {code}: {msg}")
        self.code = code


class ShapeMismatch(B200Error, AssertionError):
    """The reference panics via assert_eq!(self.n, other.n) (src/graph_csr.rs:307,351)."""


class Stats(C.Structure):
    _fields_ = [("rows", C.c_uint64), ("cols", C.c_uint64), ("nnz_a", C.c_uint64), ("nnz_b", C.c_uint64),
                ("nnz_c", C.c_uint64), ("products", C.c_uint64), ("max_row_products", C.c_uint64),
                ("max_row_nnz", C.c_uint64), ("bytes_algorithmic", C.c_uint64), ("ms_symbolic", C.c_float),
                ("ms_numeric", C.c_float), ("ms_total", C.c_float), ("acc_mode", C.c_int32),
                ("kernel_launches", C.c_int32), ("sym_bin_rows", C.c_uint32 * 16), ("pipeline", C.c_uint32),
                ("reserved", C.c_uint32 * 15)]

    def as_dict(self) -> dict:
        d = {k: getattr(self, k) for k, _ in self._fields_ if k not in ("sym_bin_rows", "reserved")}
        d["sym_bin_rows"] = list(self.sym_bin_rows)
        return d


class Config(C.Structure):
    """b200_config (include/b200_spgemm.h): the tuning switches of a context, the MagnusConfig analogue."""
    _fields_ = [("struct_bytes", C.c_uint32)] + [(n, C.c_int32) for n in (
        "pipeline", "placement", "exact_limit_mb", "force_acc_mode", "window_cap_groups", "window_mul", "circular_windows",
        "arc_window", "touched_span", "narrow_scratch", "expand_kernel", "pack_b", "lanes_per_entry_lg", "expand_div", "hash_div",
        "grid_div", "grid_mul", "aux_streams", "fused_threads", "fused_window_cols", "fused_dense_pmax", "heavy_chunk_cols",
        "heavy_kernel", "fused_ring_slots", "fused_product_slots", "rw_cap_percent", "heavy_min_products", "heavy_unit_products", "narrow_download", "commute_swap")] + [("reserved", C.c_int32 * 1)]


EXPORTS = [
    "b200_last_error", "b200_device_count", "b200_ctx_create", "b200_ctx_destroy", "b200_ctx_synchronize",
    "b200_ctx_kernel_launches", "b200_ctx_set_timing", "b200_config_default", "b200_ctx_configure", "b200_ctx_get_config",
    "b200_csr_product_stats", "b200_csr_upload", "b200_csr_upload_idx64",
    "b200_csr_from_device", "b200_csr_free", "b200_csr_info", "b200_csr_device_ptrs", "b200_csr_max_value",
    "b200_csr_download", "b200_csr_download_idx64", "b200_csr_download_async", "b200_spgemm", "b200_row_products",
    "b200_shard_rows_by_products", "b200_csr_row_block", "b200_csr_add", "b200_csr_same_pattern",
    "b200_lattice", "b200_thin", "b200_stdrng_u64", "b200_csr_from_coo", "b200_csr_from_coo_device", "b200_rmat",
    "b200_csr_bandwidth_stats", "b200_csr_permute", "b200_csr_rcm_order",
    "b200_comm_unique_id", "b200_comm_init_rank", "b200_comm_init_all", "b200_comm_destroy", "b200_comm_rank",
    "b200_comm_broadcast_csr", "b200_comm_allgather_csr", "b200_comm_allreduce",
]

_lib = None


def load():
    """Load the engine; raises if the in-tree shared library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build the CUDA engine first "
                          f"(make -C {os.path.dirname(LIB_PATH)}); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, u64, i32 = C.c_void_p, C.c_uint64, C.c_int
    L.b200_last_error.restype = C.c_char_p
    L.b200_device_count.restype = i32
    sigs = {
        "b200_ctx_create": [i32, vp, C.POINTER(vp)],
        "b200_ctx_destroy": [vp],
        "b200_ctx_synchronize": [vp],
        "b200_ctx_kernel_launches": [vp, C.POINTER(u64)],
        "b200_ctx_set_timing": [vp, i32],
        "b200_config_default": [C.POINTER(Config)],
        "b200_ctx_configure": [vp, C.POINTER(Config)],
        "b200_ctx_get_config": [vp, C.POINTER(Config)],
        "b200_csr_product_stats": [vp, vp, C.POINTER(Stats)],
        "b200_csr_upload": [vp, u64, u64, vp, vp, vp, i32, C.POINTER(vp)],
        "b200_csr_upload_idx64": [vp, u64, u64, vp, vp, vp, i32, C.POINTER(vp)],
        "b200_csr_from_device": [vp, u64, u64, u64, vp, vp, vp, i32, C.POINTER(vp)],
        "b200_csr_free": [vp, vp],
        "b200_csr_info": [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64), C.POINTER(i32)],
        "b200_csr_device_ptrs": [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)],
        "b200_csr_max_value": [vp, vp, C.POINTER(u64)],
        "b200_csr_download": [vp, vp, vp, vp, vp],
        "b200_csr_download_idx64": [vp, vp, vp, vp, vp],
        "b200_csr_download_async": [vp, vp, vp, vp, vp],
        "b200_spgemm": [vp, vp, vp, C.POINTER(vp), C.POINTER(Stats)],
        "b200_row_products": [vp, vp, vp, vp],
        "b200_shard_rows_by_products": [vp, vp, vp, i32, vp],
        "b200_csr_row_block": [vp, vp, u64, u64, C.POINTER(vp)],
        "b200_csr_add": [vp, vp, vp, C.POINTER(vp)],
        "b200_csr_same_pattern": [vp, vp, vp, C.POINTER(i32)],
        "b200_lattice": [vp, vp, i32, i32, i32, C.POINTER(vp)],
        "b200_thin": [vp, vp, C.c_double, vp, u64, C.POINTER(vp), C.POINTER(u64)],
        "b200_stdrng_u64": [vp, u64, u64, vp],
        "b200_csr_from_coo": [vp, u64, u64, u64, vp, vp, vp, i32, i32, C.POINTER(vp)],
        "b200_csr_from_coo_device": [vp, u64, u64, u64, vp, vp, vp, i32, i32, C.POINTER(vp)],
        "b200_rmat": [vp, i32, u64, C.c_double, C.c_double, C.c_double, u64, i32, C.POINTER(vp)],
        "b200_csr_bandwidth_stats": [vp, vp, C.POINTER(u64), C.POINTER(C.c_double)],
        "b200_csr_permute": [vp, vp, vp, C.POINTER(vp)],
        "b200_csr_rcm_order": [vp, vp, vp],
        "b200_comm_unique_id": [vp],
        "b200_comm_init_rank": [vp, i32, i32, vp, C.POINTER(vp)],
        "b200_comm_init_all": [vp, i32, vp],
        "b200_comm_destroy": [vp],
        "b200_comm_rank": [vp, C.POINTER(i32), C.POINTER(i32)],
        "b200_comm_broadcast_csr": [vp, vp, i32, C.POINTER(vp)],
        "b200_comm_allgather_csr": [vp, vp, C.POINTER(vp)],
        "b200_comm_allreduce": [vp, vp, i32, i32],
    }
    for name, args in sigs.items():
        f = getattr(L, name)
        f.argtypes = args
        f.restype = i32
    _lib = L
    return L


def check(code: int):
    if code != B200_OK:
        msg = load().b200_last_error().decode("utf-8", "replace")
        raise (ShapeMismatch if code == B200_ERR_SHAPE else B200Error)(code, msg)


def stdrng_u64(seed: bytes, first: int, n: int) -> np.ndarray:
    """StdRng::from_seed(seed).next_u64() outputs first..first+n-1 from the engine's own ChaCha12 (host side, no GPU)."""
    assert len(seed) == 32
    out = np.zeros(n, dtype=np.uint64)
    sb = (C.c_ubyte * 32).from_buffer_copy(seed)
    check(load().b200_stdrng_u64(sb, int(first), int(n), out.ctypes.data))
    return out


def _vdtype(bits: int):
    return np.uint32 if bits == 32 else np.uint64


class Context:
    """One engine context = one device + one stream (b200_ctx)."""

    def __init__(self, device: int = 0, stream: int | None = None):
        L = load()
        h = C.c_void_p()
        check(L.b200_ctx_create(device, C.c_void_p(stream) if stream else None, C.byref(h)))
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            load().b200_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        check(load().b200_ctx_synchronize(self._h))

    def set_timing(self, enabled: bool):
        check(load().b200_ctx_set_timing(self._h, int(enabled)))

    def config(self) -> Config:
        c = Config()
        check(load().b200_ctx_get_config(self._h, C.byref(c)))
        return c

    def configure(self, **fields) -> Config:
        """Change tuning switches (b200_ctx_configure); returns the configuration that was in force before."""
        old, new = self.config(), self.config()
        for k, v in fields.items():
            if not hasattr(new, k) or k in ("struct_bytes", "reserved"):
                raise B200Error(B200_ERR_BADARG, f"unknown configuration field {k!r}")
            setattr(new, k, int(v))
        check(load().b200_ctx_configure(self._h, C.byref(new)))
        return old

    def restore(self, cfg: Config):
        check(load().b200_ctx_configure(self._h, C.byref(cfg)))

    def kernel_launches(self) -> int:
        v = C.c_uint64()
        check(load().b200_ctx_kernel_launches(self._h, C.byref(v)))
        return int(v.value)

    # -- matrices ----------------------------------------------------------------------
    def upload(self, rows: int, cols: int, row_ptr, col_idx, values) -> "DeviceCsr":
        rp = np.ascontiguousarray(row_ptr, dtype=np.uint64)
        vv = np.ascontiguousarray(values)
        if vv.dtype not in (np.uint32, np.uint64):
            raise B200Error(B200_ERR_BADARG, f"values must be uint32 or uint64, got {vv.dtype}")
        bits = 32 if vv.dtype == np.uint32 else 64
        ci = np.ascontiguousarray(col_idx)
        if rp.ndim != 1 or rp.size != rows + 1:
            raise B200Error(B200_ERR_BADARG, f"row_ptr must hold rows + 1 = {rows + 1} entries, got {rp.size}")
        if ci.ndim != 1 or vv.ndim != 1 or ci.size != int(rp[-1]) or vv.size != int(rp[-1]):
            raise B200Error(B200_ERR_BADARG, f"col_idx / values must hold row_ptr[-1] = {int(rp[-1])} entries, got {ci.size} / {vv.size}")
        if ci.dtype.kind not in "ui":
            raise B200Error(B200_ERR_BADARG, f"col_idx must be an integer array, got {ci.dtype}")
        if ci.dtype.kind == "i" and ci.size and int(ci.min()) < 0:
            raise B200Error(B200_ERR_FORMAT, "negative column index")
        if ci.dtype.itemsize == 8:   # 64-bit columns take the path that range-checks them on the device
            ci = np.ascontiguousarray(ci, dtype=np.uint64)
        h = C.c_void_p()
        if ci.dtype == np.uint64:   # MAGNUS usize columns
            check(load().b200_csr_upload_idx64(self._h, rows, cols, rp.ctypes.data, ci.ctypes.data, vv.ctypes.data, bits, C.byref(h)))
        else:
            ci = np.ascontiguousarray(ci, dtype=np.uint32)
            check(load().b200_csr_upload(self._h, rows, cols, rp.ctypes.data, ci.ctypes.data, vv.ctypes.data, bits, C.byref(h)))
        return DeviceCsr(self, h)

    def from_device(self, rows: int, cols: int, nnz: int, d_row_ptr: int, d_col_idx: int, d_values: int, val_bits: int) -> "DeviceCsr":
        h = C.c_void_p()
        check(load().b200_csr_from_device(self._h, rows, cols, nnz, d_row_ptr, d_col_idx, d_values, val_bits, C.byref(h)))
        return DeviceCsr(self, h)

    def spgemm(self, a: "DeviceCsr", b: "DeviceCsr", want_stats: bool = False):
        """C = A x B.  Without `want_stats` the call returns while the kernels are still queued (the product's nnz is
        fetched from the device report the first time it is asked for)."""
        h = C.c_void_p()
        st = Stats() if want_stats else None
        check(load().b200_spgemm(self._h, a._h, b._h, C.byref(h), C.byref(st) if want_stats else None))
        c = DeviceCsr(self, h)
        return (c, st) if want_stats else c

    def add(self, a: "DeviceCsr", b: "DeviceCsr") -> "DeviceCsr":
        h = C.c_void_p()
        check(load().b200_csr_add(self._h, a._h, b._h, C.byref(h)))
        return DeviceCsr(self, h)

    def same_pattern(self, a: "DeviceCsr", b: "DeviceCsr") -> bool:
        s = C.c_int()
        check(load().b200_csr_same_pattern(self._h, a._h, b._h, C.byref(s)))
        return bool(s.value)

    def lattice(self, dims, torus: bool, val_bits: int = 32) -> "DeviceCsr":
        """Moore lattice / torus built on the device (src/graph_csr.rs:177-222)."""
        d = np.asarray(list(dims), dtype=np.uint64)
        h = C.c_void_p()
        check(load().b200_lattice(self._h, d.ctypes.data, int(d.size), int(bool(torus)), val_bits, C.byref(h)))
        return DeviceCsr(self, h)

    def from_coo(self, rows: int, cols: int, r, c, v, val_bits: int = 64, saturating: bool = False) -> "DeviceCsr":
        """COO triplets -> CSR on the device (src/graph_csr.rs:83-129): radix sort by (row, column), duplicate sum, zero drop."""
        r = np.ascontiguousarray(np.asarray(r).ravel())
        c = np.ascontiguousarray(np.asarray(c).ravel())
        if r.size and (int(r.max()) > 0xFFFFFFFF or int(c.max()) > 0xFFFFFFFF or int(r.min()) < 0 or int(c.min()) < 0):
            raise B200Error(B200_ERR_BADARG, "triplet index out of range")
        r32, c32 = r.astype(np.uint32), c.astype(np.uint32)
        vv = np.ascontiguousarray(np.asarray(v).ravel().astype(np.uint32 if val_bits == 32 else np.uint64))
        if not (r32.size == c32.size == vv.size):
            raise B200Error(B200_ERR_BADARG, "triplet arrays differ in length")
        h = C.c_void_p()
        check(load().b200_csr_from_coo(self._h, int(rows), int(cols), int(r32.size), r32.ctypes.data, c32.ctypes.data, vv.ctypes.data,
                                       val_bits, int(bool(saturating)), C.byref(h)))
        return DeviceCsr(self, h)

    def bandwidth_stats(self, a: "DeviceCsr"):
        """(max |r-c|, mean |r-c|) over the stored entries (src/graph_csr.rs:802-818)."""
        mx, avg = C.c_uint64(), C.c_double()
        check(load().b200_csr_bandwidth_stats(self._h, a._h, C.byref(mx), C.byref(avg)))
        return int(mx.value), float(avg.value)

    def permute(self, a: "DeviceCsr", perm) -> "DeviceCsr":
        """Rows and columns reordered by perm[new] = old (src/graph_csr.rs:727-785)."""
        p = np.ascontiguousarray(perm, dtype=np.uint32)
        if p.size != a.rows:
            raise B200Error(B200_ERR_BADARG, f"perm has {p.size} entries, the matrix {a.rows} rows")
        h = C.c_void_p()
        check(load().b200_csr_permute(self._h, a._h, p.ctypes.data, C.byref(h)))
        return DeviceCsr(self, h)

    def rcm_order(self, a: "DeviceCsr") -> np.ndarray:
        """Reverse Cuthill-McKee order of the pattern, perm[new] = old (src/graph_csr.rs:663-723)."""
        p = np.zeros(max(a.rows, 1), dtype=np.uint32)
        check(load().b200_csr_rcm_order(self._h, a._h, p.ctypes.data))
        return p[:a.rows]

    def rmat(self, scale: int, edge_factor: int, a: float, b: float, c: float, seed: int = 42, val_bits: int = 64) -> "DeviceCsr":
        """R-MAT graph generated and assembled on the device (host twin: hostgen.rmat)."""
        h = C.c_void_p()
        check(load().b200_rmat(self._h, int(scale), int(edge_factor), float(a), float(b), float(c), int(seed), val_bits, C.byref(h)))
        return DeviceCsr(self, h)

    def thin(self, a: "DeviceCsr", density: float, seed: bytes = bytes([42] * 32), skip: int = 0):
        """Symmetric thinning on the device (src/graph_csr.rs:225-247, StdRng::from_seed(seed)); returns (matrix, draws taken)."""
        assert len(seed) == 32
        sb = (C.c_ubyte * 32).from_buffer_copy(seed)
        h, n = C.c_void_p(), C.c_uint64()
        check(load().b200_thin(self._h, a._h, float(density), sb, int(skip), C.byref(h), C.byref(n)))
        return DeviceCsr(self, h), int(n.value)

    def row_products(self, a: "DeviceCsr", b: "DeviceCsr") -> np.ndarray:
        out = np.zeros(a.rows, dtype=np.uint64)
        check(load().b200_row_products(self._h, a._h, b._h, out.ctypes.data))
        return out

    def shard_rows_by_products(self, a: "DeviceCsr", b: "DeviceCsr", nparts: int) -> np.ndarray:
        cuts = np.zeros(nparts + 1, dtype=np.uint64)
        check(load().b200_shard_rows_by_products(self._h, a._h, b._h, nparts, cuts.ctypes.data))
        return cuts

    def row_block(self, a: "DeviceCsr", r0: int, r1: int) -> "DeviceCsr":
        h = C.c_void_p()
        check(load().b200_csr_row_block(self._h, a._h, r0, r1, C.byref(h)))
        return DeviceCsr(self, h)


class Comm:
    """One rank of a multi-GPU job (b200_comm): an NCCL communicator bound to an engine context."""

    def __init__(self, ctx: Context, nranks: int, rank: int, unique_id: bytes):
        assert len(unique_id) == 128
        buf = (C.c_ubyte * 128).from_buffer_copy(unique_id)
        h = C.c_void_p()
        check(load().b200_comm_init_rank(ctx._h, int(nranks), int(rank), buf, C.byref(h)))
        self.ctx, self._h, self.rank, self.size = ctx, h, int(rank), int(nranks)

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_ubyte * 128)()
        check(load().b200_comm_unique_id(buf))
        return bytes(buf)

    def broadcast(self, src: "DeviceCsr | None", root: int = 0) -> "DeviceCsr":
        """Replicate `src` (given on `root`) on every rank (four ncclBroadcast)."""
        h = C.c_void_p()
        check(load().b200_comm_broadcast_csr(self._h, src._h if src is not None else None, int(root), C.byref(h)))
        return DeviceCsr(self.ctx, h)

    def allgather(self, block: "DeviceCsr") -> "DeviceCsr":
        """Row blocks in rank order -> the whole matrix on every rank (device buffers, grouped ncclBroadcast)."""
        h = C.c_void_p()
        check(load().b200_comm_allgather_csr(self._h, block._h, C.byref(h)))
        return DeviceCsr(self.ctx, h)

    def allreduce(self, values, op: str):
        """op in {"sum_u64", "max_u64", "sum_f64", "max_f64"} over at most 32 scalars; returns the reduced numpy array."""
        code = {"sum_u64": 0, "max_u64": 1, "sum_f64": 2, "max_f64": 3}[op]
        a = np.ascontiguousarray(values, dtype=np.uint64 if code < 2 else np.float64).copy()
        check(load().b200_comm_allreduce(self._h, a.ctypes.data, int(a.size), code))
        return a

    def close(self):
        if getattr(self, "_h", None):
            load().b200_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceCsr:
    """Owning wrapper of a b200_csr handle."""

    def __init__(self, ctx: Context, handle):
        self.ctx = ctx
        self._h = handle
        r, c, b = C.c_uint64(), C.c_uint64(), C.c_int()
        check(load().b200_csr_info(self._h, C.byref(r), C.byref(c), None, C.byref(b)))
        self.rows, self.cols, self.val_bits = int(r.value), int(c.value), int(b.value)
        self._nnz = None

    @property
    def nnz(self) -> int:
        """Stored entries; for a product still in flight this waits for its device report."""
        if self._nnz is None:
            n = C.c_uint64()
            check(load().b200_csr_info(self._h, None, None, C.byref(n), None))
            self._nnz = int(n.value)
        return self._nnz

    def product_stats(self) -> Stats:
        st = Stats()
        check(load().b200_csr_product_stats(self.ctx._h, self._h, C.byref(st)))
        return st

    def free(self):
        if getattr(self, "_h", None) and getattr(self.ctx, "_h", None):
            load().b200_csr_free(self.ctx._h, self._h)
        self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def download(self, idx64: bool = False):
        rp = np.empty(self.rows + 1, dtype=np.uint64)
        ci = np.empty(self.nnz, dtype=np.uint64 if idx64 else np.uint32)
        vv = np.empty(self.nnz, dtype=_vdtype(self.val_bits))
        f = load().b200_csr_download_idx64 if idx64 else load().b200_csr_download
        check(f(self.ctx._h, self._h, rp.ctypes.data, ci.ctypes.data, vv.ctypes.data))
        return rp, ci, vv

    def download_async_into(self, rp: int, ci: int, vv: int):
        """Raw-pointer variant for pinned buffers (bench end-to-end path)."""
        check(load().b200_csr_download_async(self.ctx._h, self._h, rp, ci, vv))

    def device_ptrs(self):
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(load().b200_csr_device_ptrs(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def max_value(self) -> int:
        v = C.c_uint64()
        check(load().b200_csr_max_value(self.ctx._h, self._h, C.byref(v)))
        return int(v.value)
