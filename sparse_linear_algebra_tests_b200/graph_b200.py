"""`B200Matrix` -- host-side mirror of the reference's graph matrix types, backed by the CUDA engine.

Same inherent-method surface as `MagnusMatrix` (src/graph_magnus.rs:16-448) and `CsrMatrix`
(src/graph_csr.rs:55-657): new, identity, from_edges, from_edges_undirected, from_adjacency,
random, lattice, thin, get, nnz, matmul (+ matmul_par / matmul_seq aliases), add,
reachability_sum, power_until_stable, connected_components[_uf], num_components, print, and the
`NDIndex` / `Sparse2D` views (ndim, dim, get_opt, n_rows, row_nnz, row_entry;
einsum-dyn/src/lib.rs:126-154, einsum-dyn/src/sparse.rs:42-55).  The matrix lives on the GPU
(a `b200_csr` handle); `matmul`/`add` never leave the device.  Builders run on the host exactly
like the reference's and upload once.  Shape mismatch raises `ShapeMismatch` (an AssertionError),
matching the reference's `assert_eq!` panic.

There is no CPU fallback: every arithmetic method goes through libb200spgemm.so.
"""
from __future__ import annotations

import sys
from typing import Iterable

import numpy as np

from . import hostgen
from ._native import Context, DeviceCsr, ShapeMismatch, Stats

_default_ctx: Context | None = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


def set_default_context(ctx: Context | None):
    global _default_ctx
    _default_ctx = ctx


class B200Matrix:
    """n x n (or row-block m x n) saturating unsigned sparse matrix resident on one B200."""

    def __init__(self, dev: DeviceCsr, host: hostgen.HostCsr | None = None):
        self._dev = dev
        self._host = host          # lazily downloaded copy for get()/print()/NDIndex
        self.last_stats: Stats | None = None
        self.perm: np.ndarray | None = None   # perm[new] = old, set by rcm() / permute() (src/graph_csr.rs:51)

    # ------------------------------------------------------------------ plumbing
    @property
    def n(self) -> int:
        return self._dev.rows

    @property
    def shape(self):
        return (self._dev.rows, self._dev.cols)

    @property
    def val_bits(self) -> int:
        return self._dev.val_bits

    @property
    def device(self) -> DeviceCsr:
        return self._dev

    @classmethod
    def from_host(cls, h: hostgen.HostCsr, ctx: Context | None = None) -> "B200Matrix":
        ctx = ctx or default_context()
        return cls(ctx.upload(h.rows, h.cols, h.row_ptr, h.col_idx, h.values), h)

    @classmethod
    def from_parts(cls, rows: int, cols: int, row_ptr, col_idx, values, ctx: Context | None = None) -> "B200Matrix":
        """The three CSR arrays (SparseMatrixCSR::new at src/graph_magnus.rs:74; u64 `col_idx` accepted)."""
        ctx = ctx or default_context()
        return cls(ctx.upload(rows, cols, row_ptr, col_idx, values))

    def to_host(self) -> hostgen.HostCsr:
        if self._host is None:
            rp, ci, vv = self._dev.download()
            self._host = hostgen.HostCsr(self._dev.rows, self._dev.cols, rp, ci, vv)
        return self._host

    @property
    def row_ptr(self) -> np.ndarray:
        return self.to_host().row_ptr

    @property
    def col_idx(self) -> np.ndarray:
        return self.to_host().col_idx

    @property
    def values(self) -> np.ndarray:
        return self.to_host().values

    # ------------------------------------------------------------------ builders (host, like the reference)
    @classmethod
    def new(cls, n: int, val_bits: int = 64, ctx=None) -> "B200Matrix":
        return cls.from_host(hostgen.empty(n, val_bits), ctx)

    @classmethod
    def identity(cls, n: int, val_bits: int = 64, ctx=None) -> "B200Matrix":
        return cls.from_host(hostgen.identity(n, val_bits), ctx)

    @classmethod
    def from_coo(cls, n: int, triplets: Iterable, val_bits: int = 64, ctx=None, device: bool = True) -> "B200Matrix":
        """src/graph_csr.rs:83-129 / src/graph_magnus.rs:34-76: sort by (row, column), sum duplicates, drop zeros -- on the
        device (b200_csr_from_coo: radix sort + run sums); `device=False` assembles on the host and uploads."""
        t = np.asarray(list(triplets), dtype=np.uint64).reshape(-1, 3)
        if device:
            return cls((ctx or default_context()).from_coo(n, n, t[:, 0], t[:, 1], t[:, 2], val_bits))
        return cls.from_host(hostgen.from_coo(n, n, t[:, 0], t[:, 1], t[:, 2], val_bits), ctx)

    @classmethod
    def rmat(cls, scale: int, edge_factor: int = 16, abc=(0.45, 0.15, 0.15), seed: int = 42, val_bits: int = 64, ctx=None) -> "B200Matrix":
        """R-MAT input of BASELINE configs[3], generated on the device (b200_rmat)."""
        return cls((ctx or default_context()).rmat(scale, edge_factor, abc[0], abc[1], abc[2], seed, val_bits))

    @classmethod
    def from_edges(cls, n: int, edges: Iterable, val_bits: int = 64, ctx=None) -> "B200Matrix":
        e = np.asarray(list(edges), dtype=np.int64).reshape(-1, 2)
        return cls.from_host(hostgen.from_coo(n, n, e[:, 0], e[:, 1], np.ones(e.shape[0], hostgen.vdtype(val_bits)), val_bits), ctx)

    @classmethod
    def from_edges_undirected(cls, n: int, edges: Iterable, val_bits: int = 64, ctx=None) -> "B200Matrix":
        r, c = [], []
        for a, b in edges:
            r.append(a); c.append(b)
            if a != b:
                r.append(b); c.append(a)
        return cls.from_host(hostgen.from_coo(n, n, r, c, np.ones(len(r), hostgen.vdtype(val_bits)), val_bits), ctx)

    @classmethod
    def from_adjacency(cls, pairs: Iterable, val_bits: int = 64, ctx=None):
        """Named edge pairs -> (matrix, name -> index) (src/graph_magnus.rs:94-116); ids in first-seen order."""
        names: dict[str, int] = {}
        edges = []
        for a, b in pairs:
            ai = names.setdefault(a, len(names))
            bi = names.setdefault(b, len(names))
            edges.append((ai, bi))
        return cls.from_edges(len(names), edges, val_bits, ctx), dict(sorted(names.items()))

    @classmethod
    def random(cls, rng: np.random.Generator, n: int, m: int, val_bits: int = 64, ctx=None) -> "B200Matrix":
        """m random directed edges without self-loops (src/graph_magnus.rs:119-130; numpy Generator as the rng)."""
        assert n >= 2, "need at least 2 nodes to avoid self-loops"
        r = rng.integers(0, n, size=m)
        c = rng.integers(0, n - 1, size=m)
        c = np.where(c >= r, c + 1, c)
        return cls.from_host(hostgen.from_coo(n, n, r, c, np.ones(m, hostgen.vdtype(val_bits)), val_bits), ctx)

    @classmethod
    def lattice(cls, dims, torus: bool, val_bits: int = 64, ctx=None, device: bool = True) -> "B200Matrix":
        """N-d Moore lattice / torus (src/graph_csr.rs:177-222).  Built by the engine's device generator (up to 4
        dimensions); `device=False`, or more dimensions, builds it with the host generator and uploads."""
        if device and len(list(dims)) <= 4:
            return cls((ctx or default_context()).lattice(dims, torus, val_bits))
        return cls.from_host(hostgen.lattice(dims, torus, val_bits), ctx)

    def thin(self, density: float, seed: bytes = bytes([42] * 32), skip: int = 0, device: bool = True) -> "B200Matrix":
        """Symmetric Bernoulli thinning driven by StdRng::from_seed(seed) (src/graph_csr.rs:225-247); `skip` = draws
        already taken from the same generator.  `last_draws` on the result = draws this call consumed."""
        if device:
            dev, n = self._dev.ctx.thin(self._dev, density, seed, skip)
            out = B200Matrix(dev)
        else:
            h = self.to_host()
            out, n = B200Matrix.from_host(hostgen.thin(h, density, seed, skip), self._dev.ctx), hostgen.draws_of_thin(h)
        out.last_draws = n
        return out

    # ------------------------------------------------------------------ queries
    def get(self, r: int, c: int) -> int:
        return self.to_host().get(r, c)

    def nnz(self) -> int:
        return self._dev.nnz

    # NDIndex<u64> (src/graph_magnus.rs:434-448)
    def ndim(self) -> int:
        return 2

    def dim(self, _axis: int) -> int:
        return self.n

    def get_opt(self, ix):
        h = self.to_host()
        r, c = ix
        s, e = int(h.row_ptr[r]), int(h.row_ptr[r + 1])
        i = int(np.searchsorted(h.col_idx[s:e], c))
        return int(h.values[s + i]) if i < e - s and int(h.col_idx[s + i]) == c else None

    def set(self, _ix, _v):
        raise TypeError("B200Matrix is immutable after construction")

    # Sparse2D<u64> (einsum-dyn/src/sparse.rs:42-55; CsrMatrix's impl at src/graph_csr.rs:861-871)
    def n_rows(self) -> int:
        return self._dev.rows

    def row_nnz(self, row: int) -> int:
        h = self.to_host()
        return int(h.row_ptr[row + 1]) - int(h.row_ptr[row])

    def row_entry(self, row: int, idx: int):
        h = self.to_host()
        p = int(h.row_ptr[row]) + idx
        return int(h.col_idx[p]), int(h.values[p])

    # ------------------------------------------------------------------ arithmetic (device)
    def matmul(self, other: "B200Matrix", want_stats: bool = False) -> "B200Matrix":
        """C = self x other on the GPU (replaces CsrMatrix::matmul/_par and MagnusMatrix::matmul/_seq)."""
        if self._dev.cols != other._dev.rows or self.val_bits != other.val_bits:
            raise ShapeMismatch(2, f"matmul: {self.shape} u{self.val_bits} x {other.shape} u{other.val_bits}")
        ctx = self._dev.ctx
        if want_stats:
            dev, st = ctx.spgemm(self._dev, other._dev, True)
            out = B200Matrix(dev)
            out.last_stats = st
            return out
        return B200Matrix(ctx.spgemm(self._dev, other._dev))

    matmul_par = matmul
    matmul_seq = matmul

    def add(self, other: "B200Matrix") -> "B200Matrix":
        if self.shape != other.shape or self.val_bits != other.val_bits:
            raise ShapeMismatch(2, f"add: {self.shape} vs {other.shape}")
        return B200Matrix(self._dev.ctx.add(self._dev, other._dev))

    def same_pattern(self, other: "B200Matrix") -> bool:
        return self._dev.ctx.same_pattern(self._dev, other._dev)

    # ------------------------------------------------------------------ locality pre-pass (src/graph_csr.rs:663-818), in place like the reference
    def permute(self, perm) -> None:
        """Reorder rows and columns by perm[new] = old and remember it (src/graph_csr.rs:727-785); applied on the device."""
        perm = np.ascontiguousarray(perm, dtype=np.uint32)
        self._dev, self._host = self._dev.ctx.permute(self._dev, perm), None
        self.perm = perm.copy()

    def rcm(self) -> None:
        """Reverse Cuthill-McKee reordering (src/graph_csr.rs:663-723); the permutation stays in `perm`."""
        self.permute(self._dev.ctx.rcm_order(self._dev))

    def unpermute(self) -> None:
        """Undo the stored permutation (src/graph_csr.rs:787-799); no-op without one."""
        if self.perm is None:
            return
        inv = np.empty_like(self.perm)
        inv[self.perm] = np.arange(self.perm.size, dtype=np.uint32)
        self.permute(inv)
        self.perm = None

    def bandwidth_stats(self):
        """(max |r-c|, mean |r-c|) over the stored entries (src/graph_csr.rs:802-818)."""
        return self._dev.ctx.bandwidth_stats(self._dev)

    # ------------------------------------------------------------------ drivers built on matmul/add
    def reachability_sum(self):
        """A + A^2 + ... until nnz of the sum stops growing (src/graph_csr.rs:545-558)."""
        power, total, k = self, self, 1
        while True:
            power = power.matmul(self)
            k += 1
            new_total = total.add(power)
            if new_total.nnz() == total.nnz():
                return new_total, k
            total = new_total

    def power_until_stable(self):
        """Repeated squaring until the sparsity pattern is stable (src/graph_csr.rs:561-575)."""
        current, k = self, 0
        while True:
            nxt = current.matmul(current)
            k += 1
            if nxt.nnz() == current.nnz() and nxt.same_pattern(current):
                return nxt, k
            current = nxt

    def connected_components(self):
        """Via the transitive closure of A + I (src/graph_csr.rs:578-600)."""
        with_id = self.add(B200Matrix.identity(self.n, self.val_bits, self._dev.ctx))
        closure, _ = with_id.power_until_stable()
        h = closure.to_host()
        n = self.n
        comp = [-1] * n
        nxt = 0
        for i in range(n):
            if comp[i] != -1:
                continue
            comp[i] = nxt
            row_i = set(h.col_idx[int(h.row_ptr[i]):int(h.row_ptr[i + 1])].tolist())
            for j in row_i:
                if j > i and comp[j] == -1 and h.get(j, i) > 0:
                    comp[j] = nxt
            nxt += 1
        return comp

    def connected_components_uf(self):
        """Union-find over stored entries, canonical ids in first-seen order (src/graph_csr.rs:603-650)."""
        h = self.to_host()
        n = self.n
        parent = list(range(n))

        def find(x):
            while parent[x] != x:
                parent[x] = parent[parent[x]]
                x = parent[x]
            return x

        rows = h.row_of_entry()
        for r, c in zip(rows.tolist(), h.col_idx.tolist()):
            ra, rb = find(r), find(c)
            if ra != rb:
                parent[rb] = ra
        ids, out = {}, []
        for i in range(n):
            out.append(ids.setdefault(find(i), len(ids)))
        return out

    def num_components(self) -> int:
        comp = self.connected_components_uf()
        return max(comp) + 1 if comp else 0

    def print(self, file=sys.stdout):
        for r in range(self.n):
            print(" ".join("." if self.get(r, c) == 0 else str(self.get(r, c)) for c in range(self._dev.cols)), file=file)
