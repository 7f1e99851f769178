"""Host-side constructors feeding the SpGEMM path (numpy; not the hot path).

Mirrors the reference's CPU builders that every graph matrix type repeats:
`from_coo` (src/graph_csr.rs:83-129), `lattice` (:177-222), `thin` (:225-247), the bench
instance of `bench_repeated_exponentiation` (src/graph_magnus.rs:707-719) including rand
0.9.2's StdRng (ChaCha12) so the *exact* reference operand is reproduced, the portable
xorshift torus of linalg/benches/perf.rs:43-95, and an R-MAT generator.
Independent of oracle/ (tests cross-check the two).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

U64 = np.uint64
U32 = np.uint32


def vdtype(val_bits: int):
    if val_bits not in (32, 64):
        raise ValueError("val_bits must be 32 or 64")
    return U32 if val_bits == 32 else U64


@dataclass
class HostCsr:
    """Host CSR with the reference's field names (src/graph_csr.rs:42-53)."""
    rows: int
    cols: int
    row_ptr: np.ndarray   # uint64[rows+1]
    col_idx: np.ndarray   # uint32[nnz], ascending and unique per row
    values: np.ndarray    # uint32[nnz] or uint64[nnz]

    @property
    def n(self) -> int:
        return self.rows

    @property
    def val_bits(self) -> int:
        return 32 if self.values.dtype == U32 else 64

    def nnz(self) -> int:
        return int(self.values.shape[0])

    def get(self, r: int, c: int) -> int:
        s, e = int(self.row_ptr[r]), int(self.row_ptr[r + 1])
        i = int(np.searchsorted(self.col_idx[s:e], c))
        return int(self.values[s + i]) if i < e - s and int(self.col_idx[s + i]) == c else 0

    def row_of_entry(self) -> np.ndarray:
        return np.repeat(np.arange(self.rows, dtype=np.int64), np.diff(self.row_ptr.astype(np.int64)))

    def row_block(self, r0: int, r1: int) -> "HostCsr":
        s, e = int(self.row_ptr[r0]), int(self.row_ptr[r1])
        return HostCsr(r1 - r0, self.cols, (self.row_ptr[r0:r1 + 1] - self.row_ptr[r0]).astype(U64),
                       self.col_idx[s:e].copy(), self.values[s:e].copy())

    def astype_values(self, val_bits: int) -> "HostCsr":
        return HostCsr(self.rows, self.cols, self.row_ptr, self.col_idx, self.values.astype(vdtype(val_bits)))


def from_coo(rows: int, cols: int, r, c, v, val_bits: int = 32, saturating: bool = False) -> HostCsr:
    """src/graph_csr.rs:83-129: sort by (row, col), sum duplicates (plain `+=`, i.e. wrapping in
    release builds; `saturating=True` gives linalg/src/csr.rs:158-195), drop zeros."""
    dt = vdtype(val_bits)
    r = np.asarray(r, dtype=np.int64).ravel()
    c = np.asarray(c, dtype=np.int64).ravel()
    v = np.asarray(v, dtype=dt).ravel()
    if r.size and (r.max() >= rows or c.max() >= cols or r.min() < 0 or c.min() < 0):
        raise IndexError("triplet index out of range")
    key = r * np.int64(cols) + c
    order = np.argsort(key, kind="stable")
    key, v = key[order], v[order]
    if key.size:
        head = np.empty(key.size, dtype=bool)
        head[0] = True
        np.not_equal(key[1:], key[:-1], out=head[1:])
        starts = np.flatnonzero(head)
        if saturating:
            exact = np.add.reduceat(v.astype(object), starts) if val_bits == 64 else np.add.reduceat(v.astype(U64), starts)
            lim = int(np.iinfo(dt).max)
            sums = np.array([min(int(x), lim) for x in exact], dtype=dt) if val_bits == 64 else np.minimum(exact, lim).astype(dt)
        else:
            sums = np.add.reduceat(v, starts).astype(dt)          # wraps like a release-mode `+=`
        ukey = key[starts]
        keep = sums != 0
        ukey, sums = ukey[keep], sums[keep]
    else:
        ukey, sums = key, v
    rr = ukey // np.int64(cols) if cols else ukey
    cc = (ukey - rr * np.int64(cols)).astype(U32)
    row_ptr = np.zeros(rows + 1, dtype=U64)
    if rr.size:
        np.cumsum(np.bincount(rr, minlength=rows), out=row_ptr[1:].view(np.int64))
    return HostCsr(rows, cols, row_ptr, cc, sums)


def identity(n: int, val_bits: int = 32) -> HostCsr:
    return HostCsr(n, n, np.arange(n + 1, dtype=U64), np.arange(n, dtype=U32), np.ones(n, dtype=vdtype(val_bits)))


def empty(n: int, val_bits: int = 32) -> HostCsr:
    return HostCsr(n, n, np.zeros(n + 1, dtype=U64), np.zeros(0, dtype=U32), np.zeros(0, dtype=vdtype(val_bits)))


def lattice(dims, torus: bool, val_bits: int = 32) -> HostCsr:
    """src/graph_csr.rs:177-222: N-d Moore neighbourhood; node ids row-major with the last
    dimension fastest; the 3^N offsets enumerated with dimension 0 as least-significant digit;
    torus wrap via rem_euclid; self offset skipped; duplicates (side-2 torus) summed."""
    dims = [int(d) for d in dims]
    nd = len(dims)
    total = int(np.prod(dims)) if nd else 1
    strides = [1] * nd
    for i in range(nd - 2, -1, -1):
        strides[i] = strides[i + 1] * dims[i + 1]
    coords = np.stack(np.unravel_index(np.arange(total, dtype=np.int64), dims), axis=1) if nd else np.zeros((1, 0), np.int64)
    rs, cs = [], []
    for off in range(3 ** nd):
        tmp, deltas = off, []
        for _ in range(nd):
            deltas.append(tmp % 3 - 1)
            tmp //= 3
        if all(d == 0 for d in deltas):
            continue
        nb = np.zeros(total, dtype=np.int64)
        valid = np.ones(total, dtype=bool)
        for d in range(nd):
            cd = coords[:, d] + deltas[d]
            if torus:
                cd = np.mod(cd, dims[d])
            else:
                valid &= (cd >= 0) & (cd < dims[d])
                cd = np.clip(cd, 0, dims[d] - 1)
            nb += cd * strides[d]
        idx = np.flatnonzero(valid)
        rs.append(idx)
        cs.append(nb[idx])
    r = np.concatenate(rs) if rs else np.zeros(0, np.int64)
    c = np.concatenate(cs) if cs else np.zeros(0, np.int64)
    return from_coo(total, total, r, c, np.ones(r.size, dtype=vdtype(val_bits)), val_bits)


# ------------------------------------------------------------------ rand 0.9.2 StdRng (ChaCha12)
def _rotl(x, n):
    return ((x << U32(n)) | (x >> U32(32 - n))).astype(U32)


def chacha12_words(seed: bytes, nblocks: int) -> np.ndarray:
    """ChaCha12 keystream words for block counters 0..nblocks-1 (stream id 0), as rand_chacha 0.9
    lays them out: block after block, 16 LE u32 words each (Cargo.lock:871-878)."""
    assert len(seed) == 32
    key = np.frombuffer(seed, dtype="<u4").astype(U32)
    ctr = np.arange(nblocks, dtype=U64)
    s = [np.full(nblocks, k, dtype=U32) for k in (0x61707865, 0x3320646E, 0x79622D32, 0x6B206574)]
    s += [np.full(nblocks, k, dtype=U32) for k in key]
    s += [(ctr & U64(0xFFFFFFFF)).astype(U32), (ctr >> U64(32)).astype(U32), np.zeros(nblocks, U32), np.zeros(nblocks, U32)]
    x = [a.copy() for a in s]

    def qr(a, b, c, d):
        x[a] = (x[a] + x[b]).astype(U32); x[d] = _rotl(x[d] ^ x[a], 16)
        x[c] = (x[c] + x[d]).astype(U32); x[b] = _rotl(x[b] ^ x[c], 12)
        x[a] = (x[a] + x[b]).astype(U32); x[d] = _rotl(x[d] ^ x[a], 8)
        x[c] = (x[c] + x[d]).astype(U32); x[b] = _rotl(x[b] ^ x[c], 7)

    with np.errstate(over="ignore"):
        for _ in range(6):
            qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
            qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
        out = np.stack([(x[i] + s[i]).astype(U32) for i in range(16)], axis=1)
    return out.reshape(-1)


def stdrng_u64(seed: bytes, n: int) -> np.ndarray:
    """First n outputs of StdRng::from_seed(seed).next_u64(): consecutive word pairs, low word first."""
    w = chacha12_words(seed, (2 * n + 15) // 16 + 1)[:2 * n].astype(U64)
    return w[0::2] | (w[1::2] << U64(32))


def stdrng_f64_unit(seed: bytes, n: int) -> np.ndarray:
    """rand 0.9 `random_range(0.0..1.0)`: (u64 >> 12) as a 52-bit mantissa -> [0,1)."""
    return (stdrng_u64(seed, n) >> U64(12)).astype(np.float64) * (1.0 / 4503599627370496.0)


def thin(a: HostCsr, density: float, seed: bytes = bytes([42] * 32), skip: int = 0) -> HostCsr:
    """src/graph_csr.rs:225-247 with StdRng::from_seed(seed): entries visited row-major; one draw per
    entry with r <= c (short-circuit `&&`, :235); a kept (r,c) also keeps its mirror (c,r) when stored.
    `skip` = draws already taken from the same generator (the sweep of bench_matmul_magnus shares one StdRng across
    its grid, src/graph_magnus.rs:800-821); this call consumes `draws_of_thin(a)` more."""
    rows = a.row_of_entry()
    cols = a.col_idx.astype(np.int64)
    upper = rows <= cols
    nu = int(upper.sum())
    draws = stdrng_f64_unit(seed, skip + nu)[skip:]
    keep_u = np.zeros(a.nnz(), dtype=bool)
    keep_u[np.flatnonzero(upper)] = draws < density
    kr, kc, kv = rows[keep_u], cols[keep_u], a.values[keep_u]
    off = kr != kc
    mr, mc = kc[off], kr[off]                                     # mirrored positions (c, r)
    # value of the stored mirror entry, 0 when absent (get(c, r), :238)
    key_all = rows * np.int64(a.cols) + cols
    pos = np.searchsorted(key_all, mr * np.int64(a.cols) + mc)
    pos = np.minimum(pos, max(a.nnz() - 1, 0))
    present = key_all[pos] == mr * np.int64(a.cols) + mc if a.nnz() else np.zeros(0, bool)
    mv = a.values[pos][present]
    r = np.concatenate([kr, mr[present]])
    c = np.concatenate([kc, mc[present]])
    v = np.concatenate([kv, mv])
    return from_coo(a.rows, a.cols, r, c, v, a.val_bits)


def draws_of_thin(a: HostCsr) -> int:
    """Generator outputs one `thin` of `a` consumes: one per stored entry on or above the diagonal."""
    return int((a.row_of_entry() <= a.col_idx.astype(np.int64)).sum())


def thinned_torus(dims, density: float, seed: bytes = bytes([42] * 32), val_bits: int = 32) -> HostCsr:
    """`lattice(dims, torus=True).thin(density)` without materialising the full lattice's COO sort -- same result
    (checked in tests) for sides >= 3, where the Moore torus has 3^N - 1 distinct unit-valued neighbours per node.
    Used for the 200^3 configuration (8 M nodes, 208 M lattice entries)."""
    dims = [int(d) for d in dims]
    if min(dims) < 3:
        return thin(lattice(dims, True, val_bits), density, seed)
    nd, total = len(dims), int(np.prod(dims))
    strides = [1] * nd
    for i in range(nd - 2, -1, -1):
        strides[i] = strides[i + 1] * dims[i + 1]
    coords = np.unravel_index(np.arange(total, dtype=np.int64), dims)
    nnb = 3 ** nd - 1
    cols = np.empty((total, nnb), dtype=np.int64 if total >= 2 ** 31 else np.int32)
    j = 0
    for off in range(3 ** nd):
        tmp, deltas = off, []
        for _ in range(nd):
            deltas.append(tmp % 3 - 1)
            tmp //= 3
        if all(d == 0 for d in deltas):
            continue
        nb = np.zeros(total, dtype=np.int64)
        for d in range(nd):
            nb += np.mod(coords[d] + deltas[d], dims[d]) * strides[d]
        cols[:, j] = nb
        j += 1
    del coords
    cols.sort(axis=1)
    rows = np.arange(total, dtype=cols.dtype)[:, None]
    upper = cols >= rows
    draws = stdrng_f64_unit(seed, int(upper.sum()))
    keep = np.zeros(upper.shape, dtype=bool)
    keep[upper] = draws < density
    del draws, upper
    kr = np.broadcast_to(rows, cols.shape)[keep].astype(np.int64)
    kc = cols[keep].astype(np.int64)
    del cols, keep
    r = np.concatenate([kr, kc])
    c = np.concatenate([kc, kr])
    return from_coo(total, total, r, c, np.ones(r.size, dtype=vdtype(val_bits)), val_bits)


def reference_bench_instance(side: int = 30, target_epn: float = 3.0, val_bits: int = 32) -> HostCsr:
    """The operand of bench_repeated_exponentiation (src/graph_magnus.rs:707-719): lattice([s,s,s],
    torus) thinned with density target_epn / full_epn under StdRng::from_seed([42; 32])."""
    full = lattice([side, side, side], True, val_bits)
    density = target_epn / (full.nnz() / full.rows)
    return thin(full, density, bytes([42] * 32))


def lattice_csr_xorshift(s: int, target_epn: float, seed: int, val_bits: int = 32) -> HostCsr:
    """linalg/benches/perf.rs:62-95 with the xorshift64 of :43-52 (sequential generator: the draws are
    produced by a Python loop, fine up to s = 30)."""
    n = s * s * s
    node = np.arange(n, dtype=np.int64)
    x, y, z = node // (s * s), (node // s) % s, node % s
    cols = []
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dz in (-1, 0, 1):
                if dx == 0 and dy == 0 and dz == 0:
                    continue
                cols.append(((x + dx) % s) * s * s + ((y + dy) % s) * s + (z + dz) % s)
    cmat = np.stack(cols, axis=1)                                  # [n, 26] in the reference's emission order
    total = n * 26
    keep_p = min(max(target_epn / 26.0, 0.0), 1.0)
    st = max(int(seed), 1)
    mask64 = (1 << 64) - 1
    draws = np.empty(total, dtype=np.float64)
    umax = float(mask64)
    for i in range(total):
        st ^= (st << 13) & mask64
        st ^= st >> 7
        st ^= (st << 17) & mask64
        draws[i] = st / umax
    keep = (draws < keep_p).reshape(n, 26)
    r = np.repeat(node, 26).reshape(n, 26)[keep]
    c = cmat[keep]
    return from_coo(n, n, r, c, np.ones(r.size, dtype=vdtype(val_bits)), val_bits, saturating=True)


def _splitmix64_at(seed: int, idx: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = U64(seed) + (idx.astype(U64) + U64(1)) * U64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> U64(30))) * U64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> U64(27))) * U64(0x94D049BB133111EB)
        return z ^ (z >> U64(31))


def rmat(scale: int, edge_factor: int, a: float, b: float, c: float, seed: int, val_bits: int = 64,
         chunk: int = 1 << 22) -> HostCsr:
    """R-MAT: 2^scale nodes, edge_factor*2^scale edges; draw number
    e*scale+l of a counter-based splitmix64 picks the quadrant of edge e at level l (bit l)."""
    n, m = 1 << scale, edge_factor << scale
    rs, cs = [], []
    for e0 in range(0, m, chunk):
        e = np.arange(e0, min(m, e0 + chunk), dtype=U64)
        r = np.zeros(e.size, dtype=np.int64)
        cc = np.zeros(e.size, dtype=np.int64)
        for l in range(scale):
            u = (_splitmix64_at(seed, e * U64(scale) + U64(l)) >> U64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
            cc |= ((u >= a) & (u < a + b) | (u >= a + b + c)).astype(np.int64) << l
            r |= (u >= a + b).astype(np.int64) << l
        rs.append(r); cs.append(cc)
    r = np.concatenate(rs); cc = np.concatenate(cs)
    return from_coo(n, n, r, cc, np.ones(r.size, dtype=vdtype(val_bits)), val_bits)
