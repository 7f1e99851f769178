"""Regenerates tests/golden/*.npz.

The reference is Rust (nightly, git dependencies) and cannot be executed in the build image, so these
fixtures are produced by the CPU restatement in oracle/ -- which is itself pinned to the reference's own
known-answer tests and to the nnz column of the reference README (tests/test_oracle_*.py) -- and every
product is cross-checked here against scipy.sparse (an independent implementation; valid because none of
these cases saturates) before it is written.  Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O  # noqa: E402


def to_sp(m):
    return sp.csr_matrix((m.values.astype(np.uint64), m.col_idx.astype(np.int64), m.row_ptr.astype(np.int64)), shape=(m.rows, m.cols))


def check_scipy(a, b, c):
    ref = (to_sp(a) @ to_sp(b)).tocsr()
    ref.sort_indices()
    assert np.array_equal(ref.indptr.astype(np.uint64), c.row_ptr)
    assert np.array_equal(ref.indices.astype(np.uint32), c.col_idx)
    assert np.array_equal(ref.data.astype(np.uint64), c.values.astype(np.uint64))


def save(name, **mats):
    flat = {}
    for k, m in mats.items():
        flat[k + "_shape"] = np.array([m.rows, m.cols], dtype=np.uint64)
        flat[k + "_row_ptr"], flat[k + "_col_idx"], flat[k + "_values"] = m.row_ptr, m.col_idx, m.values
    np.savez_compressed(os.path.join(HERE, name), **flat)


# 1. reference bench instance at side 8 (StdRng([42;32]) thinning), powers A^2..A^5, u32 and u64
for bits in (32, 64):
    a = O.reference_bench_instance(8, 3.0, bits)
    mats, p = {"A": a}, a
    for k in range(2, 6):
        c = O.matmul(p, a)
        assert O.matmul_par(p, a, 4).equals(c)
        check_scipy(p, a, c)
        mats[f"A{k}"] = c
        p = c
    save(f"torus8_chain_u{bits}.npz", **mats)

# 2. sweep cell: full Moore torus side 6 (26 e/n), A x A
full = O.lattice([6, 6, 6], True, 64)
c = O.matmul(full, full)
check_scipy(full, full, c)
save("torus6_full_u64.npz", A=full, AA=c)

# 3. portable xorshift torus (linalg/benches/perf.rs lattice_csr(10, 3.0, 42)), A x A
x = O.lattice_csr_xorshift(10, 3.0, 42, 32)
c = O.matmul(x, x)
check_scipy(x, x, c)
save("xorshift10_u32.npz", A=x, AA=c)

# 4. R-MAT scale 10, A x A (skewed rows)
r = O.rmat(10, 8, 0.57, 0.19, 0.19, 42, 64)
c = O.matmul(r, r)
check_scipy(r, r, c)
save("rmat10_u64.npz", A=r, AA=c)

# 5. saturation: 64-node chain (A + I) squared repeatedly (src/graph_csr.rs:931-939), u32; no scipy check
n = 64
a = O.add(O.from_edges(n, [(i, i + 1) for i in range(n - 1)], 32), O.identity(n, 32))
mats, p = {"M0": a}, a
for k in range(1, 9):
    p = O.matmul(p, p)
    mats[f"M{k}"] = p
assert int(p.values.max()) == 0xFFFFFFFF
save("chain64_saturating_u32.npz", **mats)
print("golden fixtures written:", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))
