"""The reference's own known-answer tests, run against the CPU oracle (oracle/spgemm_oracle.c).

Each test names the reference test it restates.  These pin the oracle; the GPU parity tests then
compare the CUDA engine with the oracle.  Both value widths: u32 = CsrMatrix (src/graph_csr.rs),
u64 = MagnusMatrix / Sat64 (src/graph_magnus.rs, src/graph_sprs.rs), Csr<u32,u64> (linalg/src/csr.rs).
"""
import numpy as np
import pytest

BITS = [32, 64]


def reachability_sum(O, m):
    power, total, k = m, m, 1                     # src/graph_csr.rs:545-558
    while True:
        power = O.matmul(power, m); k += 1
        new = O.add(total, power)
        if new.nnz() == total.nnz():
            return new, k
        total = new


def power_until_stable(O, m):
    cur, k = m, 0                                 # src/graph_csr.rs:561-575
    while True:
        nxt = O.matmul(cur, cur); k += 1
        if nxt.nnz() == cur.nnz() and np.array_equal(nxt.col_idx, cur.col_idx) and np.array_equal(nxt.row_ptr, cur.row_ptr):
            return nxt, k
        cur = nxt


def components(O, m):
    closure, _ = power_until_stable(O, O.add(m, O.identity(m.rows, m.val_bits)))   # src/graph_csr.rs:578-600
    comp, nxt = [-1] * m.rows, 0
    for i in range(m.rows):
        if comp[i] != -1:
            continue
        comp[i] = nxt
        for j in range(i + 1, m.rows):
            if closure.get(i, j) > 0 and closure.get(j, i) > 0:
                comp[j] = nxt
        nxt += 1
    return comp


@pytest.mark.parametrize("bits", BITS)
def test_identity_matmul(oracle, bits):                       # graph_csr.rs:878-887, graph_magnus.rs:455-464, linalg csr.rs:797-806
    m = oracle.from_edges(3, [(0, 1), (1, 2)], bits)
    r = oracle.matmul(m, oracle.identity(3, bits))
    assert (r.get(0, 1), r.get(1, 2), r.get(0, 2), r.nnz()) == (1, 1, 0, 2)


@pytest.mark.parametrize("bits", BITS)
def test_path_counting_triangle(oracle, bits):                # graph_csr.rs:890-900, linalg csr.rs:809-818
    m = oracle.from_edges(3, [(0, 1), (1, 2), (2, 0)], bits)
    m2 = oracle.matmul(m, m)
    assert (m2.get(0, 2), m2.get(1, 0), m2.get(2, 1), m2.get(0, 0), m2.nnz()) == (1, 1, 1, 0, 3)
    m3 = oracle.matmul(m2, m)
    assert (m3.get(0, 0), m3.get(1, 1), m3.get(2, 2)) == (1, 1, 1)


@pytest.mark.parametrize("bits", BITS)
def test_parallel_paths_and_diamond(oracle, bits):            # graph_csr.rs:903-915, linalg csr.rs:821-826
    assert oracle.from_edges(2, [(0, 1), (0, 1)], bits).get(0, 1) == 2
    d = oracle.from_edges(4, [(0, 1), (0, 2), (1, 3), (2, 3)], bits)
    assert oracle.matmul(d, d).get(0, 3) == 2


@pytest.mark.parametrize("bits", BITS)
def test_from_coo_dedups_and_drops_zeros(oracle, bits):       # linalg csr.rs:839-849; graph_csr.rs:108
    m = oracle.from_coo(3, 3, [0, 0, 1, 2], [1, 1, 2, 2], [2, 3, 1, 0], bits)
    assert (m.get(0, 1), m.get(1, 2), m.nnz()) == (5, 1, 2)


@pytest.mark.parametrize("bits", BITS)
def test_add_with_overlap(oracle, bits):                      # linalg csr.rs:829-837
    a = oracle.from_edges(3, [(0, 1), (1, 2)], bits)
    b = oracle.from_edges(3, [(0, 1), (2, 0)], bits)
    c = oracle.add(a, b)
    assert (c.get(0, 1), c.get(1, 2), c.get(2, 0), c.nnz()) == (2, 1, 1, 3)


@pytest.mark.parametrize("bits", BITS)
def test_reachability_chain(oracle, bits):                    # graph_csr.rs:918-929
    s, _ = reachability_sum(oracle, oracle.from_edges(4, [(0, 1), (1, 2), (2, 3)], bits))
    assert all(s.get(i, j) > 0 for i, j in [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)])
    assert s.get(3, 0) == 0 and s.get(2, 0) == 0


def test_power_until_stable_chain_saturates(oracle):          # graph_csr.rs:931-939 (u32 saturates on purpose)
    n = 64
    m = oracle.from_edges(n, [(i, i + 1) for i in range(n - 1)], 32)
    stable, iters = power_until_stable(oracle, oracle.add(m, oracle.identity(n, 32)))
    assert iters <= 8
    assert int(stable.values.max()) == 0xFFFFFFFF


@pytest.mark.parametrize("bits", BITS)
def test_connected_components(oracle, bits):                  # graph_csr.rs:942-961,1096-1104; linalg csr.rs:852-864
    m = oracle.from_edges_undirected(6, [(0, 1), (1, 2), (2, 0), (3, 4), (4, 5), (5, 3)], bits)
    c = components(oracle, m)
    assert c[0] == c[1] == c[2] and c[3] == c[4] == c[5] and c[0] != c[3]
    assert len(set(components(oracle, oracle.empty(5, bits)))) == 5
    c = components(oracle, oracle.from_edges_undirected(4, [(0, 1), (1, 2), (2, 3)], bits))
    assert c[0] == c[1] == c[2] == c[3]


@pytest.mark.parametrize("bits", BITS)
def test_lattices(oracle, bits):                              # graph_csr.rs:1011-1093
    m = oracle.lattice([5], False, bits)
    assert (m.rows, m.get(0, 1), m.get(1, 0), m.get(0, 0), m.get(4, 3), m.get(4, 0), m.nnz()) == (5, 1, 1, 0, 1, 0, 8)
    t = oracle.lattice([5], True, bits)
    assert (t.get(0, 4), t.get(4, 0), t.nnz()) == (1, 1, 10)
    g = oracle.lattice([3, 3], False, bits)
    assert (g.rows, g.get(0, 1), g.get(0, 3), g.get(0, 4)) == (9, 1, 1, 1) and g.nnz() % 2 == 0
    assert sum(1 for j in range(9) if g.get(4, j) > 0) == 8
    gt = oracle.lattice([3, 3], True, bits)
    assert all(sum(1 for j in range(9) if gt.get(i, j) > 0) == 8 for i in range(9)) and gt.nnz() == 72
    c = oracle.lattice([2, 2, 2], False, bits)
    assert c.rows == 8 and sum(1 for j in range(8) if c.get(0, j) > 0) == 7 and c.nnz() == 56
    s = oracle.lattice([4, 3], False, bits)
    for r in range(s.rows):
        for k in range(int(s.row_ptr[r]), int(s.row_ptr[r + 1])):
            assert s.get(int(s.col_idx[k]), r) == int(s.values[k])
    assert len(set(components(oracle, oracle.lattice([4, 4], False, bits)))) == 1


@pytest.mark.parametrize("bits", BITS)
def test_matmul_par_agrees_with_seq(oracle, bits):            # linalg csr.rs:973-988, graph_magnus.rs:689-697
    edges = [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (5, 6), (6, 7), (7, 8), (8, 9), (9, 0), (0, 5), (1, 6), (2, 7)]
    a = oracle.from_edges(10, edges, bits)
    assert oracle.matmul(a, a).equals(oracle.matmul_par(a, a, 4))
    d = oracle.from_edges(4, [(0, 1), (0, 2), (1, 3), (2, 3)], bits)
    assert oracle.matmul_par(d, d, 2).get(0, 3) == 2


def test_dense_2x2_through_builder_values(oracle):            # graph_csr_builder.rs:163-175
    a = oracle.from_coo(2, 2, [0, 0, 1, 1], [0, 1, 0, 1], [1, 2, 3, 4], 32)
    b = oracle.from_coo(2, 2, [0, 0, 1, 1], [0, 1, 0, 1], [5, 6, 7, 8], 32)
    c = oracle.matmul(a, b)
    assert [c.get(0, 0), c.get(0, 1), c.get(1, 0), c.get(1, 1)] == [19, 22, 43, 50]


def test_saturating_scalar_ops(oracle):                       # graph_csr.rs:30-37, graph_sprs.rs:29-51
    L = oracle.lib()
    assert L.oracle_sadd_u32(0xFFFFFFFF, 1) == 0xFFFFFFFF and L.oracle_sadd_u32(7, 8) == 15
    assert L.oracle_smul_u32(0x10000, 0x10000) == 0xFFFFFFFF and L.oracle_smul_u32(3, 5) == 15
    assert L.oracle_sadd_u64(0xFFFFFFFFFFFFFFFF, 5) == 0xFFFFFFFFFFFFFFFF
    assert L.oracle_smul_u64(1 << 40, 1 << 40) == 0xFFFFFFFFFFFFFFFF and L.oracle_smul_u64(1 << 20, 1 << 20) == 1 << 40


def test_matmul_against_scipy_random(oracle):
    """Independent cross-check (no saturation): scipy.sparse product with sorted indices."""
    import scipy.sparse as sp
    rng = np.random.default_rng(0)
    for bits in BITS:
        for n, nnz in ((50, 300), (400, 4000), (3000, 20000)):
            a = oracle.from_coo(n, n, rng.integers(0, n, nnz), rng.integers(0, n, nnz), rng.integers(1, 5, nnz), bits)
            b = oracle.from_coo(n, n, rng.integers(0, n, nnz), rng.integers(0, n, nnz), rng.integers(1, 5, nnz), bits)
            c = oracle.matmul(a, b)
            assert oracle.matmul_par(a, b, 3).equals(c)
            ref = (a.to_scipy() @ b.to_scipy()).tocsr(); ref.sort_indices()
            assert np.array_equal(ref.indptr.astype(np.uint64), c.row_ptr) and np.array_equal(ref.indices.astype(np.uint32), c.col_idx)
            assert np.array_equal(ref.data.astype(np.uint64), c.values.astype(np.uint64))


# ------------------------------------------------------------------ locality pre-pass (src/graph_csr.rs:663-818)
def _same(a, b):
    return np.array_equal(a.row_ptr, b.row_ptr) and np.array_equal(a.col_idx, b.col_idx) and np.array_equal(a.values, b.values)


def _inverse(perm):
    inv = np.empty_like(perm)
    inv[perm] = np.arange(perm.size, dtype=perm.dtype)
    return inv


@pytest.mark.parametrize("which", ["small", "lattice", "directed"])
def test_rcm_unpermute_roundtrip(oracle, which):
    O = oracle
    """test_rcm_unpermute_roundtrip / _lattice / _directed (src/graph_csr.rs:1107-1146): rcm then unpermute restores the
    three arrays; the order is a permutation."""
    if which == "small":
        a = O.from_edges_undirected(6, [(0, 3), (1, 4), (2, 5), (0, 1), (3, 4)])
    elif which == "lattice":
        a = O.lattice([4, 4], False)
    else:
        a = O.from_edges(5, [(0, 1), (1, 2), (2, 3), (3, 4), (4, 0), (0, 3)])
    perm = O.rcm_order(a)
    assert sorted(perm.tolist()) == list(range(a.rows))
    p = O.permute(a, perm)
    assert p.nnz() == a.nnz()
    assert _same(O.permute(p, _inverse(perm)), a)


def test_bandwidth_stats_and_rcm_on_a_shuffled_band(oracle):
    O = oracle
    """A path graph relabelled at random has bandwidth ~n; Cuthill-McKee from a pseudo-peripheral end brings it back to 1."""
    n = 200
    rng = np.random.default_rng(5)
    lab = rng.permutation(n)
    a = O.from_edges_undirected(n, [(int(lab[i]), int(lab[i + 1])) for i in range(n - 1)])
    mx, avg = O.bandwidth_stats(a)
    assert mx > 50 and avg > 10
    p = O.permute(a, O.rcm_order(a))
    assert O.bandwidth_stats(p) == (1, 1.0)
    assert O.bandwidth_stats(O.empty(4)) == (0, 0.0)
