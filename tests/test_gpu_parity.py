"""Parity tests proper: the CUDA engine, called through the C ABI, against the CPU oracle.

Bar: bit-exact row_ptr / col_idx / values (integer path).  Cases follow the reference's own tests
(src/graph_csr.rs:878-1104, src/graph_magnus.rs:455-697, linalg/src/csr.rs:796-989) plus the bins
of the engine (tiny / hash / heavy, bitmap-rank and sort emission, 32/64-bit/saturating accumulators).
"""
import numpy as np
import pytest

from sparse_linear_algebra_tests_b200 import B200Matrix, ShapeMismatch, hostgen

pytestmark = pytest.mark.gpu


def to_o(O, h):
    return O.Csr(h.rows, h.cols, h.row_ptr, h.col_idx, h.values)


def assert_same(got, want, what=""):
    assert got.rows == want.rows and got.cols == want.cols, what
    assert np.array_equal(got.row_ptr, want.row_ptr), f"row_ptr differs {what}"
    assert np.array_equal(got.col_idx, want.col_idx), f"col_idx differs {what}"
    assert got.values.dtype == want.values.dtype, what
    assert np.array_equal(got.values, want.values), f"values differ {what}"


def check_product(O, ctx, a_h, b_h, what=""):
    a, b = B200Matrix.from_host(a_h, ctx), B200Matrix.from_host(b_h, ctx)
    c = a.matmul(b, want_stats=True)
    want = O.matmul(to_o(O, a_h), to_o(O, b_h))
    assert_same(c.to_host(), want, what)
    assert c.nnz() == want.nnz()
    assert c.last_stats.products == int(O.row_products(to_o(O, a_h), to_o(O, b_h)).sum())
    return c


def rand_csr(rng, rows, cols, nnz, bits, vmax=3):
    r = rng.integers(0, rows, size=nnz)
    c = rng.integers(0, cols, size=nnz)
    v = rng.integers(1, vmax + 1, size=nnz, dtype=np.uint64)
    return hostgen.from_coo(rows, cols, r, c, v.astype(hostgen.vdtype(bits)), bits)


# ------------------------------------------------------------------ reference KATs through the engine
@pytest.mark.parametrize("bits", [32, 64])
def test_identity_matmul(gpu_ctx, bits):
    m = B200Matrix.from_edges(3, [(0, 1), (1, 2)], bits)
    r = m.matmul(B200Matrix.identity(3, bits))
    assert (r.get(0, 1), r.get(1, 2), r.get(0, 2), r.nnz()) == (1, 1, 0, 2)


@pytest.mark.parametrize("bits", [32, 64])
def test_path_counting_triangle_diamond(gpu_ctx, bits):
    m = B200Matrix.from_edges(3, [(0, 1), (1, 2), (2, 0)], bits)
    m2 = m.matmul(m)
    assert (m2.get(0, 2), m2.get(1, 0), m2.get(2, 1), m2.nnz()) == (1, 1, 1, 3)
    m3 = m2.matmul(m)
    assert (m3.get(0, 0), m3.get(1, 1), m3.get(2, 2)) == (1, 1, 1)
    d = B200Matrix.from_edges(4, [(0, 1), (0, 2), (1, 3), (2, 3)], bits)
    assert d.matmul(d).get(0, 3) == 2
    assert B200Matrix.from_edges(2, [(0, 1), (0, 1)], bits).get(0, 1) == 2


def test_matmul_par_agrees_with_seq_10_node(gpu_ctx, oracle):
    edges = [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (5, 6), (6, 7), (7, 8), (8, 9), (9, 0), (0, 5), (1, 6), (2, 7)]
    a_h = hostgen.from_coo(10, 10, [e[0] for e in edges], [e[1] for e in edges], np.ones(13, np.uint32), 32)
    check_product(oracle, gpu_ctx, a_h, a_h)


def test_dense_2x2(gpu_ctx):
    a = B200Matrix.from_coo(2, [(0, 0, 1), (0, 1, 2), (1, 0, 3), (1, 1, 4)])
    b = B200Matrix.from_coo(2, [(0, 0, 5), (0, 1, 6), (1, 0, 7), (1, 1, 8)])
    c = a.matmul(b)
    assert [c.get(0, 0), c.get(0, 1), c.get(1, 0), c.get(1, 1)] == [19, 22, 43, 50]


def test_shape_mismatch_panics(gpu_ctx):
    with pytest.raises(ShapeMismatch):
        B200Matrix.identity(3).matmul(B200Matrix.identity(4))
    with pytest.raises(AssertionError):
        B200Matrix.identity(3, 32).matmul(B200Matrix.identity(3, 64))


def test_reachability_power_components(gpu_ctx):
    m = B200Matrix.from_edges(4, [(0, 1), (1, 2), (2, 3)])
    s, _ = m.reachability_sum()
    assert all(s.get(i, j) > 0 for i, j in [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)])
    assert s.get(3, 0) == 0 and s.get(2, 0) == 0
    n = 64
    chain = B200Matrix.from_edges(n, [(i, i + 1) for i in range(n - 1)], 32)
    _, iters = chain.add(B200Matrix.identity(n, 32)).power_until_stable()
    assert iters <= 8
    tri2 = B200Matrix.from_edges_undirected(6, [(0, 1), (1, 2), (2, 0), (3, 4), (4, 5), (5, 3)])
    comp = tri2.connected_components()
    assert comp[0] == comp[1] == comp[2] and comp[3] == comp[4] == comp[5] and comp[0] != comp[3]
    assert len(set(B200Matrix.new(5).connected_components())) == 5
    assert tri2.num_components() == 2


# ------------------------------------------------------------------ saturation (graph_csr.rs:931-939 saturates on purpose)
@pytest.mark.parametrize("bits", [32, 64])
def test_chain_closure_saturates_like_oracle(gpu_ctx, oracle, bits):
    n = 64
    e = [(i, i + 1) for i in range(n - 1)]
    a_h = hostgen.from_coo(n, n, [x[0] for x in e] + list(range(n)), [x[1] for x in e] + list(range(n)), np.ones(2 * n - 1, hostgen.vdtype(bits)), bits)
    cur, cur_o = B200Matrix.from_host(a_h, gpu_ctx), to_o(oracle, a_h)
    for _ in range(9 if bits == 32 else 12):
        cur = cur.matmul(cur)
        cur_o = oracle.matmul(cur_o, cur_o)
        assert_same(cur.to_host(), cur_o)
    lim = (1 << bits) - 1
    if bits == 32:
        assert int(cur_o.values.max()) == lim


@pytest.mark.parametrize("bits,vmax", [(32, 70000), (32, 0xFFFFFFFF), (64, 1 << 33), (64, 1 << 62), (64, 0xFFFFFFFFFFFFFFFF)])
def test_large_values_all_accumulator_modes(gpu_ctx, oracle, bits, vmax):
    rng = np.random.default_rng(7)
    n = 600
    a_h = rand_csr(rng, n, n, 6000, bits, 3)
    a_h.values[:] = rng.integers(max(1, vmax // 3), vmax, size=a_h.nnz(), dtype=np.uint64, endpoint=True).astype(a_h.values.dtype)
    b_h = rand_csr(rng, n, n, 9000, bits, 3)
    b_h.values[::2] = np.array(vmax, dtype=b_h.values.dtype)
    c = check_product(oracle, gpu_ctx, a_h, b_h, f"u{bits} vmax={vmax}")
    if vmax >= (1 << bits) - 1:
        assert int(c.values.max()) == (1 << bits) - 1


# ------------------------------------------------------------------ the benchmark family at oracle-friendly sizes
@pytest.mark.parametrize("bits", [32, 64])
@pytest.mark.parametrize("side,epn", [(5, 2), (5, 26), (10, 3), (10, 8), (10, 26), (20, 4)])
def test_sweep_grid_a_times_a(gpu_ctx, oracle, side, epn, bits):
    full = hostgen.lattice([side] * 3, True, bits)
    a_h = full if epn >= 26 else hostgen.thin(full, epn / 26.0)
    check_product(oracle, gpu_ctx, a_h, a_h, f"side={side} epn={epn}")


@pytest.mark.parametrize("bits", [32, 64])
def test_repeated_exponentiation_chain(gpu_ctx, oracle, bits):
    a_h = hostgen.reference_bench_instance(12, 3.0, bits)
    a, a_o = B200Matrix.from_host(a_h, gpu_ctx), to_o(oracle, a_h)
    p, p_o = a, a_o
    for k in range(2, 8):
        p = p.matmul(a)
        p_o = oracle.matmul_par(p_o, a_o)
        assert_same(p.to_host(), p_o, f"A^{k}")


def test_reference_instance_30_first_powers(gpu_ctx, oracle):
    a_h = hostgen.reference_bench_instance(30, 3.0, 64)
    assert a_h.nnz() == 81434
    a, a_o = B200Matrix.from_host(a_h, gpu_ctx), to_o(oracle, a_h)
    p, p_o = a, a_o
    for k, nnz in zip(range(2, 6), (251590, 655391, 1574848, 3383207)):
        p = p.matmul(a)
        p_o = oracle.matmul_par(p_o, a_o)
        assert p.nnz() == nnz
        assert_same(p.to_host(), p_o, f"A^{k}")


# ------------------------------------------------------------------ bins and edge cases
def test_empty_and_degenerate(gpu_ctx, oracle):
    for bits in (32, 64):
        z = hostgen.empty(7, bits)
        i7 = hostgen.identity(7, bits)
        for x, y in ((z, z), (z, i7), (i7, z)):
            c = check_product(oracle, gpu_ctx, x, y)
            assert c.nnz() == 0
    one = hostgen.from_coo(1, 1, [0], [0], np.array([5], np.uint64), 64)
    assert check_product(oracle, gpu_ctx, one, one).get(0, 0) == 25


def test_rows_pointing_at_empty_b_rows(gpu_ctx, oracle):
    rng = np.random.default_rng(3)
    n = 500
    a_h = rand_csr(rng, n, n, 40000, 64)           # ~80 entries per row
    b_h = rand_csr(rng, n, n, 300, 64)             # most B rows empty: P <= 32 with deg_A > 32
    check_product(oracle, gpu_ctx, a_h, b_h)


@pytest.mark.parametrize("bits", [32, 64])
@pytest.mark.parametrize("n,nnz_a,nnz_b", [(200, 1000, 1000), (3000, 30000, 60000), (20000, 200000, 400000), (400000, 1200000, 2000000)])
def test_random_graphs(gpu_ctx, oracle, n, nnz_a, nnz_b, bits):
    rng = np.random.default_rng(n)
    check_product(oracle, gpu_ctx, rand_csr(rng, n, n, nnz_a, bits), rand_csr(rng, n, n, nnz_b, bits), f"n={n}")


@pytest.mark.parametrize("bits", [32, 64])
def test_ragged_rows_hit_every_hash_bin(gpu_ctx, oracle, bits):
    """Row lengths 1..20000 against a fairly dense B: exercises every hash bin, bitmap-rank emission
    (small column space) and the heavy path."""
    rng = np.random.default_rng(11)
    n = 30000
    lens = [1, 2, 5, 20, 33, 64, 100, 200, 400, 800, 1500, 3000, 6000, 12000, 20000, 0, 7]
    r = np.concatenate([np.full(l, i) for i, l in enumerate(lens)])
    c = np.concatenate([rng.choice(n, size=l, replace=False) for l in lens])
    a_h = hostgen.from_coo(n, n, r, c, rng.integers(1, 4, size=r.size, dtype=np.uint64).astype(hostgen.vdtype(bits)), bits)
    b_h = rand_csr(rng, n, n, 5 * n, bits)
    c = check_product(oracle, gpu_ctx, a_h, b_h)
    assert c.last_stats.sym_bin_rows[9] > 0, "heavy numeric bin not exercised"


def test_wide_column_space_sort_emission_and_global_heavy(gpu_ctx, oracle):
    """n = 2.5M columns: no shared-memory bitmap anywhere (sort emission, global-bitmap heavy rows)."""
    rng = np.random.default_rng(5)
    n = 2_500_000
    lens = [3, 40, 150, 700, 2500, 9000, 30000]
    r = np.concatenate([np.full(l, i * 1000) for i, l in enumerate(lens)])
    c = np.concatenate([rng.choice(n, size=l, replace=False) for l in lens])
    a_h = hostgen.from_coo(n, n, r, c, np.ones(r.size, np.uint64), 64)
    b_h = rand_csr(rng, n, n, 4 * n, 64)
    c = check_product(oracle, gpu_ctx, a_h, b_h)
    assert c.last_stats.sym_bin_rows[9] > 0


def test_rectangular_row_block_matches_rows_of_full_product(gpu_ctx, oracle):
    a_h = hostgen.reference_bench_instance(10, 4.0, 64)
    a = B200Matrix.from_host(a_h, gpu_ctx)
    full = a.matmul(a).to_host()
    cuts = gpu_ctx.shard_rows_by_products(a.device, a.device, 3)
    assert cuts[0] == 0 and cuts[-1] == a_h.rows and np.all(np.diff(cuts.astype(np.int64)) >= 0)
    prods = gpu_ctx.row_products(a.device, a.device)
    assert np.array_equal(prods, oracle.row_products(to_o(oracle, a_h), to_o(oracle, a_h)))
    per_part = [int(prods[int(cuts[i]):int(cuts[i + 1])].sum()) for i in range(3)]
    assert max(per_part) - min(per_part) <= 2 * int(prods.max()) + 3
    for i in range(3):
        r0, r1 = int(cuts[i]), int(cuts[i + 1])
        blk = B200Matrix(gpu_ctx.row_block(a.device, r0, r1))
        assert_same(blk.to_host(), a_h.row_block(r0, r1))
        got = blk.matmul(a).to_host()
        assert_same(got, full.row_block(r0, r1), f"block {i}")


def test_magnus_layout_usize_columns_roundtrip(gpu_ctx):
    a_h = hostgen.reference_bench_instance(6, 3.0, 64)
    m = B200Matrix.from_parts(a_h.rows, a_h.cols, a_h.row_ptr, a_h.col_idx.astype(np.uint64), a_h.values, gpu_ctx)
    rp, ci, vv = m.device.download(idx64=True)
    assert ci.dtype == np.uint64 and np.array_equal(ci, a_h.col_idx) and np.array_equal(rp, a_h.row_ptr) and np.array_equal(vv, a_h.values)


def test_upload_rejects_explicit_zero(gpu_ctx):
    from sparse_linear_algebra_tests_b200 import B200Error
    with pytest.raises(B200Error):
        gpu_ctx.upload(2, 2, np.array([0, 1, 2], np.uint64), np.array([0, 1], np.uint32), np.array([1, 0], np.uint64))
    with pytest.raises(B200Error):
        gpu_ctx.upload(2, 2, np.array([0, 1, 2], np.uint64), np.array([0, 5], np.uint32), np.array([1, 1], np.uint64))


@pytest.mark.parametrize("bits", [32, 64])
def test_add_matches_oracle(gpu_ctx, oracle, bits):
    rng = np.random.default_rng(2)
    a_h, b_h = rand_csr(rng, 3000, 3000, 20000, bits), rand_csr(rng, 3000, 3000, 25000, bits)
    a_h.values[:50] = np.array((1 << bits) - 1, dtype=a_h.values.dtype)
    got = B200Matrix.from_host(a_h, gpu_ctx).add(B200Matrix.from_host(b_h, gpu_ctx)).to_host()
    assert_same(got, oracle.add(to_o(oracle, a_h), to_o(oracle, b_h)))


# ------------------------------------------------------------------ size-independent properties at larger scale
def test_properties_on_full_size_torus(gpu_ctx):
    """30^3 reference instance up to A^7: nnz matches the reference README table (README.md:42-47, rounded),
    row sums obey (A^k 1) = A (A^(k-1) 1), symmetry of A^k, sorted unique columns."""
    a_h = hostgen.reference_bench_instance(30, 3.0, 64)
    a = B200Matrix.from_host(a_h, gpu_ctx)
    p = a
    readme = {2: "252k", 3: "655k", 4: "1.57M", 5: "3.38M", 6: "6.59M", 7: "11.7M"}
    exact = {2: 251590, 3: 655391, 4: 1574848, 5: 3383207, 6: 6590100, 7: 11736555}
    rowsum_prev = np.add.reduceat(a_h.values.astype(np.float64), a_h.row_ptr[:-1].astype(np.int64)) * (np.diff(a_h.row_ptr.astype(np.int64)) > 0)
    import scipy.sparse as sp
    a_sp = sp.csr_matrix((a_h.values.astype(np.float64), a_h.col_idx.astype(np.int64), a_h.row_ptr.astype(np.int64)), shape=(a_h.rows, a_h.cols))
    for k in range(2, 8):
        p = p.matmul(a)
        assert p.nnz() == exact[k]
        s = f"{p.nnz() / 1e3:.0f}k" if p.nnz() < 1e6 else f"{p.nnz() / 1e6:.3g}M"
        assert s == readme[k]
        h = p.to_host()
        lens = np.diff(h.row_ptr.astype(np.int64))
        rows = np.repeat(np.arange(h.rows), lens)
        key = rows * h.cols + h.col_idx.astype(np.int64)
        assert np.all(np.diff(key) > 0), "columns not strictly ascending within rows"
        rowsum = np.bincount(rows, weights=h.values.astype(np.float64), minlength=h.rows)
        # A^k 1 = A^(k-1) (A 1); with symmetric A also = A (A^(k-1) 1)
        assert np.allclose(rowsum, a_sp @ rowsum_prev, rtol=0, atol=0.5)
        rowsum_prev = rowsum
        m = sp.csr_matrix((h.values.astype(np.float64), h.col_idx.astype(np.int64), h.row_ptr.astype(np.int64)), shape=(h.rows, h.cols))
        assert (m != m.T).nnz == 0, "A^k of a symmetric A must be symmetric"


# ------------------------------------------------------------------ committed golden fixtures (tests/golden/*.npz)
import glob as _glob
import os as _os

_GOLD = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "golden")


def _load_golden(path):
    z = np.load(path)
    names = sorted({k[:-len("_shape")] for k in z.files if k.endswith("_shape")})
    return {n: hostgen.HostCsr(int(z[n + "_shape"][0]), int(z[n + "_shape"][1]), z[n + "_row_ptr"], z[n + "_col_idx"], z[n + "_values"]) for n in names}


@pytest.mark.parametrize("path", sorted(_glob.glob(_os.path.join(_GOLD, "*.npz"))), ids=_os.path.basename)
def test_engine_reproduces_golden_fixture(gpu_ctx, path):
    m = _load_golden(path)
    up = lambda h: B200Matrix.from_host(h, gpu_ctx)
    if "AA" in m:
        assert_same(up(m["A"]).matmul(up(m["A"])).to_host(), m["AA"])
    elif "A2" in m:
        a = up(m["A"]); p = a
        for k in range(2, 6):
            p = p.matmul(a)
            assert_same(p.to_host(), m[f"A{k}"], f"A^{k}")
    else:
        p = up(m["M0"])
        for k in range(1, 9):
            p = p.matmul(p)
            assert_same(p.to_host(), m[f"M{k}"], f"squaring {k}")


def test_two_pass_fallback_path_is_bit_identical(gpu_ctx, oracle, cfg):
    """placement=1 (b200_config) forces the exact-allocation path (counts -> row_ptr -> numeric) used
    when the scratch CSR would not fit; both must give the same bytes."""
    a_h = hostgen.reference_bench_instance(12, 3.0, 64)
    a = B200Matrix.from_host(a_h, gpu_ctx)
    one = a.matmul(a).matmul(a).to_host()
    cfg(pipeline=2, placement=1)
    two = a.matmul(a).matmul(a).to_host()
    assert_same(one, two)
    assert_same(one, oracle.matmul(oracle.matmul(to_o(oracle, a_h), to_o(oracle, a_h)), to_o(oracle, a_h)))


# ------------------------------------------------------------------ column windows: bitmap path vs hash path of the one-pass numeric
@pytest.mark.parametrize("wincap", ["0", "1", "3", "16"])
@pytest.mark.parametrize("bits", [32, 64])
def test_column_window_split_is_bit_identical(gpu_ctx, oracle, cfg, wincap, bits):
    """Rows whose column window exceeds their bin's shared-memory bitmap leave k_num_expand for the hash + sort kernels.
    window_cap_groups shrinks the window (in 128-column groups; 0 = hash only) so that both lists are populated on a small
    torus; every split must give the reference's bytes."""
    cfg(pipeline=2, window_cap_groups=int(wincap))
    a_h = hostgen.reference_bench_instance(12, 3.0, bits)
    a, a_o = B200Matrix.from_host(a_h, gpu_ctx), to_o(oracle, a_h)
    p, p_o = a, a_o
    for k in range(2, 6):
        p = p.matmul(a)
        p_o = oracle.matmul(p_o, a_o)
        assert_same(p.to_host(), p_o, f"A^{k} wincap={wincap}")


def test_window_of_a_banded_matrix_far_from_column_zero(gpu_ctx, oracle):
    """The bitmap is relative to the row's first column (a multiple of 128): a band placed at large column indices in a
    wide, rectangular right operand must land in the same places."""
    rng = np.random.default_rng(11)
    rows, inner, cols = 300, 500, 3_000_000
    a_h = rand_csr(rng, rows, inner, 9000, 64, 5)
    r = rng.integers(0, inner, size=20000)
    c = 2_900_000 + (r * 37 + rng.integers(0, 900, size=20000)) % 90_000          # banded, near the right edge
    b_h = hostgen.from_coo(inner, cols, r, c, rng.integers(1, 4, size=20000).astype(np.uint64), 64)
    check_product(oracle, gpu_ctx, a_h, b_h, "band at columns 2.9M..3.0M")


# ------------------------------------------------------------------ exact mode (count pass, C written once) vs one-pass scratch + compaction
@pytest.mark.parametrize("exact", ["0", "1"])
@pytest.mark.parametrize("bits", [32, 64])
def test_exact_and_scratch_modes_give_the_same_bytes(gpu_ctx, oracle, cfg, exact, bits):
    """placement=1: every list gets a count kernel first and the numeric kernels write C at its final offsets;
    placement=0: rows go to a scratch CSR at bound offsets and are compacted.  The engine picks by size; both are
    forced here over the torus chain, ragged rows through every bin including the heavy one, and a wide column space."""
    cfg(pipeline=2, placement=int(exact))
    a_h = hostgen.reference_bench_instance(12, 3.0, bits)
    a, a_o = B200Matrix.from_host(a_h, gpu_ctx), to_o(oracle, a_h)
    p, p_o = a, a_o
    for k in range(2, 6):
        p = p.matmul(a)
        p_o = oracle.matmul(p_o, a_o)
        assert_same(p.to_host(), p_o, f"A^{k} exact={exact}")
    rng = np.random.default_rng(3)
    n = 30000
    lens = [1, 2, 5, 20, 33, 64, 100, 200, 400, 800, 1500, 3000, 6000, 12000, 0, 7]
    r = np.concatenate([np.full(l, i) for i, l in enumerate(lens)])
    c = np.concatenate([rng.choice(n, size=l, replace=False) for l in lens])
    ragged = hostgen.from_coo(n, n, r, c, rng.integers(1, 4, size=r.size, dtype=np.uint64).astype(hostgen.vdtype(bits)), bits)
    check_product(oracle, gpu_ctx, ragged, rand_csr(rng, n, n, 5 * n, bits), f"ragged exact={exact}")
    wide_n = 1_500_000
    lens = [3, 40, 150, 700, 2500, 9000]
    r = np.concatenate([np.full(l, i * 1000) for i, l in enumerate(lens)])
    c = np.concatenate([rng.choice(wide_n, size=l, replace=False) for l in lens])
    wide = hostgen.from_coo(wide_n, wide_n, r, c, np.ones(r.size, hostgen.vdtype(bits)), bits)
    check_product(oracle, gpu_ctx, wide, rand_csr(rng, wide_n, wide_n, 3 * wide_n, bits), f"wide exact={exact}")


def test_exact_mode_saturating_values(gpu_ctx, oracle, cfg):
    cfg(pipeline=2, placement=1)
    rng = np.random.default_rng(9)
    a_h = rand_csr(rng, 500, 500, 6000, 64, 3)
    a_h.values[:] = rng.integers(1 << 61, 1 << 63, size=a_h.nnz(), dtype=np.uint64)
    check_product(oracle, gpu_ctx, a_h, a_h, "u64 saturation, exact mode")


# ------------------------------------------------------------------ einsum front-end (src/graph_csr.rs:1593-1631: einsum == matmul)
def test_einsum_ab_bc_ac_equals_matmul(gpu_ctx, oracle):
    from sparse_linear_algebra_tests_b200 import einsum_sparse_driven, einsum_sparse_hash
    edges = [(0, 1), (0, 2), (1, 2), (1, 3), (2, 3), (3, 4), (4, 5), (5, 0), (2, 5)]
    a = B200Matrix.from_edges(6, edges, 64)
    want = a.matmul(a).to_host()
    assert_same(einsum_sparse_driven("ab,bc->ac", a, a).to_host(), want)
    assert_same(einsum_sparse_hash("ij,jk->ik", a, a).to_host(), want)
    assert_same(want, oracle.matmul(to_o(oracle, a.to_host()), to_o(oracle, a.to_host())))
    with pytest.raises(AssertionError):
        einsum_sparse_driven("ab,cb->ac", a, a)                      # needs a transpose: not the row-wise arrangement
    # Sparse2D view of the result (einsum-dyn/src/sparse.rs:42-55)
    c = a.matmul(a)
    assert c.n_rows() == 6 and sum(c.row_nnz(r) for r in range(6)) == c.nnz()
    col, val = c.row_entry(0, 0)
    assert c.get(0, col) == val


# ------------------------------------------------------------------ circular column windows (torus rows that wrap around the index space)
@pytest.mark.parametrize("circular", ["1", "0"])
@pytest.mark.parametrize("wincap", ["2", "8", "64"])
def test_circular_windows_on_a_long_torus(gpu_ctx, oracle, cfg, circular, wincap):
    """48 x 6 x 6 torus (1728 columns = 14 groups of 128): with a window of a few groups only rows away from the ends of
    the index space fit a plain window; measured from the row's own reference column every row does.  The rows that wrap
    come out of the kernel rotated and must land in ascending column order; both settings must give the reference bytes."""
    cfg(pipeline=2, window_cap_groups=int(wincap), circular_windows=int(circular))
    full = hostgen.lattice([48, 6, 6], True, 64)
    a_h = hostgen.thin(full, 0.2, bytes([7] * 32))
    a, a_o = B200Matrix.from_host(a_h, gpu_ctx), to_o(oracle, a_h)
    p, p_o = a, a_o
    for k in range(2, 6):
        p = p.matmul(a, want_stats=True)
        p_o = oracle.matmul(p_o, a_o)
        assert_same(p.to_host(), p_o, f"A^{k} circular={circular} wincap={wincap}")
    # rectangular row block (a GPU's share in the multi-GPU run) against the replicated square operand
    blk = B200Matrix(gpu_ctx.row_block(a.device, 0, 300))
    got = blk.matmul(a).matmul(a).to_host()
    want = oracle.matmul(oracle.matmul(to_o(oracle, a_h.row_block(0, 300)), a_o), a_o)
    assert_same(got, want, "row block")


def test_circular_window_exact_mode(gpu_ctx, oracle, cfg):
    cfg(pipeline=2, window_cap_groups=4, placement=1)
    full = hostgen.lattice([40, 5, 5], True, 32)
    a_h = hostgen.thin(full, 0.25, bytes([9] * 32))
    a, a_o = B200Matrix.from_host(a_h, gpu_ctx), to_o(oracle, a_h)
    p, p_o = a, a_o
    for k in range(2, 5):
        p = p.matmul(a)
        p_o = oracle.matmul(p_o, a_o)
        assert_same(p.to_host(), p_o, f"A^{k}")


# ------------------------------------------------------------------ operand-level arc windows (row blocks of a long torus: one window for all rows)
@pytest.mark.parametrize("arc", ["1", "0"])
@pytest.mark.parametrize("block", [(0, 500), (1800, 2400), (3600, 4096)])
def test_arc_window_of_a_row_block(gpu_ctx, oracle, cfg, arc, block):
    """A GPU's row block of a 64 x 8 x 8 torus touches a short circular arc of the 4096 columns, and each multiply by A widens
    the arc by A's column span.  With arc_window=1 the engine uses that arc as the one bitmap window of every row (blocks at
    the ends of the index space wrap around it); with 0 it falls back to per-row windows.  Both must give the reference bytes."""
    cfg(pipeline=2, arc_window=int(arc))
    full = hostgen.lattice([64, 8, 8], True, 64)
    a_h = hostgen.thin(full, 0.15, bytes([11] * 32))
    a, a_o = B200Matrix.from_host(a_h, gpu_ctx), to_o(oracle, a_h)
    p = B200Matrix(gpu_ctx.row_block(a.device, block[0], block[1]))
    p_o = to_o(oracle, a_h.row_block(block[0], block[1]))
    for k in range(2, 7):
        p = p.matmul(a, want_stats=True)
        p_o = oracle.matmul(p_o, a_o)
        assert_same(p.to_host(), p_o, f"A^{k} block={block} arc={arc}")


def test_arc_window_exact_mode_and_add(gpu_ctx, oracle, cfg):
    """Arc windows with the exact-memory (count first) path, and a product handle that went through add() keeps a valid arc."""
    cfg(pipeline=2, placement=1)
    full = hostgen.lattice([64, 8, 8], True, 32)
    a_h = hostgen.thin(full, 0.15, bytes([12] * 32))
    a, a_o = B200Matrix.from_host(a_h, gpu_ctx), to_o(oracle, a_h)
    p = B200Matrix(gpu_ctx.row_block(a.device, 3900, 4096))
    p_o = to_o(oracle, a_h.row_block(3900, 4096))
    for k in range(2, 5):
        q, q_o = p.matmul(a), oracle.matmul(p_o, a_o)
        assert_same(q.to_host(), q_o, f"A^{k}")
        p, p_o = q.add(q), oracle.add(q_o, q_o)
        assert_same(p.to_host(), p_o, f"sum {k}")


# ------------------------------------------------------------------ fixture generators on the device
@pytest.mark.parametrize("bits", [32, 64])
@pytest.mark.parametrize("dims,torus", [([5], False), ([5], True), ([1], True), ([3, 3], True), ([2, 2, 2], False), ([2, 2, 2], True),
                                        ([4, 3], False), ([6, 5, 4], True), ([3, 1, 4], True), ([3, 2, 2, 3], True), ([7, 6, 5], False), ([], True)])
def test_device_lattice_equals_host_builder(gpu_ctx, dims, torus, bits):
    """CsrMatrix::lattice on the device (src/graph_csr.rs:177-222) against the host builder that the oracle pins:
    1-D 5 -> 8 entries, torus 10; 3x3 torus 72; 2x2x2 -> 56 (the reference's own counts, src/graph_csr.rs:1011-1093),
    side-1 and side-2 torus dimensions (self loops and summed duplicates), four dimensions, the empty shape."""
    got = B200Matrix.lattice(dims, torus, bits, gpu_ctx).to_host()
    want = hostgen.lattice(dims, torus, bits)
    assert_same(got, want, f"lattice {dims} torus={torus}")
    counts = {((5,), False): 8, ((5,), True): 10, ((3, 3), True): 72, ((2, 2, 2), False): 56}
    if (tuple(dims), torus) in counts:
        assert got.nnz() == counts[(tuple(dims), torus)]


@pytest.mark.parametrize("bits", [32, 64])
def test_device_thin_equals_host_builder_and_shares_a_generator(gpu_ctx, bits):
    """CsrMatrix::thin on the device (src/graph_csr.rs:225-247) with StdRng([42;32]) and other seeds; a second thinning that
    continues the same generator (the sweep of bench_matmul_magnus, src/graph_magnus.rs:800-821) must skip the first one's
    draws; side-2 torus values (2) and self loops ride through."""
    taken = 0
    for dims, dens, seed in (([6, 6, 6], 3.0 / 26.0, bytes([42] * 32)), ([9, 7], 0.4, bytes(range(32))), ([2, 3, 4], 0.5, bytes([7] * 32)),
                             ([12], 0.9, bytes([1] * 32)), ([5, 5, 5], 0.0, bytes([3] * 32)), ([5, 5, 5], 1.0, bytes([3] * 32))):
        full_h = hostgen.lattice(dims, True, bits)
        full = B200Matrix.lattice(dims, True, bits, gpu_ctx)
        got = full.thin(dens, seed)
        assert_same(got.to_host(), hostgen.thin(full_h, dens, seed), f"thin {dims} {dens}")
        assert got.last_draws == hostgen.draws_of_thin(full_h)
    shared = bytes([42] * 32)
    for side, epn in ((5, 2.0), (5, 8.0), (6, 3.0)):                 # three instances from one generator, as the sweep does
        full_h = hostgen.lattice([side] * 3, True, bits)
        got = B200Matrix.lattice([side] * 3, True, bits, gpu_ctx).thin(epn / 26.0, shared, skip=taken)
        assert_same(got.to_host(), hostgen.thin(full_h, epn / 26.0, shared, taken), f"shared generator side={side}")
        taken += got.last_draws


def test_device_built_reference_instance_multiplies_like_the_host_built_one(gpu_ctx, oracle):
    """The 30^3 bench instance built entirely on the device: same matrix (README nnz 81 434) and the same A^2."""
    a = B200Matrix.lattice([30, 30, 30], True, 64, gpu_ctx).thin(3.0 / 26.0, bytes([42] * 32))
    a_h = hostgen.reference_bench_instance(30, 3.0, 64)
    assert_same(a.to_host(), a_h, "device-built bench instance")
    assert a.nnz() == 81434
    assert_same(a.matmul(a).to_host(), oracle.matmul(to_o(oracle, a_h), to_o(oracle, a_h)), "A^2 of the device-built instance")


def test_device_thin_of_a_non_symmetric_matrix(gpu_ctx):
    """thin() looks the mirror up with get(c, r) (src/graph_csr.rs:238): on a matrix that is not symmetric a kept upper entry
    whose mirror is absent stays alone, and a lower entry whose mirror is absent can never be kept.  Random directed graph with
    self loops added, values > 1."""
    rng = np.random.default_rng(5)
    n = 300
    r = rng.integers(0, n, 4000); c = rng.integers(0, n, 4000); v = rng.integers(1, 9, 4000)
    a_h = hostgen.from_coo(n, n, np.concatenate([r, np.arange(0, n, 7)]), np.concatenate([c, np.arange(0, n, 7)]),
                           np.concatenate([v, np.full(len(range(0, n, 7)), 3)]), 64)
    a = B200Matrix.from_host(a_h, gpu_ctx)
    for dens, seed in ((0.5, bytes([42] * 32)), (0.9, bytes([9] * 32))):
        got = a.thin(dens, seed)
        assert_same(got.to_host(), hostgen.thin(a_h, dens, seed), f"non-symmetric thin {dens}")
        assert got.last_draws == hostgen.draws_of_thin(a_h)
