"""Parity of the chunked heavy-row kernels (csrc/heavy.cu): column-space chunks, TMA-staged B-row segments, a dense
accumulator per chunk, rows cut into several work units.  The reference reaches the corresponding MAGNUS levels
through MagnusMatrix::matmul (src/graph_magnus.rs:225-232); the checker is the CSR restatement (oracle.matmul_par,
src/graph_csr.rs:350-484).  Everything goes through the C ABI.
"""
import numpy as np
import pytest

from sparse_linear_algebra_tests_b200 import B200Matrix, hostgen

pytestmark = pytest.mark.gpu


def to_o(O, h):
    return O.Csr(h.rows, h.cols, h.row_ptr, h.col_idx, h.values)


def assert_same(got, want, what=""):
    assert np.array_equal(got.row_ptr, want.row_ptr), f"row_ptr differs {what}"
    assert np.array_equal(got.col_idx, want.col_idx), f"col_idx differs {what}"
    assert got.values.dtype == want.values.dtype, what
    assert np.array_equal(got.values, want.values), f"values differ {what}"


def hub_graph(n, hubs, hub_fill, deg, links, bits, vmax, seed, frac=0.25):
    """`hubs` dense rows (a fraction hub_fill of all columns each), every other row `deg` random columns plus `links`
    columns among the hubs: rows that link to hubs are heavy (their B rows are the hubs' long rows -- the segments the
    chunked kernel stages through shared memory)."""
    rng = np.random.default_rng(seed)
    r, c = [], []
    for h in range(hubs):
        cols = np.nonzero(rng.random(n) < hub_fill)[0]
        r.append(np.full(cols.size, h)); c.append(cols)
    rest = np.arange(hubs, n)
    r.append(np.repeat(rest, deg)); c.append(rng.integers(0, n, rest.size * deg))
    linked = rest[rng.random(rest.size) < frac]
    r.append(np.repeat(linked, links)); c.append(rng.integers(0, hubs, linked.size * links))
    r = np.concatenate(r); c = np.concatenate(c)
    v = rng.integers(1, vmax + 1, r.size, dtype=np.uint64).astype(hostgen.vdtype(bits))
    return hostgen.from_coo(n, n, r, c, v, bits, saturating=True)


@pytest.mark.parametrize("placement", [-1, 0, 1], ids=["auto", "scratch", "exact"])
@pytest.mark.parametrize("bits,vmax", [(64, 1), (64, 1 << 20), (32, 3000), (64, (1 << 64) - 1)], ids=["u64-pattern", "u64-wide", "u32", "u64-saturating"])
def test_hub_rows_through_small_chunks(gpu_ctx, oracle, cfg, bits, vmax, placement):
    """1024-column chunks, rows cut into work units of ~4096 products: many chunks per row, several CTAs per row, long
    segments (hub rows hold ~300 entries per chunk) through the bulk-copy stages, short ones through the enumeration; all
    three accumulator modes and both placements."""
    a_h = hub_graph(6000, 12, 0.3, 6, 6, bits, vmax, 7)
    cfg(pipeline=2, placement=placement, heavy_chunk_cols=1024, heavy_min_products=8193, heavy_unit_products=4096)
    a = B200Matrix.from_host(a_h, gpu_ctx)
    c = a.matmul(a, want_stats=True)
    assert c.last_stats.sym_bin_rows[9] > 0, "no heavy row"
    assert_same(c.to_host(), oracle.matmul_par(to_o(oracle, a_h), to_o(oracle, a_h)), f"hub graph u{bits} vmax {vmax} placement {placement}")


def test_chunked_kernel_on_and_off_agree(gpu_ctx, oracle, cfg):
    """The same multiply with the chunked kernels switched off (global-memory table for every heavy row)."""
    a_h = hub_graph(70000, 6, 0.06, 5, 3, 64, 50, 11, frac=0.03)
    want = oracle.matmul_par(to_o(oracle, a_h), to_o(oracle, a_h))
    for on in (1, 0):
        cfg(pipeline=2, heavy_kernel=on)
        a = B200Matrix.from_host(a_h, gpu_ctx)
        assert_same(a.matmul(a).to_host(), want, f"heavy_kernel={on}")


@pytest.mark.parametrize("scale,bits", [(16, 64), (17, 32)])
def test_graph500_skew_default_plan(gpu_ctx, oracle, scale, bits):
    """Graph500-skew R-MAT with the automatic plan (column chunks sized from shared memory): hub rows of > 10^5 products."""
    a_h = hostgen.rmat(scale, 16, 0.57, 0.19, 0.19, 42, bits)
    a = B200Matrix.from_host(a_h, gpu_ctx)
    a_o = to_o(oracle, a_h)
    c = a.matmul(a, want_stats=True)
    assert c.last_stats.sym_bin_rows[9] > 0
    assert_same(c.to_host(), oracle.matmul_par(a_o, a_o), f"graph500 scale {scale} u{bits}")


def test_rectangular_right_operand_and_empty_chunks(gpu_ctx, oracle, cfg):
    """Columns only in the first and the last chunk of a wide rectangular B: empty chunks in between, a last chunk that is
    not full."""
    rng = np.random.default_rng(3)
    m, k, n = 40, 3000, 100_000 + 37
    ar = np.repeat(np.arange(m), 600); ac = rng.integers(0, k, ar.size)
    a_h = hostgen.from_coo(m, k, ar, ac, rng.integers(1, 9, ar.size, dtype=np.uint64), 64, saturating=True)
    br = np.repeat(np.arange(k), 40)
    bc = np.where(rng.random(br.size) < 0.5, rng.integers(0, 900, br.size), n - 1 - rng.integers(0, 900, br.size))
    b_h = hostgen.from_coo(k, n, br, bc, rng.integers(1, 9, br.size, dtype=np.uint64), 64, saturating=True)
    cfg(pipeline=2, heavy_chunk_cols=2048, heavy_min_products=8193)
    a, b = B200Matrix.from_host(a_h, gpu_ctx), B200Matrix.from_host(b_h, gpu_ctx)
    c = a.matmul(b, want_stats=True)
    assert c.last_stats.sym_bin_rows[9] > 0
    assert_same(c.to_host(), oracle.matmul_par(to_o(oracle, a_h), to_o(oracle, b_h)), "rectangular")


# ------------------------------------------------------------------ COO -> CSR and the R-MAT generator on the device
@pytest.mark.parametrize("saturating", [False, True])
@pytest.mark.parametrize("bits", [32, 64])
@pytest.mark.parametrize("rows,cols,n", [(1, 1, 5), (7, 3, 40), (300, 70000, 20000), (100000, 100000, 700000), (5, 5, 0)])
def test_from_coo_on_device(gpu_ctx, bits, saturating, rows, cols, n):
    """b200_csr_from_coo against the host restatement of CsrMatrix::from_coo (src/graph_csr.rs:83-129; saturating flavour
    linalg/src/csr.rs:158-195): duplicates (a third of the triplets repeat), sums that wrap to zero and explicit zeros are
    dropped, trailing empty rows closed."""
    rng = np.random.default_rng(rows * 31 + n)
    r = rng.integers(0, rows, n); c = rng.integers(0, cols, n)
    if n >= 9:
        r[n // 3: 2 * (n // 3)] = r[: n // 3]; c[n // 3: 2 * (n // 3)] = c[: n // 3]       # duplicates
    top = (1 << bits) - 1
    v = rng.integers(0, 4, n).astype(np.uint64)
    big = rng.random(n) < 0.2
    v[big] = np.uint64(top) - rng.integers(0, 3, int(big.sum())).astype(np.uint64)          # near the top: wraps / saturates
    if n >= 9:
        half = np.uint64(1 << (bits - 1))
        v[0] = half; v[n // 3] = half                                                        # wraps to exactly zero (plain +=)
    want = hostgen.from_coo(rows, cols, r, c, v.astype(hostgen.vdtype(bits)), bits, saturating=saturating)
    got = gpu_ctx.from_coo(rows, cols, r, c, v, bits, saturating)
    rp, ci, vv = got.download()
    assert got.rows == rows and got.cols == cols
    assert np.array_equal(rp, want.row_ptr) and np.array_equal(ci, want.col_idx) and np.array_equal(vv, want.values)


def test_from_coo_rejects_out_of_range(gpu_ctx):
    from sparse_linear_algebra_tests_b200 import B200Error
    with pytest.raises(B200Error):
        gpu_ctx.from_coo(4, 4, [0, 4], [1, 1], [1, 1], 64)
    with pytest.raises(B200Error):
        gpu_ctx.from_coo(4, 4, [0, 1], [1, 9], [1, 1], 64)


@pytest.mark.parametrize("scale,abc,bits", [(10, (0.45, 0.15, 0.15), 64), (16, (0.57, 0.19, 0.19), 32), (18, (0.45, 0.15, 0.15), 64)])
def test_device_rmat_equals_host_generator(gpu_ctx, scale, abc, bits):
    want = hostgen.rmat(scale, 16, abc[0], abc[1], abc[2], 42, bits)
    got = gpu_ctx.rmat(scale, 16, abc[0], abc[1], abc[2], 42, bits)
    rp, ci, vv = got.download()
    assert np.array_equal(rp, want.row_ptr) and np.array_equal(ci, want.col_idx) and np.array_equal(vv, want.values)


def test_graph_type_from_coo_uses_the_device(gpu_ctx, oracle):
    """B200Matrix.from_coo (the MagnusMatrix::from_coo surface): duplicate-edge sum = 2 + 3 = 5 (linalg/src/csr.rs:839-849)."""
    m = B200Matrix.from_coo(3, [(0, 1, 2), (0, 1, 3), (2, 0, 7)], 64, gpu_ctx)
    assert m.get(0, 1) == 5 and m.get(2, 0) == 7 and m.nnz() == 2


# ------------------------------------------------------------------ locality pre-pass: rcm / permute / unpermute / bandwidth_stats
@pytest.mark.parametrize("bits", [32, 64])
def test_permute_and_bandwidth_match_the_oracle(gpu_ctx, oracle, bits):
    """b200_csr_permute / _bandwidth_stats / _rcm_order against the restatement of src/graph_csr.rs:663-818 on the thinned
    10^3 torus (isolated nodes, many components), a random permutation, and the RCM order itself."""
    a_h = hostgen.reference_bench_instance(10, 3.0, bits)
    a_o = to_o(oracle, a_h)
    a = B200Matrix.from_host(a_h, gpu_ctx)
    assert a.bandwidth_stats() == oracle.bandwidth_stats(a_o)
    rng = np.random.default_rng(1)
    perm = rng.permutation(a_h.rows).astype(np.uint32)
    got = gpu_ctx.permute(a.device, perm)
    want = oracle.permute(a_o, perm)
    rp, ci, vv = got.download()
    assert np.array_equal(rp, want.row_ptr) and np.array_equal(ci, want.col_idx) and np.array_equal(vv, want.values)
    order = gpu_ctx.rcm_order(a.device)
    assert np.array_equal(order, oracle.rcm_order(a_o))
    a.rcm()
    assert a.perm is not None and np.array_equal(a.perm, order)
    assert_same(a.to_host(), oracle.permute(a_o, order), "rcm")
    assert a.bandwidth_stats() == oracle.bandwidth_stats(oracle.permute(a_o, order))
    a.unpermute()
    assert a.perm is None
    assert_same(a.to_host(), a_h, "unpermute")


def test_rcm_roundtrips_of_the_reference(gpu_ctx, oracle):
    """test_rcm_unpermute_roundtrip / _lattice / _directed (src/graph_csr.rs:1107-1146) on the device handles; the product
    of a permuted matrix has the nnz of the original's (analyze_graph_structure, :1549)."""
    cases = [B200Matrix.from_edges_undirected(6, [(0, 3), (1, 4), (2, 5), (0, 1), (3, 4)], 64, gpu_ctx),
             B200Matrix.lattice([4, 4], False, 64, gpu_ctx),
             B200Matrix.from_edges(5, [(0, 1), (1, 2), (2, 3), (3, 4), (4, 0), (0, 3)], 64, gpu_ctx)]
    for m in cases:
        orig = m.to_host()
        nnz2 = m.matmul(m).nnz()
        m.rcm()
        assert m.perm is not None
        assert m.matmul(m).nnz() == nnz2
        m.unpermute()
        assert m.perm is None
        assert_same(m.to_host(), orig, "rcm + unpermute")


def test_permute_rejects_a_non_permutation(gpu_ctx):
    from sparse_linear_algebra_tests_b200 import B200Error
    m = B200Matrix.identity(4, 64, gpu_ctx)
    for bad in ([0, 1, 1, 2], [0, 1, 2, 7]):
        with pytest.raises(B200Error):
            gpu_ctx.permute(m.device, np.array(bad, np.uint32))


def test_device_cut_points_follow_the_host_rule(gpu_ctx):
    """b200_shard_rows_by_products (tile sums on the device, one tile per cut read back) gives the cuts of
    distributed.product_balanced_cuts on the same per-row counts -- with heavy rows (weighted 2.5 x) in the operand."""
    from sparse_linear_algebra_tests_b200.distributed import product_balanced_cuts
    a = gpu_ctx.rmat(14, 16, 0.57, 0.19, 0.19, 42, 64)
    prods = gpu_ctx.row_products(a, a)
    assert int(prods.max()) > 8192
    for parts in (1, 2, 3, 8):
        assert np.array_equal(gpu_ctx.shard_rows_by_products(a, a, parts), product_balanced_cuts(prods, parts))
