"""Parity of the fused pipeline (csrc/fused.cu: pre-pass + one persistent numeric/placement kernel) and the parity scope
the north star names: A^2..A^7 on the 30^3 torus bit for bit, the whole side x e_per_n sweep of bench_matmul_magnus, a
100^3 torus, R-MAT graphs at both skews.  Everything goes through the C ABI; the oracle is the checker only.
"""
import numpy as np
import pytest

from sparse_linear_algebra_tests_b200 import B200Error, B200Matrix, hostgen

pytestmark = pytest.mark.gpu
DEFAULT_PIPELINE = 6     # what `pipeline = 0` picks for A^7 of the headline chain (the operands commute: evaluated as A x A^6, leftmul.cu)


def to_o(O, h):
    return O.Csr(h.rows, h.cols, h.row_ptr, h.col_idx, h.values)


def assert_same(got, want, what=""):
    assert got.rows == want.rows and got.cols == want.cols, what
    assert np.array_equal(got.row_ptr, want.row_ptr), f"row_ptr differs {what}"
    assert np.array_equal(got.col_idx, want.col_idx), f"col_idx differs {what}"
    assert got.values.dtype == want.values.dtype, what
    assert np.array_equal(got.values, want.values), f"values differ {what}"


def chain_check(O, ctx, a_h, powers, left_h=None, what=""):
    """A^k = A^(k-1) x A against oracle.matmul_par for k = 2..powers (the protocol of bench_repeated_exponentiation,
    src/graph_magnus.rs:740-786); the GPU chain is queued without ever asking for a size in between."""
    a = B200Matrix.from_host(a_h, ctx)
    a_o = to_o(O, a_h)
    p = a if left_h is None else B200Matrix.from_host(left_h, ctx)
    p_o = a_o if left_h is None else to_o(O, left_h)
    gpu = []
    for _k in range(2, powers + 1):
        p = p.matmul(a)
        gpu.append(p)
    for k, g in zip(range(2, powers + 1), gpu):
        p_o = O.matmul_par(p_o, a_o)
        assert_same(g.to_host(), p_o, f"A^{k} {what}")
    return gpu


# ------------------------------------------------------------------ the headline instance, every power, both value widths
@pytest.mark.parametrize("pipeline", [0, 1, 2, 3, 4, 5, 6], ids=["default", "fused", "binned", "rowwarp", "onelaunch", "onepass", "leftmul"])
@pytest.mark.parametrize("bits", [64, 32])
def test_reference_instance_30_every_power_bit_exact(gpu_ctx, oracle, cfg, bits, pipeline):
    """BASELINE configs[1]: the reference's exact operand (StdRng([42;32]) thinning of the 30^3 Moore torus, 81 434 nnz),
    A^2..A^7, row_ptr / col_idx / values byte for byte, through both pipelines; nnz per power = the README column
    (README.md:42-47)."""
    cfg(pipeline=pipeline)
    a_h = hostgen.reference_bench_instance(30, 3.0, bits)
    assert a_h.nnz() == 81434
    gpu = chain_check(oracle, gpu_ctx, a_h, 7, what=f"u{bits} pipeline {pipeline}")
    assert [g.nnz() for g in gpu] == [251590, 655391, 1574848, 3383207, 6590100, 11736555]
    st = gpu[-1].device.product_stats()
    assert st.nnz_c == 11736555 and st.pipeline == {0: DEFAULT_PIPELINE, 1: 1, 2: 2, 3: 3, 4: 4, 5: 5, 6: 6}[pipeline]
    if pipeline == 1:
        assert st.sym_bin_rows[2] == 0, "a row of the headline multiply left the fused kernel for the counted lists"


@pytest.mark.parametrize("fields", [
    dict(pipeline=1),                                          # fused, engine's own choices
    dict(pipeline=1, fused_threads=128),
    dict(pipeline=1, fused_threads=64),
    dict(pipeline=1, pack_b=0),                                # plain descriptors + column gathers instead of packed records
    dict(pipeline=1, pack_b=0, lanes_per_entry_lg=2),
    dict(pipeline=1, force_acc_mode=1),                        # 64-bit sums as two 32-bit words
    dict(pipeline=1, force_acc_mode=2),                        # saturating CAS accumulators
    dict(pipeline=1, arc_window=0),                            # whole column space as the window
    dict(pipeline=1, fused_window_cols=2048),                  # window too short for the later powers: rows fall to the counted lists
    dict(pipeline=1, fused_window_cols=2048, heavy_kernel=0),
    dict(pipeline=1, fused_dense_pmax=40),                     # "heavy" rows from 41 products on
    dict(pipeline=2),                                          # binned pipeline, for the same bytes
], ids=lambda f: ",".join(f"{k}={v}" for k, v in f.items()))
@pytest.mark.parametrize("bits", [64, 32])
def test_fused_variants_on_a_torus_chain(gpu_ctx, oracle, cfg, fields, bits):
    if fields.get("force_acc_mode") == 2 and bits == 32:
        pytest.skip("u32 values have no saturating-CAS mode (sums are kept in 64 bits)")
    cfg(**fields)
    a_h = hostgen.reference_bench_instance(12, 3.0, bits)
    chain_check(oracle, gpu_ctx, a_h, 6, what=str(fields))


@pytest.mark.parametrize("fields", [dict(pipeline=1), dict(pipeline=1, fused_window_cols=512), dict(pipeline=1, arc_window=0)],
                         ids=lambda f: ",".join(f"{k}={v}" for k, v in f.items()))
def test_fused_windows_that_wrap_around_the_index_space(gpu_ctx, oracle, cfg, fields):
    """48 x 6 x 6 torus: rows near both ends of the index space have windows that wrap (origin > 0, emit rotated);
    plus a rectangular row block (a GPU's share of the multi-GPU run) against the replicated square operand."""
    cfg(**fields)
    full = hostgen.lattice([48, 6, 6], True, 64)
    a_h = hostgen.thin(full, 0.2, bytes([7] * 32))
    chain_check(oracle, gpu_ctx, a_h, 6, what="48x6x6")
    for r0, r1 in ((0, 300), (1500, 1728), (700, 1100)):
        chain_check(oracle, gpu_ctx, a_h, 5, left_h=a_h.row_block(r0, r1), what=f"block {r0}:{r1}")


def test_fused_rows_longer_than_ring_and_product_buffer(gpu_ctx, oracle, cfg):
    """Tiny ring and product buffer: rows that spill to the per-CTA global slots (longer than the ring) and rows that
    generate their products twice (more products than the buffer), interleaved with rows that fit."""
    a_h = hostgen.reference_bench_instance(12, 4.0, 64)
    for fields in (dict(fused_ring_slots=256, fused_product_slots=64), dict(fused_ring_slots=256, fused_product_slots=4096), dict(fused_ring_slots=8192, fused_product_slots=64)):
        cfg(pipeline=1, **fields)
        chain_check(oracle, gpu_ctx, a_h, 6, what=str(fields))


def test_fused_mixed_classes_and_empty_rows(gpu_ctx, oracle, cfg):
    """Tiny, dense, heavy and empty rows interleaved in one left operand (unit boundaries at every class change), random
    columns (no arc structure: the window is the whole column space)."""
    cfg(pipeline=1)
    rng = np.random.default_rng(17)
    n = 20000
    lens = rng.choice([0, 0, 1, 2, 3, 5, 9, 40, 120, 700, 3000], size=1500)
    lens[7] = 19000                                         # ~76 k products: beyond the dense class, a heavy row
    rows = rng.choice(n, size=lens.size, replace=False)
    r = np.concatenate([np.full(l, i) for i, l in zip(rows, lens)])
    c = np.concatenate([rng.choice(n, size=l, replace=False) for l in lens])
    for bits in (64, 32):
        a_h = hostgen.from_coo(n, n, r, c, rng.integers(1, 4, size=r.size).astype(hostgen.vdtype(bits)), bits)
        rb = rng.integers(0, n, size=4 * n); cb = rng.integers(0, n, size=4 * n)
        b_h = hostgen.from_coo(n, n, rb, cb, rng.integers(1, 4, size=4 * n).astype(hostgen.vdtype(bits)), bits)
        a, b = B200Matrix.from_host(a_h, gpu_ctx), B200Matrix.from_host(b_h, gpu_ctx)
        c_g = a.matmul(b, want_stats=True)
        assert_same(c_g.to_host(), oracle.matmul(to_o(oracle, a_h), to_o(oracle, b_h)), f"mixed u{bits}")
        st = c_g.last_stats
        assert st.pipeline == 1 and st.sym_bin_rows[0] > 0 and st.sym_bin_rows[1] > 0 and st.sym_bin_rows[2] > 0 and st.sym_bin_rows[3] > 0


def test_fused_saturation(gpu_ctx, oracle, cfg):
    cfg(pipeline=1)
    rng = np.random.default_rng(9)
    n = 800
    r = rng.integers(0, n, size=12000); c = rng.integers(0, n, size=12000)
    a_h = hostgen.from_coo(n, n, r, c, rng.integers(1 << 61, 1 << 63, size=12000, dtype=np.uint64), 64)
    a = B200Matrix.from_host(a_h, gpu_ctx)
    assert_same(a.matmul(a).to_host(), oracle.matmul(to_o(oracle, a_h), to_o(oracle, a_h)), "u64 saturating")
    a32 = hostgen.from_coo(n, n, r, c, rng.integers(1 << 29, 1 << 31, size=12000).astype(np.uint32), 32)
    b32 = B200Matrix.from_host(a32, gpu_ctx)
    assert_same(b32.matmul(b32).to_host(), oracle.matmul(to_o(oracle, a32), to_o(oracle, a32)), "u32 saturating")
    # the reference's own saturating case: (A + I) of a 64-node chain squared until stable (src/graph_csr.rs:931-939)
    chain = B200Matrix.from_edges(64, [(i, i + 1) for i in range(63)], 32, gpu_ctx)
    m = chain.add(B200Matrix.identity(64, 32, gpu_ctx))
    m_o = to_o(oracle, m.to_host())
    for _ in range(8):
        m, m_o = m.matmul(m), oracle.matmul(m_o, m_o)
        assert_same(m.to_host(), m_o, "64-chain squaring")


def test_async_products_and_stats(gpu_ctx, oracle, cfg):
    """b200_spgemm without stats returns before the product's size is known: more than B200_REPORT_SLOTS products may be
    in flight, handles may be freed unread, and product_stats / nnz / download wait for the right report."""
    cfg(pipeline=1)
    a_h = hostgen.reference_bench_instance(10, 3.0, 64)
    a = B200Matrix.from_host(a_h, gpu_ctx)
    a_o = to_o(oracle, a_h)
    want2 = oracle.matmul(a_o, a_o)
    prods = [a.matmul(a) for _ in range(40)]                 # 40 unread products: the report ring wraps twice
    del prods[5:20]                                          # freed without ever being read
    for p in prods:
        assert_same(p.to_host(), want2, "A^2 in flight")
    st = prods[-1].device.product_stats()
    assert st.nnz_c == want2.nnz() and st.products == int(oracle.row_products(a_o, a_o).sum()) and st.ms_total > 0
    with pytest.raises(B200Error):
        a.device.product_stats()                             # an uploaded matrix is not a product


# ------------------------------------------------------------------ BASELINE configs[2]: every cell of the reference sweep
def test_every_cell_of_the_reference_sweep(gpu_ctx, oracle):
    """bench_matmul_magnus (src/graph_magnus.rs:790-929): side in {5,10,20,30} x e_per_n in {2,3,4,8,26}, A x A, ONE
    StdRng([42;32]) shared across the grid exactly like the reference loop (:800, :817-821)."""
    seed, used = bytes([42] * 32), 0
    for side in (5, 10, 20, 30):
        full = hostgen.lattice([side] * 3, True, 64)
        full_epn = full.nnz() / full.rows
        for epn in (2.0, 3.0, 4.0, 8.0, 26.0):
            density = epn / full_epn
            if density >= 1.0:
                a_h = full
            else:
                a_h = hostgen.thin(full, density, seed, skip=used)
                used += hostgen.draws_of_thin(full)
            a = B200Matrix.from_host(a_h, gpu_ctx)
            a_o = to_o(oracle, a_h)
            assert_same(a.matmul(a).to_host(), oracle.matmul_par(a_o, a_o), f"side {side} e/n {epn}")


# ------------------------------------------------------------------ larger inputs under the driver's eye
def test_torus_100_low_powers(gpu_ctx, oracle):
    """100^3 Moore torus (10^6 nodes, ~3 e/n), A^2..A^4 (91 M products at A^4): per-row windows far wider than shared
    memory, so this runs the tiny rows inside the fused kernel and everything else through the counted lists."""
    a_h = hostgen.reference_bench_instance(100, 3.0, 64)
    chain_check(oracle, gpu_ctx, a_h, 4, what="100^3")


@pytest.mark.parametrize("scale,abc,bits", [(16, (0.57, 0.19, 0.19), 64), (17, (0.45, 0.15, 0.15), 64), (18, (0.45, 0.15, 0.15), 32), (14, (0.57, 0.19, 0.19), 32)])
def test_rmat_square(gpu_ctx, oracle, scale, abc, bits):
    """R-MAT A^2 (SURVEY.md App. C generator): the Graph500 skew reaches rows of > 10^5 products (heavy kernel, 64-bit
    sums), the milder skew is the hash lists' workload."""
    a_h = hostgen.rmat(scale, 16, abc[0], abc[1], abc[2], 42, bits)
    a = B200Matrix.from_host(a_h, gpu_ctx)
    a_o = to_o(oracle, a_h)
    c = a.matmul(a, want_stats=True)
    assert_same(c.to_host(), oracle.matmul_par(a_o, a_o), f"rmat scale {scale} {abc}")
    if abc[0] > 0.5:
        assert c.last_stats.sym_bin_rows[9] > 0, "no heavy row in a Graph500-skew R-MAT?"


# ------------------------------------------------------------------ format checks and sort-key edge
def test_tiny_row_at_the_top_of_a_2_pow_27_column_space(gpu_ctx, oracle, cfg):
    """cols == 2^27 exactly: a 32-product tiny row whose last product is column cols-1 held by lane 31 would pack to the
    empty-key word if the narrow sort key were used at this width."""
    n = 1 << 27
    for pipeline in (1, 2):
        cfg(pipeline=pipeline)
        a_h = hostgen.from_coo(4, 32, [1] * 32, list(range(32)), np.ones(32, np.uint64), 64)
        b_cols = [n - 1 - 3 * (31 - i) for i in range(32)]       # row k of B holds one column; row 31 holds cols - 1
        b_h = hostgen.from_coo(32, n, list(range(32)), b_cols, np.arange(1, 33, dtype=np.uint64), 64)
        a, b = B200Matrix.from_host(a_h, gpu_ctx), B200Matrix.from_host(b_h, gpu_ctx)
        assert_same(a.matmul(b).to_host(), oracle.matmul(to_o(oracle, a_h), to_o(oracle, b_h)), f"2^27 columns, pipeline {pipeline}")


def test_upload_rejects_unsorted_or_repeated_columns(gpu_ctx):
    rp = np.array([0, 3, 5], np.uint64)
    ok = gpu_ctx.upload(2, 8, rp, np.array([1, 4, 7, 0, 2], np.uint32), np.ones(5, np.uint64))
    assert ok.nnz == 5
    for cols in ([4, 1, 7, 0, 2], [1, 1, 7, 0, 2], [1, 4, 7, 2, 2]):
        with pytest.raises(B200Error) as e:
            gpu_ctx.upload(2, 8, rp, np.array(cols, np.uint32), np.ones(5, np.uint64))
        assert e.value.code == 5
    with pytest.raises(B200Error):
        gpu_ctx.upload(2, 8, rp, np.array([1, 4, 7], np.uint32), np.ones(3, np.uint64))        # lengths do not match row_ptr
    with pytest.raises(B200Error):
        gpu_ctx.upload(3, 8, rp, np.array([1, 4, 7, 0, 2], np.uint32), np.ones(5, np.uint64))  # row_ptr too short


# ------------------------------------------------------------------ the one-pass multiply (dense.cu, pipeline 5)
@pytest.mark.parametrize("ctas", [3, 8], ids=["window-18k", "window-6k-pieces"])
@pytest.mark.parametrize("bits", [64, 32])
def test_one_pass_dense_windows(gpu_ctx, oracle, cfg, bits, ctas):
    """Pipeline 5 on tori whose rows wrap around the index space (arc origin > 0, rows written rotated), with the window
    shrunk (8 CTAs per SM: ~6 K columns) so that the later powers of the 30^3 chain are produced in column pieces."""
    cfg(pipeline=5, fused_threads=ctas)
    a_h = hostgen.reference_bench_instance(30, 3.0, bits)
    gpu = chain_check(oracle, gpu_ctx, a_h, 6 if ctas == 8 else 7, what=f"one-pass u{bits} ctas {ctas}")
    assert gpu[-1].device.product_stats().pipeline == 5
    long_h = hostgen.thinned_torus([64, 6, 6], 4.0 / 26.0, bytes([7] * 32), bits)
    chain_check(oracle, gpu_ctx, long_h, 5, what="one-pass long torus")


def test_one_pass_row_blocks_empty_rows_and_values(gpu_ctx, oracle, cfg):
    """Row blocks at both ends of the index space (rectangular left operand), rows without products, values > 1 on both sides."""
    cfg(pipeline=5)
    rng = np.random.default_rng(11)
    a_h = hostgen.thinned_torus([40, 10, 10], 3.5 / 26.0, bytes([3] * 32), 64)      # (mean row <= 4: sector-packed right operand)
    a_h.values[:] = rng.integers(1, 30, a_h.values.size)
    a_o = to_o(oracle, a_h)
    full = oracle.matmul_par(oracle.matmul_par(a_o, a_o), a_o)
    a = B200Matrix.from_host(a_h, gpu_ctx)
    for r0, r1 in ((0, 700), (1500, 2600), (3300, 4000)):
        blk = B200Matrix.from_host(a_h.row_block(r0, r1), gpu_ctx)
        first = blk.matmul(a)
        assert first.device.product_stats().pipeline == 5
        got = first.matmul(a)                                                # (64-bit sums from here on: the engine picks another pipeline by itself)
        want = hostgen.HostCsr(full.rows, full.cols, full.row_ptr, full.col_idx, full.values).row_block(r0, r1)
        assert_same(got.to_host(), want, f"rows {r0}..{r1}")


def test_narrow_download_widens_to_the_same_bytes(gpu_ctx, oracle, cfg):
    """b200_config.narrow_download: u64 values proven < 2^32 cross PCIe as u32 and host threads widen them -- the arrays the
    caller gets are the same (asynchronous variant into pinned memory, several products in flight, then one synchronize)."""
    import torch
    a_h = hostgen.reference_bench_instance(30, 3.0, 64)
    a = gpu_ctx.upload(a_h.rows, a_h.cols, a_h.row_ptr, a_h.col_idx, a_h.values)
    a_o = to_o(oracle, a_h)
    want, p_o = [], a_o
    for _ in range(4):
        p_o = oracle.matmul_par(p_o, a_o); want.append(p_o)
    for on in (1, 0):
        cfg(narrow_download=on)
        p, got, bufs = a, [], []
        for w in want:
            p = gpu_ctx.spgemm(p, a)
            b = (torch.empty((a_h.rows + 1) * 8, dtype=torch.uint8).pin_memory(), torch.empty(w.nnz() * 4, dtype=torch.uint8).pin_memory(),
                 torch.empty(w.nnz() * 8, dtype=torch.uint8).pin_memory())
            p.download_async_into(b[0].data_ptr(), b[1].data_ptr(), b[2].data_ptr())
            got.append(p); bufs.append(b)
        gpu_ctx.synchronize()
        for w, b in zip(want, bufs):
            assert np.array_equal(b[0].numpy().view(np.uint64), w.row_ptr) and np.array_equal(b[1].numpy().view(np.uint32), w.col_idx)
            assert np.array_equal(b[2].numpy().view(np.uint64), w.values), f"narrow_download={on}"
