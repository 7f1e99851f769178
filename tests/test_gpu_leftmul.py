"""Parity of the left multiply (csrc/leftmul.cu, pipeline 6: short rows in A, long sorted rows in B streamed as lists) and of
the commuting-operand swap that routes the steps of a power chain to it (b200_spgemm, `commute_swap`).  Everything goes
through the C ABI; the oracle (CsrMatrix::matmul_par restated, src/graph_csr.rs:350-484) is the checker only.
"""
import numpy as np
import pytest

from sparse_linear_algebra_tests_b200 import B200Matrix, hostgen

pytestmark = pytest.mark.gpu


def to_o(O, h):
    return O.Csr(h.rows, h.cols, h.row_ptr, h.col_idx, h.values)


def assert_same(got, want, what=""):
    assert got.rows == want.rows and got.cols == want.cols, what
    assert np.array_equal(got.row_ptr, want.row_ptr), f"row_ptr differs {what}"
    assert np.array_equal(got.col_idx, want.col_idx), f"col_idx differs {what}"
    assert got.values.dtype == want.values.dtype, what
    assert np.array_equal(got.values, want.values), f"values differ {what}"


def random_csr(rng, rows, cols, per_row, vmax, bits, empty_every=0):
    r, c, v = [], [], []
    for i in range(rows):
        if empty_every and i % empty_every == 0:
            continue
        k = int(rng.integers(1, per_row + 1))
        cc = rng.choice(cols, size=min(k, cols), replace=False)
        r += [i] * len(cc); c += list(cc); v += list(rng.integers(1, vmax + 1, size=len(cc), dtype=np.uint64))
    return hostgen.from_coo(rows, cols, np.array(r, np.uint32), np.array(c, np.uint32), np.array(v, np.uint64), bits)


def left_check(O, ctx, a_h, b_h, what):
    a, b = B200Matrix.from_host(a_h, ctx), B200Matrix.from_host(b_h, ctx)
    c = a.matmul(b, want_stats=True)
    assert c.last_stats.pipeline == 6, what
    want = O.matmul_par(to_o(O, a_h), to_o(O, b_h))
    assert_same(c.to_host(), want, what)
    st = c.last_stats
    assert st.nnz_c == want.nnz() and st.products == int(O.row_products(to_o(O, a_h), to_o(O, b_h)).sum()), what
    return c


@pytest.mark.parametrize("bits", [64, 32])
@pytest.mark.parametrize("fields", [
    dict(pipeline=6),
    dict(pipeline=6, force_acc_mode=1),                       # 64-bit sums as two 32-bit words
    dict(pipeline=6, force_acc_mode=2),                       # saturating CAS accumulators
    dict(pipeline=6, rw_cap_percent=10),                      # rows longer than the accumulator: several passes over their lists
    dict(pipeline=6, circular_windows=0),                     # whole column space as the window instead of the travelling one
], ids=lambda f: ",".join(f"{k}={v}" for k, v in f.items()))
def test_left_multiply_on_a_torus_power(gpu_ctx, oracle, cfg, fields, bits):
    """A x A^4 on a thinned 14^3 torus: lists of ~100 entries, rows at both ends of the index space wrap around it."""
    if fields.get("force_acc_mode") == 2 and bits == 32:
        pytest.skip("u32 values have no saturating-CAS mode (sums are kept in 64 bits)")
    a_h = hostgen.reference_bench_instance(14, 3.0, bits)
    a_o = to_o(oracle, a_h)
    a = B200Matrix.from_host(a_h, gpu_ctx)
    cfg(pipeline=2, commute_swap=0)                           # A^4 by the binned kernels, in the reference's order
    p, p_o = a, a_o
    for _ in range(3):
        p = p.matmul(a)
        p_o = oracle.matmul_par(p_o, a_o)
    cfg(commute_swap=0, **fields)
    c = a.matmul(p, want_stats=True)                          # the offsets bound of A^4 came with the handle: travelling window
    assert c.last_stats.pipeline == 6
    assert_same(c.to_host(), oracle.matmul_par(a_o, p_o), str(fields))
    # the same right operand as a fresh upload (nothing known about its offsets: whole column space as the window)
    p_h = hostgen.HostCsr(p_o.rows, p_o.cols, p_o.row_ptr, p_o.col_idx, p_o.values)
    left_check(oracle, gpu_ctx, a_h, p_h, str(fields))


@pytest.mark.parametrize("bits,vmax", [(64, 1), (64, 3_000_000_000), (32, 70_000), (64, (1 << 63) + 5)], ids=["ones", "u64-big", "u32-clip", "u64-saturating"])
def test_left_multiply_random_rectangular(gpu_ctx, oracle, cfg, bits, vmax):
    """Rectangular operands (no travelling window), empty rows in A, A rows of up to 40 entries (two blocks of lists), lists
    from 1 to 600 entries, values up to the saturation boundary."""
    rng = np.random.default_rng(7)
    a_h = random_csr(rng, 300, 500, 40, min(vmax, 1 << 20), bits, empty_every=7)
    b_h = random_csr(rng, 500, 2000, 600, vmax, bits, empty_every=11)
    cfg(pipeline=6)
    left_check(oracle, gpu_ctx, a_h, b_h, f"u{bits} vmax {vmax}")


def test_left_multiply_wide_column_space_falls_back(gpu_ctx, oracle, cfg):
    """A window of more than 65536 columns does not fit the kernel's 16-bit offsets: the engine picks another pipeline."""
    rng = np.random.default_rng(3)
    a_h = random_csr(rng, 64, 64, 3, 5, 64)
    b_h = random_csr(rng, 64, 200_000, 300, 5, 64)
    cfg(pipeline=6)
    a, b = B200Matrix.from_host(a_h, gpu_ctx), B200Matrix.from_host(b_h, gpu_ctx)
    c = a.matmul(b, want_stats=True)
    assert c.last_stats.pipeline != 6
    assert_same(c.to_host(), oracle.matmul_par(to_o(oracle, a_h), to_o(oracle, b_h)))


@pytest.mark.parametrize("bits", [64, 32])
@pytest.mark.parametrize("swap", [1, 0])
def test_power_chain_commuting_swap_bit_exact(gpu_ctx, oracle, cfg, bits, swap):
    """The reference's chain A^k = A^(k-1) x A (src/graph_magnus.rs:740-786) on its own 30^3 operand: with the swap the later
    steps are evaluated as A x A^(k-1) by the left multiply, and every power is still the oracle's, byte for byte."""
    cfg(commute_swap=swap)
    a_h = hostgen.reference_bench_instance(30, 3.0, bits)
    a = B200Matrix.from_host(a_h, gpu_ctx)
    a_o = to_o(oracle, a_h)
    p, p_o, gpu = a, a_o, []
    for _k in range(2, 8):
        p = p.matmul(a)
        gpu.append(p)
    for k, g in zip(range(2, 8), gpu):
        p_o = oracle.matmul_par(p_o, a_o)
        assert_same(g.to_host(), p_o, f"A^{k} u{bits} swap {swap}")
    assert [g.nnz() for g in gpu] == [251590, 655391, 1574848, 3383207, 6590100, 11736555]
    st = gpu[-1].device.product_stats()
    assert st.pipeline == (6 if swap else 4) and st.products == 23765080


def test_swap_needs_a_common_base(gpu_ctx, oracle, cfg):
    """Two uploads of the same matrix are different bases: nothing is known to commute, the operands stay in order; and a
    product of non-commuting matrices is evaluated in the order given."""
    a_h = hostgen.reference_bench_instance(14, 3.0, 64)
    a1, a2 = B200Matrix.from_host(a_h, gpu_ctx), B200Matrix.from_host(a_h, gpu_ctx)
    p = a1
    for _ in range(4):
        p = p.matmul(a1)
    c = p.matmul(a2, want_stats=True)                         # A^5 (of base a1) x a2: no common base
    assert c.last_stats.pipeline != 6
    rng = np.random.default_rng(11)
    x_h = random_csr(rng, a_h.rows, a_h.cols, 3, 4, 64)
    x = B200Matrix.from_host(x_h, gpu_ctx)
    px = p.matmul(x, want_stats=True)
    want = oracle.matmul_par(to_o(oracle, p.to_host()), to_o(oracle, x_h))
    assert_same(px.to_host(), want, "A^5 x X")
    xp = x.matmul(p, want_stats=True)                         # short rows on the left, long on the right: pipeline 6 by shape
    assert xp.last_stats.pipeline == 6
    assert_same(xp.to_host(), oracle.matmul_par(to_o(oracle, x_h), to_o(oracle, p.to_host())), "X x A^5")


@pytest.mark.parametrize("bits", [64, 32])
def test_left_multiply_edge_shapes(gpu_ctx, oracle, cfg, bits):
    """Forced pipeline 6 on the shapes the reference's own tests use (src/graph_csr.rs:1000-1100: identity, empty, single
    entries) plus ragged ones: 1 x 1, identity, an empty left operand row block, a single full row, 33 columns (one bit past
    a bitmap word), a left row of 70 entries (three blocks of lists) against lists of one entry."""
    rng = np.random.default_rng(5)
    cfg(pipeline=6)
    ident = lambda n: hostgen.identity(n, bits)
    cases = [
        (ident(1), ident(1)),
        (ident(5), ident(5)),
        (random_csr(rng, 7, 33, 5, 9, bits), random_csr(rng, 33, 33, 33, 9, bits)),
        (random_csr(rng, 4, 90, 70, 3, bits), ident(90)),
        (hostgen.from_coo(3, 40, np.array([1] * 40, np.uint32), np.arange(40, dtype=np.uint32), np.full(40, 2, np.uint64), bits),
         random_csr(rng, 40, 64, 64, 1000, bits)),
        (random_csr(rng, 50, 50, 2, 1, bits, empty_every=2), random_csr(rng, 50, 300, 120, 7, bits, empty_every=3)),
    ]
    for i, (a_h, b_h) in enumerate(cases):
        a, b = B200Matrix.from_host(a_h, gpu_ctx), B200Matrix.from_host(b_h, gpu_ctx)
        c = a.matmul(b, want_stats=True)
        assert_same(c.to_host(), oracle.matmul_par(to_o(oracle, a_h), to_o(oracle, b_h)), f"case {i} u{bits} (pipeline {c.last_stats.pipeline})")


@pytest.mark.parametrize("arc", [(300, 400), (3600, 800), (0, 4000), (1234, 1), (3999, 2)])
def test_left_multiply_over_an_active_row_arc(gpu_ctx, oracle, cfg, arc):
    """A whole-size left operand whose rows outside an arc of the row circle are empty (a rank's rows of a sharded chain,
    possibly wrapping past the last row): the kernel hands out the rows of the arc only, the others come out empty."""
    from sparse_linear_algebra_tests_b200.distributed import restrict_rows
    O = oracle
    a_h = hostgen.thin(hostgen.lattice([40, 10, 10], True, 64), 3.0 / 26.0, bytes([42] * 32))
    ao = to_o(O, a_h)
    p4 = O.matmul_par(O.matmul_par(O.matmul_par(ao, ao), ao), ao)
    p4_h = hostgen.HostCsr(p4.rows, p4.cols, p4.row_ptr, p4.col_idx, p4.values)
    mask = np.zeros(a_h.rows, dtype=bool)
    start, length = arc                                        # rows (start + o) mod rows, o < length: the second and last wrap
    mask[np.arange(start, start + length) % a_h.rows] = True
    cfg(pipeline=6)
    c = left_check(O, gpu_ctx, restrict_rows(a_h, mask), p4_h, f"arc {arc}")
    lens = np.diff(c.to_host().row_ptr.astype(np.int64))
    assert not lens[~mask].any() and lens[mask].any()


@pytest.mark.parametrize("world", [2, 4])
def test_halo_power_chain_blocks_match_the_reference_chain(gpu_ctx, oracle, world):
    """distributed.HaloPowerChain on the CUDA engine, every rank of a `world`-rank job in turn on this GPU: rows [r0, r1) of
    A^2..A^6 through left multiplies over block + halo equal the blocks of the reference chain (graph_magnus.rs:758-772),
    and the late powers do take the left-multiply kernel (whole-size handles with empty rows outside the arc)."""
    from sparse_linear_algebra_tests_b200.distributed import CudaEngine, HaloPowerChain, product_balanced_cuts
    O = oracle
    a_h = hostgen.thin(hostgen.lattice([48, 10, 10], True, 64), 3.0 / 26.0, bytes([42] * 32))
    ao = to_o(O, a_h)
    ref, p = [], ao
    for _ in range(2, 7):
        p = O.matmul_par(p, ao)
        ref.append(p)
    cuts = product_balanced_cuts(O.row_products(ao, ao), world)
    eng = CudaEngine(gpu_ctx)
    for rank in range(world):
        r0, r1 = int(cuts[rank]), int(cuts[rank + 1])
        ch = HaloPowerChain(eng, a_h, r0, r1, 6)
        outs = ch.run()
        for k, (c, want) in zip(range(2, 7), zip(outs, ref)):
            got = eng.download(ch.block(c))
            assert_same(got, want.row_block(r0, r1), f"rank {rank}/{world} A^{k}")
        assert outs[-1].product_stats().pipeline == 6, f"rank {rank}/{world}: A^6 did not take the left multiply"
