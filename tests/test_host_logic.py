"""Host-side logic (no GPU): builders against the oracle's independent C versions, sharding rule, ABI surface."""
import ctypes
import os
import re

import numpy as np
import pytest

from sparse_linear_algebra_tests_b200 import EXPORTS, LIB_PATH, hostgen
from sparse_linear_algebra_tests_b200.distributed import product_balanced_cuts

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def same(h, o):
    return (h.rows == o.rows and np.array_equal(h.row_ptr, o.row_ptr) and np.array_equal(h.col_idx, o.col_idx)
            and h.values.dtype == o.values.dtype and np.array_equal(h.values, o.values))


@pytest.mark.parametrize("bits", [32, 64])
@pytest.mark.parametrize("dims,torus", [([5], False), ([5], True), ([3, 3], True), ([2, 2, 2], False), ([2, 2, 2], True), ([4, 3], False), ([6, 5, 4], True)])
def test_lattice_matches_oracle(oracle, dims, torus, bits):
    assert same(hostgen.lattice(dims, torus, bits), oracle.lattice(dims, torus, bits))


@pytest.mark.parametrize("bits", [32, 64])
def test_thin_and_bench_instance_match_oracle(oracle, bits):
    for side, epn in ((6, 3.0), (10, 4.0), (12, 8.0)):
        assert same(hostgen.reference_bench_instance(side, epn, bits), oracle.reference_bench_instance(side, epn, bits))
    full = hostgen.lattice([7, 7, 7], True, bits)
    t = hostgen.thin(full, 0.3, bytes(range(32)))
    assert same(t, oracle.thin_stdrng(oracle.lattice([7, 7, 7], True, bits), bytes(range(32)), 0.3))
    rows = t.row_of_entry()
    for r, c, v in zip(rows.tolist(), t.col_idx.tolist(), t.values.tolist()):
        assert t.get(c, r) == v                                    # thinning preserves symmetry (graph_csr.rs:224)


@pytest.mark.parametrize("dims", [[5, 4, 3], [6, 6, 6], [7, 3], [9]])
def test_thinned_torus_equals_lattice_then_thin(dims):
    """The large-instance builder (200^3 config) skips the full-lattice COO sort but must give the same matrix."""
    for bits, density, seed in ((32, 3.0 / 26.0, bytes([42] * 32)), (64, 0.4, bytes(range(32)))):
        want = hostgen.thin(hostgen.lattice(dims, True, bits), density, seed)
        assert same(hostgen.thinned_torus(dims, density, seed, bits), want)


def test_portable_generators_match_oracle(oracle):
    assert same(hostgen.lattice_csr_xorshift(8, 3.0, 42), oracle.lattice_csr_xorshift(8, 3.0, 42))
    assert same(hostgen.rmat(9, 8, 0.57, 0.19, 0.19, 42, 64), oracle.rmat(9, 8, 0.57, 0.19, 0.19, 42, 64))
    assert same(hostgen.rmat(8, 4, 0.45, 0.15, 0.15, 7, 32), oracle.rmat(8, 4, 0.45, 0.15, 0.15, 7, 32))


def test_from_coo_semantics(oracle):
    h = hostgen.from_coo(4, 4, [3, 0, 0, 2, 0], [1, 2, 2, 2, 0], np.array([1, 2, 3, 0, 9], np.uint32), 32)
    assert (h.get(0, 2), h.get(0, 0), h.get(3, 1), h.nnz()) == (5, 9, 1, 3) and h.row_ptr.tolist() == [0, 2, 2, 2, 3]
    big = np.array([0xFFFFFFFF, 2], np.uint32)
    wrap = hostgen.from_coo(1, 1, [0, 0], [0, 0], big, 32)                    # plain += (graph_csr.rs:93) wraps in release
    sat = hostgen.from_coo(1, 1, [0, 0], [0, 0], big, 32, saturating=True)   # linalg/src/csr.rs:167
    assert wrap.get(0, 0) == 1 and sat.get(0, 0) == 0xFFFFFFFF
    assert same(wrap, oracle.from_coo(1, 1, [0, 0], [0, 0], big, 32, False)) and same(sat, oracle.from_coo(1, 1, [0, 0], [0, 0], big, 32, True))
    with pytest.raises(IndexError):
        hostgen.from_coo(2, 2, [2], [0], [1])


def test_product_balanced_cuts(oracle):
    a = oracle.reference_bench_instance(10, 3.0, 64)
    p = oracle.row_products(a, a)
    for parts in (1, 2, 3, 8):
        cuts = product_balanced_cuts(p, parts)
        assert cuts[0] == 0 and cuts[-1] == a.rows and np.all(np.diff(cuts.astype(np.int64)) >= 0)
        loads = [int(p[int(cuts[i]):int(cuts[i + 1])].sum()) + int(cuts[i + 1] - cuts[i]) for i in range(parts)]
        assert max(loads) - min(loads) <= 2 * (int(p.max()) + 1)
    skew = np.array([1000] + [0] * 99, dtype=np.uint64)            # one heavy row: it gets a part of its own
    cuts = product_balanced_cuts(skew, 4)
    assert cuts[1] == 1


def test_shared_library_loads_and_exports_every_declared_symbol():
    assert os.path.exists(LIB_PATH), "build the CUDA engine first (python -c 'import __graft_entry__ as g; g.build()')"
    lib = ctypes.CDLL(LIB_PATH)
    header = open(os.path.join(ROOT, "include", "b200_spgemm.h")).read()
    declared = sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no declarations found in include/b200_spgemm.h"
    for name in declared:
        assert hasattr(lib, name), f"libb200spgemm.so does not export {name}"
    assert sorted(EXPORTS) == declared


def test_engine_fails_loudly_without_a_gpu():
    """No CPU fallback: on a box without a CUDA device the context cannot be created."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from sparse_linear_algebra_tests_b200 import B200Error, Context
    with pytest.raises(B200Error) as e:
        Context(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_einsum_spec_validation_needs_no_gpu():
    """Spec handling of the einsum front-end (einsum-dyn/src/sparse.rs:81-112): malformed specs are InvalidSpec, specs
    that are well-formed but not the matmul arrangement panic."""
    from sparse_linear_algebra_tests_b200.einsum import InvalidSpec, parse_matmul_spec
    assert parse_matmul_spec("ab,bc->ac") == ("a", "b", "b", "c", "ac")
    assert parse_matmul_spec(" ij , jk -> ik ") == ("i", "j", "j", "k", "ik")
    for bad in ("ab,bc", "ab->ab", "ab,bc,cd->ad", "a1,bc->ac", "ab,bc->ad"):
        with pytest.raises(InvalidSpec):
            parse_matmul_spec(bad)
    with pytest.raises(AssertionError):
        parse_matmul_spec("abc,cd->abd")


def test_thin_with_a_shared_generator_consumes_the_stream_in_order():
    """bench_matmul_magnus thins every grid point with ONE StdRng (src/graph_magnus.rs:800-821): `skip` continues the
    stream where the previous thin stopped, one draw per stored entry with r <= c."""
    full = hostgen.lattice([5, 5, 5], True, 64)
    n = hostgen.draws_of_thin(full)
    assert n == int((full.row_of_entry() <= full.col_idx.astype(np.int64)).sum()) and n == full.nnz() // 2    # no diagonal in a lattice
    seed = bytes([42] * 32)
    stream = hostgen.stdrng_f64_unit(seed, 2 * n)
    first, second = hostgen.thin(full, 0.3, seed), hostgen.thin(full, 0.3, seed, skip=n)
    rows, cols = full.row_of_entry(), full.col_idx.astype(np.int64)
    upper = np.flatnonzero(rows <= cols)
    for got, draws in ((first, stream[:n]), (second, stream[n:])):
        kept = upper[draws < 0.3]
        want = {(int(rows[i]), int(cols[i])) for i in kept} | {(int(cols[i]), int(rows[i])) for i in kept}
        have = set(zip(got.row_of_entry().tolist(), got.col_idx.tolist()))
        assert have == want
    assert not same(first, second)


def test_engine_chacha12_host_twin_matches_the_pinned_generator():
    """gen.cuh's ChaCha12 is one __host__ __device__ function; its host side is exported (b200_stdrng_u64) so that the
    generator the device thinning kernels run can be pinned without a GPU: StdRng::from_seed next_u64 outputs must equal
    hostgen's (which reproduce the README's nnz column through the oracle), from any starting draw."""
    from sparse_linear_algebra_tests_b200 import _native
    for seed in (bytes([42] * 32), bytes(range(32)), bytes([0] * 32)):
        want = hostgen.stdrng_u64(seed, 5000)
        assert np.array_equal(_native.stdrng_u64(seed, 0, 5000), want)
        assert np.array_equal(_native.stdrng_u64(seed, 4093, 907), want[4093:])
    # counters beyond 2^32 blocks use the high counter word
    far = (1 << 35) + 5
    assert np.array_equal(_native.stdrng_u64(bytes([42] * 32), far, 16)[8:], _native.stdrng_u64(bytes([42] * 32), far + 8, 8))


def test_halo_chain_gives_the_ranks_row_blocks(oracle):
    """distributed.HaloPowerChain (left multiplies over a rank's rows plus a halo, nothing exchanged): rows [r0, r1) of every
    power equal the rank's block of the reference chain A^k = A^(k-1) x A (graph_magnus.rs:758-772), also for a block whose
    halo wraps round the torus; the row sets shrink to the block at the last power; without locality the halo is everything."""
    from sparse_linear_algebra_tests_b200.distributed import HaloPowerChain, halo_row_sets, restrict_rows

    class Eng:
        def upload(self, h): return oracle.Csr(h.rows, h.cols, h.row_ptr, h.col_idx, h.values)
        def spgemm(self, a, b): return oracle.matmul(a, b)
        def row_block(self, a, r0, r1): return a.row_block(r0, r1)

    a = hostgen.thin(hostgen.lattice([24, 4, 4], True, 64), 3.0 / 26.0, bytes([42] * 32))
    full = Eng().upload(a)
    ref, p = [], full
    for _ in range(2, 6):
        p = oracle.matmul(p, full)
        ref.append(p)
    for r0, r1 in ((0, 96), (96, 200), (288, 384)):
        ch = HaloPowerChain(Eng(), a, r0, r1, 5)
        assert int(ch.need[5].sum()) == r1 - r0 and all(ch.need[k][r0:r1].all() for k in range(1, 6))
        assert all((ch.need[k] <= ch.need[k - 1]).all() for k in range(2, 6))
        assert 0.0 < ch.overhead < 1.5
        for c, want in zip(ch.run(), ref):
            b, w = ch.block(c), want.row_block(r0, r1)
            assert np.array_equal(b.row_ptr, w.row_ptr) and np.array_equal(b.col_idx, w.col_idx) and np.array_equal(b.values, w.values)
    r = restrict_rows(a, halo_row_sets(a, 10, 20, 3)[3])
    assert r.rows == a.rows and r.nnz() == int(a.row_ptr[20] - a.row_ptr[10]) and int(r.row_ptr[10]) == 0
    g = hostgen.rmat(8, 8, 0.45, 0.15, 0.15, 42, 64) if hasattr(hostgen, "rmat") else None
    if g is not None:
        assert halo_row_sets(g, 0, 32, 4)[1].sum() > 0.5 * g.rows
