"""`"ab,bc->ac"` over two sparse matrices, routed to the GPU engine.

The reference's sparse einsum entry points (einsum-dyn/src/sparse.rs:70-150 `einsum_sparse_driven`, :721
`einsum_sparse_hash`, and the VM of :216-498) all restrict two 2-D sparse inputs to the matmul arrangement -- A's column
index is B's row index, the output is (A's row index, B's column index) -- and then run Gustavson's loop through
`Sparse2D::row_entry` calls on one thread.  That arrangement IS the hot path, so here the spec is validated the same
way and the contraction is one `b200_spgemm`; the result is a `B200Matrix` (the reference writes into a dense or
`CsrBuilder` output, src/graph_csr_builder.rs:53-85).  Arithmetic is the engine's saturating add/multiply, which equals
the reference's plain `+=`/`*` whenever nothing overflows (the reference's own einsum-vs-matmul tests,
src/graph_csr.rs:1593-1631, compare exactly these two).
"""
from __future__ import annotations

from .graph_b200 import B200Matrix


class InvalidSpec(ValueError):
    """Malformed einsum specification (einsum-dyn/src/lib.rs `InvalidSpec`)."""


def parse_matmul_spec(spec: str):
    """Return (a0, a1, b0, b1, out) for a two-operand 2-D spec, or raise InvalidSpec."""
    s = spec.replace(" ", "")
    if "->" not in s:
        raise InvalidSpec(f"missing '->' in {spec!r}")
    lhs, out = s.split("->", 1)
    ins = lhs.split(",")
    if len(ins) != 2:
        raise InvalidSpec(f"expected 2 operands, got {len(ins)} in {spec!r}")
    for term in ins + [out]:
        if not term.isalpha():
            raise InvalidSpec(f"indices must be letters in {spec!r}")
    if len(ins[0]) != 2 or len(ins[1]) != 2:
        raise AssertionError("sparse-driven requires 2D inputs")            # the reference asserts (sparse.rs:89-90)
    for ch in out:
        if ch not in lhs:
            raise InvalidSpec(f"output index {ch!r} does not appear in the inputs of {spec!r}")
    return ins[0][0], ins[0][1], ins[1][0], ins[1][1], out


def einsum_sparse_driven(spec: str, a: B200Matrix, b: B200Matrix) -> B200Matrix:
    """C = einsum(spec, a, b) for the matmul arrangement (`"ab,bc->ac"` up to renaming); anything else panics like the
    reference (einsum-dyn/src/sparse.rs:99-112)."""
    a0, a1, b0, b1, out = parse_matmul_spec(spec)
    if a1 != b0:
        raise AssertionError(f"sparse-driven einsum requires the matmul pattern (A's column index = B's row index), got {spec!r}")
    if out != a0 + b1 or a0 == a1 or b0 == b1 or a0 == b1:
        raise AssertionError(f"sparse-driven einsum supports only 'ab,bc->ac'-shaped specs, got {spec!r}")
    if a.shape[1] != b.shape[0]:
        raise InvalidSpec(f"dimension mismatch on index {a1!r}: {a.shape[1]} vs {b.shape[0]}")
    return a.matmul(b)


einsum_sparse_hash = einsum_sparse_driven      # same contract, different CPU accumulator in the reference (sparse.rs:721)
