"""Developer probe: A^2..A^k on the reference bench instance with per-phase device times + oracle check."""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparse_linear_algebra_tests_b200 import B200Matrix, Context, hostgen, set_default_context

ap = argparse.ArgumentParser()
ap.add_argument("--side", type=int, default=30); ap.add_argument("--epn", type=float, default=3.0)
ap.add_argument("--steps", type=int, default=7); ap.add_argument("--bits", type=int, default=64)
ap.add_argument("--check", type=int, default=1); ap.add_argument("--iters", type=int, default=5)
args = ap.parse_args()
ctx = Context(0); set_default_context(ctx)
a_h = hostgen.reference_bench_instance(args.side, args.epn, args.bits)
a = B200Matrix.from_host(a_h)
if args.check:
    from oracle import oracle as O
    a_o = O.Csr(a_h.rows, a_h.cols, a_h.row_ptr, a_h.col_idx, a_h.values); p_o = a_o
p = a
print(f"side={args.side} n={a.n} nnz={a.nnz()} bits={args.bits}")
for k in range(2, args.steps + 1):
    best = None
    for it in range(args.iters):
        c = p.matmul(a, want_stats=True); st = c.last_stats
        if best is None or st.ms_total < best.ms_total: best = st
    d = best.as_dict()
    ok = ""
    if args.check:
        p_o = O.matmul_par(p_o, a_o); h = c.to_host()
        ok = "OK" if (np.array_equal(h.row_ptr, p_o.row_ptr) and np.array_equal(h.col_idx, p_o.col_idx) and np.array_equal(h.values, p_o.values)) else "MISMATCH"
    gbs = d["bytes_algorithmic"] / (d["ms_total"] * 1e-3) / 1e9
    print(f"A^{k}: nnz={d['nnz_c']} prod={d['products']} sym={d['ms_symbolic']:.3f}ms num={d['ms_numeric']:.3f}ms total={d['ms_total']:.3f}ms "
          f"launches={d['kernel_launches']} mode={d['acc_mode']} {gbs:.0f}GB/s Mprod/s={d['products']/d['ms_total']/1e3:.0f} {ok}")
    print("    sym bins", d["sym_bin_rows"][:10], "num bins", d["num_bin_rows"][:10])
    p = c
