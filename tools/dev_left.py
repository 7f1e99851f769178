"""Developer probe: the power chain multiplied from the left (A^k = A . A^(k-1)) against the reference order (A^(k-1) . A):
per-power device times of both orders and a bit compare of the two results."""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparse_linear_algebra_tests_b200 import B200Matrix, Context, hostgen, set_default_context

ap = argparse.ArgumentParser()
ap.add_argument("--side", type=int, default=30); ap.add_argument("--epn", type=float, default=3.0)
ap.add_argument("--steps", type=int, default=7); ap.add_argument("--bits", type=int, default=64)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--cfg", default="", help="b200_config fields for the left order, e.g. pipeline=4")
args = ap.parse_args()
ctx = Context(0); set_default_context(ctx)
a_h = hostgen.reference_bench_instance(args.side, args.epn, args.bits)
a = B200Matrix.from_host(a_h)
print(f"side={args.side} n={a.n} nnz={a.nnz()} bits={args.bits}")


def best_of(f):
    best = None; c = None
    for _ in range(args.iters):
        c = f(); st = c.last_stats
        if best is None or st.ms_total < best.ms_total: best = st
    return c, best.as_dict()


p = a
for k in range(2, args.steps + 1):
    ctx.configure(pipeline=0)
    cr, dr = best_of(lambda: p.matmul(a, want_stats=True))
    if args.cfg:
        ctx.configure(**{kk: int(v) for kk, v in (kv.split("=") for kv in args.cfg.split(","))})
    cl, dl = best_of(lambda: a.matmul(p, want_stats=True))
    hr, hl = cr.to_host(), cl.to_host()
    same = np.array_equal(hr.row_ptr, hl.row_ptr) and np.array_equal(hr.col_idx, hl.col_idx) and np.array_equal(hr.values, hl.values)
    print(f"A^{k}: nnz={dr['nnz_c']} right: prod={dr['products']} {dr['ms_total']:.3f} ms (pipeline {dr['pipeline']}, {dr['kernel_launches']} launches, mode {dr['acc_mode']})"
          f"   left: prod={dl['products']} {dl['ms_total']:.3f} ms (pipeline {dl['pipeline']}, {dl['kernel_launches']} launches, mode {dl['acc_mode']})   {'SAME' if same else 'DIFFERENT'}")
    p = cr
