"""Developer probe: A^2..A^k on the reference bench instance with per-phase device times + oracle check."""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparse_linear_algebra_tests_b200 import B200Matrix, Context, hostgen, set_default_context

ap = argparse.ArgumentParser()
ap.add_argument("--side", type=int, default=30); ap.add_argument("--epn", type=float, default=3.0)
ap.add_argument("--steps", type=int, default=7); ap.add_argument("--bits", type=int, default=64)
ap.add_argument("--check", type=int, default=1); ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--block", type=int, default=0, help="with --xmul N: multiply only rank 0's row block (the per-GPU work of the N-GPU run)")
ap.add_argument("--cfg", default="", help="b200_config fields, e.g. pipeline=1,fused_threads=128")
ap.add_argument("--xmul", type=int, default=1, help="torus is (side*xmul) x side x side: the weak-scaling workload of bench.py --gpus xmul, on one GPU")
args = ap.parse_args()
ctx = Context(0); set_default_context(ctx)
if args.cfg:
    ctx.configure(**{k: int(v) for k, v in (kv.split("=") for kv in args.cfg.split(","))})
if args.xmul == 1:
    a_h = hostgen.reference_bench_instance(args.side, args.epn, args.bits)
else:
    full = hostgen.lattice([args.side * args.xmul, args.side, args.side], True, args.bits)
    a_h = hostgen.thin(full, args.epn / (full.nnz() / full.rows), bytes([42] * 32))
a = B200Matrix.from_host(a_h)
if args.block:                                     # rank 0's share of bench.py --gpus xmul: the first 1/xmul of the rows (by products)
    cuts = ctx.shard_rows_by_products(a.device, a.device, args.xmul)
    p0 = B200Matrix(ctx.row_block(a.device, 0, int(cuts[1])))
    args.check = 0
if args.check:
    from oracle import oracle as O
    a_o = O.Csr(a_h.rows, a_h.cols, a_h.row_ptr, a_h.col_idx, a_h.values); p_o = a_o
p = p0 if args.block else a
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
    print("    pipeline", d["pipeline"], "rows per list / class", d["sym_bin_rows"])
    p = c
