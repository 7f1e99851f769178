"""BASELINE.json configs[2]: the side x e_per_n sweep of `bench_matmul_magnus` (src/graph_magnus.rs:790-929) with the GPU
engine as one more column.  A * A on the side^3 Moore torus thinned to e_per_n edges/node, one StdRng([42;32]) shared
across the whole grid exactly like the reference loop (:800, :817-821), so every instance is the reference's.

Prints the reference's CSV schema (csv2table.py / plot_surface.py parse it by header name) with the CPU columns the
reference gets from its other matrix types replaced by what can run here: `csr_par_us` = the OpenMP restatement of
CsrMatrix::matmul_par, `csr_us` = the sequential restatement; `b200_us` = device time of b200_spgemm (best of ITERS),
`x_*` = csr_us / t as in the reference's `x` closure but against the sequential CSR time (the BTreeMap baseline the
reference divides by is out of scope).  Every GPU result is compared bit for bit with the oracle before timing counts.
"""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparse_linear_algebra_tests_b200 import B200Matrix, Context, hostgen, set_default_context
from oracle import oracle as O

ap = argparse.ArgumentParser()
ap.add_argument("--sides", type=int, nargs="+", default=[5, 10, 20, 30])
ap.add_argument("--epn", type=float, nargs="+", default=[2.0, 3.0, 4.0, 8.0, 26.0])
ap.add_argument("--iters", type=int, default=10); ap.add_argument("--bits", type=int, default=64)
args = ap.parse_args()
ctx = Context(0); set_default_context(ctx)
seed, used = bytes([42] * 32), 0
nt = O.max_threads()
print("side,nodes,e_per_n,nnz,products,nnz_c,csr_us,csr_par_us,b200_us,x_csr_par,x_b200,gprod_per_s,hbm_frac,launches,cores")
for s in args.sides:
    full = hostgen.lattice([s, s, s], True, args.bits)
    full_epn = full.nnz() / full.rows
    for epn in args.epn:
        density = epn / full_epn
        if density >= 1.0:
            a_h = full
        else:
            a_h = hostgen.thin(full, density, seed, skip=used)
            used += hostgen.draws_of_thin(full)
        a_o = O.Csr(a_h.rows, a_h.cols, a_h.row_ptr, a_h.col_idx, a_h.values)
        want = O.matmul(a_o, a_o)
        t_seq = O.time_matmul(a_o, a_o, False, 1, args.iters)
        t_par = O.time_matmul(a_o, a_o, True, nt, args.iters)
        a = B200Matrix.from_host(a_h)
        best = None
        for _ in range(args.iters + 1):                                  # 1 warm-up + ITERS, as the reference (:856)
            c = a.matmul(a, want_stats=True)
            if best is None or c.last_stats.ms_total < best.ms_total:
                best = c.last_stats
        h = c.to_host()
        assert np.array_equal(h.row_ptr, want.row_ptr) and np.array_equal(h.col_idx, want.col_idx) and np.array_equal(h.values, want.values), (s, epn)
        d = best.as_dict()
        us = d["ms_total"] * 1e3
        frac = d["bytes_algorithmic"] / (d["ms_total"] * 1e-3) / 1e9 / 6554.9
        print(f"{s},{a_h.rows},{epn:.0f},{a_h.nnz()},{d['products']},{d['nnz_c']},{t_seq * 1e6:.0f},{t_par * 1e6:.0f},{us:.1f},"
              f"{t_seq / t_par:.4f},{t_seq * 1e6 / us:.4f},{d['products'] / d['ms_total'] / 1e6:.2f},{frac:.4f},{d['kernel_launches']},{nt}", flush=True)
