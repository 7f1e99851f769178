"""Large configurations of BASELINE.json on one GPU (configs[3], configs[4]), checked bit-for-bit against the CPU oracle.

  torus : A^2..A^k on a side^3 Moore torus thinned to ~epn edges/node (StdRng([42;32])), A^k = A^(k-1) * A
  rmat  : A^2 of an R-MAT graph (2^scale nodes, ef edges/node, quadrant probabilities a/b/c)

Prints one line per multiply: products, nnz, device ms, products/s, algorithmic GB/s and its fraction of the measured
HBM copy bandwidth; with --check the result is downloaded and compared with oracle.matmul_par (all host cores).
"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparse_linear_algebra_tests_b200 import B200Matrix, Context, hostgen, set_default_context

ap = argparse.ArgumentParser()
ap.add_argument("kind", choices=["torus", "rmat"])
ap.add_argument("--side", type=int, default=200); ap.add_argument("--epn", type=float, default=3.0); ap.add_argument("--power", type=int, default=5)
ap.add_argument("--scale", type=int, default=20); ap.add_argument("--ef", type=int, default=16)
ap.add_argument("--abc", type=float, nargs=3, default=[0.45, 0.15, 0.15])
ap.add_argument("--bits", type=int, default=64); ap.add_argument("--check", type=int, default=1); ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--device-build", type=int, default=0, help="torus: build the instance with the engine's device generators (b200_lattice + b200_thin) instead of hostgen")
args = ap.parse_args()

peaks = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
peak = float(json.load(open(peaks))["hbm_gbs"]) if os.path.exists(peaks) else 6650.0

t0 = time.time()
ctx = Context(0); set_default_context(ctx)
a = None
if args.kind == "torus" and args.device_build:
    a = B200Matrix.lattice([args.side] * 3, True, args.bits, ctx).thin(args.epn / 26.0, bytes([42] * 32))
    ctx.synchronize()
    a_h = a.to_host() if args.check else None
    steps = list(range(2, args.power + 1))
    print(f"torus (device-built): n={a.n} nnz={a.nnz()} built in {time.time() - t0:.2f}s", flush=True)
elif args.kind == "torus":
    a_h = hostgen.thinned_torus([args.side] * 3, args.epn / 26.0, bytes([42] * 32), args.bits)
    steps = list(range(2, args.power + 1))
else:
    a_h = hostgen.rmat(args.scale, args.ef, args.abc[0], args.abc[1], args.abc[2], 42, args.bits)
    steps = [2]
if a is None:
    print(f"{args.kind}: n={a_h.rows} nnz={a_h.nnz()} built in {time.time() - t0:.1f}s", flush=True)
    a = B200Matrix.from_host(a_h)
if args.check:
    from oracle import oracle as O
    a_o = O.Csr(a_h.rows, a_h.cols, a_h.row_ptr, a_h.col_idx, a_h.values); p_o = a_o
p = a
for k in steps:
    best = None
    for it in range(args.iters):
        c = p.matmul(a, want_stats=True)
        if best is None or c.last_stats.ms_total < best.ms_total:
            best = c.last_stats
        if it + 1 < args.iters:
            del c
    d = best.as_dict()
    gbs = d["bytes_algorithmic"] / (d["ms_total"] * 1e-3) / 1e9
    ok = ""
    if args.check:
        t1 = time.time(); p_o = O.matmul_par(p_o, a_o, O.max_threads()); cpu_s = time.time() - t1
        h = c.to_host()
        same = np.array_equal(h.row_ptr, p_o.row_ptr) and np.array_equal(h.col_idx, p_o.col_idx) and np.array_equal(h.values, p_o.values)
        ok = f"{'BIT-EXACT' if same else 'MISMATCH'} (cpu oracle {cpu_s * 1e3:.0f} ms, {O.max_threads()} threads)"
        c._host = None
    print(f"A^{k}: products={d['products']} nnz={d['nnz_c']} maxrowP={d['max_row_products']} ms={d['ms_total']:.3f} "
          f"(pre+numeric+scan {d['ms_numeric']:.3f}, compaction {d['ms_symbolic']:.3f}) Gprod/s={d['products'] / d['ms_total'] / 1e6:.1f} "
          f"alg={gbs:.0f}GB/s frac={gbs / peak:.3f} launches={d['kernel_launches']} mode={d['acc_mode']} bins={d['sym_bin_rows'][:10]} {ok}", flush=True)
    p = c
