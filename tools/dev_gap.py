"""Developer probe: host wall time of every multiply of a torus chain beside its device time (where does a step wait?)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparse_linear_algebra_tests_b200 import Context
side = int(sys.argv[1]) if len(sys.argv) > 1 else 200
maxp = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ctx = Context(0)
full = ctx.lattice([side] * 3, True, 64)
a, _ = ctx.thin(full, 3.0 / 26.0, bytes([42] * 32))
full.free()
blk = a
if len(sys.argv) > 3:                                  # rank r of n: the product-balanced row block the multi-GPU bench gives that rank
    r, n = int(sys.argv[3]), int(sys.argv[4])
    cuts = ctx.shard_rows_by_products(a, a, n)
    blk = ctx.row_block(a, int(cuts[r]), int(cuts[r + 1]))
for rep in range(4):
    p, keep = blk, []
    ctx.synchronize(); t_step = time.perf_counter()
    for k in range(2, maxp + 1):
        t0 = time.perf_counter()
        c = ctx.spgemm(p, a)
        t1 = time.perf_counter()
        ctx.synchronize()
        t2 = time.perf_counter()
        st = c.product_stats()
        print(f"rep {rep} A^{k}: call {1e3*(t1-t0):8.2f} ms  call+sync {1e3*(t2-t0):8.2f} ms  device {st.ms_total:8.2f} ms (sym {st.ms_symbolic:.2f} num {st.ms_numeric:.2f}) nnz {st.nnz_c} launches {st.kernel_launches} pipeline {st.pipeline}", flush=True)
        keep.append(c); p = c
    t3 = time.perf_counter()
    for c in keep: c.free()
    ctx.synchronize()
    print(f"rep {rep}: chain {1e3*(t3-t_step):.2f} ms, frees {1e3*(time.perf_counter()-t3):.2f} ms", flush=True)
