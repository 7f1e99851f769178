"""Developer probe: A^2 of one rank's row block of an R-MAT graph on one GPU (B200_TRACE=1 prints the kernel timeline)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparse_linear_algebra_tests_b200 import Context
scale, r, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
ctx = Context(0)
a = ctx.rmat(scale, 16, 0.45, 0.15, 0.15, 42, 64)
cuts = ctx.shard_rows_by_products(a, a, n)
blk = ctx.row_block(a, int(cuts[r]), int(cuts[r + 1]))
print("rows of the block", blk.rows, "nnz", blk.nnz, flush=True)
for it in range(3):
    c, st = ctx.spgemm(blk, a, True)
    d = st.as_dict()
    print(f"block {r}/{n}: products {d['products']} nnz {d['nnz_c']} maxP {d['max_row_products']} ms {d['ms_total']:.2f} bins {d['sym_bin_rows']}", flush=True)
    c.free()
