"""Developer probe: one rank of an N-rank torus chain alone on one GPU, per power: the right multiply block(A^(k-1)) x A against
the halo chain's left multiply A_k x A^(k-1) (distributed.HaloPowerChain) -- pipeline taken, launches, device ms.
    python tools/halo_probe.py --gpus 8 --rank 3 [--side 30] [--max-power 7]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=8)
    ap.add_argument("--rank", type=int, default=0)
    ap.add_argument("--side", type=int, default=30)
    ap.add_argument("--max-power", type=int, default=7)
    ap.add_argument("--reps", type=int, default=4)
    a = ap.parse_args()
    import torch
    from sparse_linear_algebra_tests_b200 import Context, hostgen
    from sparse_linear_algebra_tests_b200.distributed import CudaEngine, HaloPowerChain
    ctx = Context(0)
    full = hostgen.lattice([a.side * a.gpus, a.side, a.side], True, 64)
    a_h = hostgen.thin(full, 3.0 / (full.nnz() / full.rows), bytes([42] * 32))
    A = ctx.upload(a_h.rows, a_h.cols, a_h.row_ptr, a_h.col_idx, a_h.values)
    cuts = ctx.shard_rows_by_products(A, A, a.gpus)
    r0, r1 = int(cuts[a.rank]), int(cuts[a.rank + 1])
    blk = ctx.row_block(A, r0, r1)
    hc = HaloPowerChain(CudaEngine(ctx), a_h, r0, r1, a.max_power, a_dev=A)
    print(f"rank {a.rank}/{a.gpus} rows [{r0},{r1}) need {[int(m.sum()) for m in hc.need[1:]]} overhead {hc.overhead:.3f}")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
    best = {}
    for rep in range(a.reps):
        for mode in ("right", "halo"):
            flush.fill_(1)
            torch.cuda.synchronize()
            p = blk if mode == "right" else A
            for i in range(a.max_power - 1):
                x, y = (p, A) if mode == "right" else (hc.left[i], p)
                c, st = ctx.spgemm(x, y, True)
                d = st.as_dict()
                key = (mode, i + 2)
                if key not in best or d["ms_total"] < best[key]["ms_total"]:
                    best[key] = d
                p = c
    for mode in ("right", "halo"):
        tot = 0.0
        for k in range(2, a.max_power + 1):
            d = best[(mode, k)]
            tot += d["ms_total"]
            print(f"{mode:5s} A^{k}: pipeline {d['pipeline']} launches {d['kernel_launches']:2d} ms {d['ms_total']:.4f} products {d['products']} nnz {d['nnz_c']}")
        print(f"{mode:5s} sum of per-multiply ms: {tot:.4f}")


if __name__ == "__main__":
    main()
