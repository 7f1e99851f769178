"""Worker of tests/test_gpu_distributed.py (one process per rank, launched through torch.distributed.run).

mode "gloo":  both ranks share cuda:0; torch.distributed (gloo) carries A and the gathered blocks through the host, the
              multiplies run on the CUDA engine -- the N > 1 plumbing with the real engine on a one-GPU box.
mode "nccl":  one GPU per rank; everything below the C ABI: b200_comm_broadcast_csr, product-balanced row blocks, the
              resident-block power chain, b200_comm_allgather_csr, and the squaring chain (power_until_stable).
Prints "RANK r OK" or raises.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def same(h, o):
    return np.array_equal(h.row_ptr, o.row_ptr) and np.array_equal(h.col_idx, o.col_idx) and np.array_equal(h.values, o.values)


def main():
    import torch.distributed as dist
    from oracle import oracle as O
    from sparse_linear_algebra_tests_b200 import Context, hostgen
    from sparse_linear_algebra_tests_b200.distributed import (CudaEngine, ShardedPowerChain, ShardedSquaring, allgather_csr, broadcast_host_csr,
                                                              make_comm)
    mode = sys.argv[1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    ctx = Context(int(os.environ.get("LOCAL_RANK", 0)) if mode == "nccl" else 0)
    eng = CudaEngine(ctx)
    a_h = hostgen.reference_bench_instance(12, 3.0, 64)
    a_o = O.Csr(a_h.rows, a_h.cols, a_h.row_ptr, a_h.col_idx, a_h.values)
    if mode == "gloo":
        a_b = broadcast_host_csr(a_h if rank == 0 else None, 0)
        chain = ShardedPowerChain(eng, a_b, rank, world)
        p_o = a_o
        for k in range(2, 6):
            blk = chain.step()
            p_o = O.matmul_par(p_o, a_o)
            g = allgather_csr(eng.download(blk))
            assert same(g, p_o), f"rank {rank}: A^{k} gathered from the CUDA row blocks differs from the oracle"
    else:
        comm = make_comm(ctx, rank, world)
        a_dev = comm.broadcast(eng.upload(a_h) if rank == 0 else None, 0)
        assert same(eng.download(a_dev), a_o), "broadcast operand differs"
        cuts = ctx.shard_rows_by_products(a_dev, a_dev, world)
        blk = ctx.row_block(a_dev, int(cuts[rank]), int(cuts[rank + 1]))
        p_o = a_o
        for k in range(2, 6):
            blk = ctx.spgemm(blk, a_dev)
            p_o = O.matmul_par(p_o, a_o)
            full = comm.allgather(blk)
            assert same(eng.download(full), p_o), f"rank {rank}: A^{k} (NCCL all-gather of the row blocks) differs from the oracle"
        tot = comm.allreduce([blk.nnz], "sum_u64")
        assert int(tot[0]) == p_o.nnz()
        # squaring chain: (A + I)^(2^k) until the pattern is stable (src/graph_csr.rs:561-575, test :931-939)
        chain_h = hostgen.from_coo(40, 40, list(range(39)) + list(range(40)), list(range(1, 40)) + list(range(40)), np.ones(79, np.uint64), 64)
        cur_o = O.Csr(40, 40, chain_h.row_ptr, chain_h.col_idx, chain_h.values)
        sq = ShardedSquaring(ctx, comm, comm.broadcast(eng.upload(chain_h) if rank == 0 else None, 0))
        final, steps = sq.run()
        k = 0
        while True:
            nxt = O.matmul_par(cur_o, cur_o); k += 1
            stable = nxt.nnz() == cur_o.nnz() and np.array_equal(nxt.col_idx, cur_o.col_idx) and np.array_equal(nxt.row_ptr, cur_o.row_ptr)
            cur_o = nxt
            if stable:
                break
        assert steps == k and same(eng.download(final), cur_o), f"rank {rank}: squaring chain differs ({steps} vs {k} steps)"
        comm.close()
    dist.barrier()
    print(f"RANK {rank} OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
