"""N > 1 host logic on CPU: world_size 2 over gloo.  The multiply is injected (CPU oracle) -- the product's
CudaEngine needs a B200 -- everything else (broadcast of A, product-balanced row blocks, resident-block power
chain without communication -- as right multiplies and as the halo chain of left multiplies --, variable-size all-gather
of C) is the code the GPU ranks run."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleEngine:
    def __init__(self):
        from oracle import oracle as O
        self.O = O

    def upload(self, h):
        return self.O.Csr(h.rows, h.cols, h.row_ptr, h.col_idx, h.values)

    def download(self, m):
        from sparse_linear_algebra_tests_b200 import hostgen
        return hostgen.HostCsr(m.rows, m.cols, m.row_ptr, m.col_idx, m.values)

    def row_products(self, a, b):
        return self.O.row_products(a, b)

    def row_block(self, a, r0, r1):
        return a.row_block(r0, r1)

    def spgemm(self, a, b):
        return self.O.matmul(a, b)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from sparse_linear_algebra_tests_b200 import hostgen
        from sparse_linear_algebra_tests_b200.distributed import ShardedPowerChain, allgather_csr, broadcast_host_csr
        eng = OracleEngine()
        a_h = hostgen.reference_bench_instance(8, 3.0, 64) if rank == 0 else None
        a_h = broadcast_host_csr(a_h, 0)
        chain = ShardedPowerChain(eng, a_h, rank, world)
        full = eng.upload(a_h)
        ok = True
        for k in range(2, 5):
            blk = chain.step()
            full = eng.spgemm(full, eng.upload(a_h))
            gathered = allgather_csr(eng.download(blk))
            ok &= (np.array_equal(gathered.row_ptr, full.row_ptr) and np.array_equal(gathered.col_idx, full.col_idx)
                   and np.array_equal(gathered.values, full.values))
        # the same blocks through left multiplies over block + halo (HaloPowerChain): nothing exchanged between the ranks
        from sparse_linear_algebra_tests_b200.distributed import HaloPowerChain
        halo = HaloPowerChain(eng, a_h, chain.r0, chain.r1, 4)
        full = eng.upload(a_h)
        for c in halo.run():
            full = eng.spgemm(full, eng.upload(a_h))
            gathered = allgather_csr(eng.download(halo.block(c)))
            ok &= (np.array_equal(gathered.row_ptr, full.row_ptr) and np.array_equal(gathered.col_idx, full.col_idx)
                   and np.array_equal(gathered.values, full.values))
        loads = [None] * world
        dist.all_gather_object(loads, int(chain.row_products[chain.r0:chain.r1].sum()))
        q.put((rank, ok, chain.r0, chain.r1, loads))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_power_chain_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), "all-gathered row blocks differ from the single-process product"
    assert res[0][2] == 0 and res[0][3] == res[1][2] and res[1][3] == 512        # contiguous cover of the 8^3 rows
    loads = res[0][4]
    assert abs(loads[0] - loads[1]) <= 0.05 * sum(loads) + 64, f"product split not balanced: {loads}"
