"""Developer probe: pinned D2H bandwidth by copy size, and where the end-to-end step of bench.py spends its time."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
dev = torch.device("cuda", 0)
for mb in (1, 8, 47, 94, 141, 282):
    n = mb << 20
    d = torch.empty(n, dtype=torch.uint8, device=dev); h = torch.empty(n, dtype=torch.uint8).pin_memory()
    for _ in range(2): h.copy_(d, non_blocking=True); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(f"D2H {mb:4d} MiB: {dt*1e3:7.3f} ms  {n/dt/1e9:6.1f} GB/s", flush=True)
from sparse_linear_algebra_tests_b200 import Context, hostgen
for bits in (64, 32):
    ctx = Context(0)
    a_h = hostgen.reference_bench_instance(30, 3.0, bits)
    A = ctx.upload(a_h.rows, a_h.cols, a_h.row_ptr, a_h.col_idx, a_h.values)
    p, powers = A, []
    for k in range(6): p = ctx.spgemm(p, A); powers.append(p)
    ctx.synchronize()
    vb = bits // 8
    bufs = [(torch.empty((a_h.rows + 1) * 8, dtype=torch.uint8).pin_memory(), torch.empty(c.nnz * 4, dtype=torch.uint8).pin_memory(), torch.empty(c.nnz * vb, dtype=torch.uint8).pin_memory()) for c in powers]
    for rep in range(3):
        ctx.synchronize(); t0 = time.perf_counter()
        for c, b in zip(powers, bufs): c.download_async_into(b[0].data_ptr(), b[1].data_ptr(), b[2].data_ptr())
        ctx.synchronize(); dt = time.perf_counter() - t0
        tot = sum(x.numel() for b in bufs for x in b)
        print(f"u{bits}: download of all six powers ({tot/1e6:.1f} MB): {dt*1e3:.3f} ms = {tot/dt/1e9:.1f} GB/s", flush=True)
    del powers, A
