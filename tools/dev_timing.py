"""Developer probe: wall / event time of the chain without engine-side stats (the bench's timed-loop conditions)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparse_linear_algebra_tests_b200 import Context, hostgen
dev = torch.device("cuda", 0); stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream)
ctx = Context(0, stream.cuda_stream)
if len(sys.argv) > 1:
    ctx.configure(**{k: int(v) for k, v in (kv.split("=") for kv in sys.argv[1].split(","))})
a_h = hostgen.reference_bench_instance(30, 3.0, 64)
A = ctx.upload(a_h.rows, a_h.cols, a_h.row_ptr, a_h.col_idx, a_h.values)
def chain(stats):
    p, keep, tw = A, [], []
    for k in range(2, 8):
        t0 = time.perf_counter()
        c = ctx.spgemm(p, A, True)[0] if stats else ctx.spgemm(p, A)
        tw.append((time.perf_counter() - t0) * 1e3)
        keep.append(c); p = c
    return keep, tw
for stats in (True, False, True, False):
    for it in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        e0.record(stream); keep, tw = chain(stats); e1.record(stream); e1.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        print(f"stats={stats} it={it} event={e0.elapsed_time(e1):8.3f} ms wall={wall:8.3f} ms per-call wall ms: " + " ".join(f"{x:.2f}" for x in tw))
        del keep
