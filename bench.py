#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 SpGEMM engine (driver contract: ONE JSON line on rank 0).

Workload (BASELINE.json configs[1], the reference's bench_repeated_exponentiation,
src/graph_magnus.rs:699-788): A^2..A^7 on the 30x30x30 Moore torus thinned to ~3 edges/node with
StdRng::from_seed([42;32]) -- the reference's exact operand (81 434 nnz) -- u64 saturating values,
A^k = A^(k-1) * A.  One "step" = the whole chain (6 multiplies, 45.8 M intermediate products).
metric = intermediate products / second (whole job); ms per A^k multiply is reported beside it.

  value     operands resident in HBM, CUDA-event time on the engine's stream (= torch's current
            stream), L2 flushed between steps, max over ranks
  e2e       same chain through the public API with HOST buffers: pinned H2D of A, the six multiplies,
            pinned D2H of every power (row_ptr, col_idx, values), wall clock
  roofline  the largest multiply of the step (A^7 = A^6 * A): algorithmic bytes over
            its CUDA-event time, against the measured HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline / --impl reference
            the CPU restatement of the reference's rayon CsrMatrix::matmul_par (oracle/, OpenMP, all
            host cores), reference protocol: 1 warm-up + 3 timed multiplies per power, wall clock
N > 1 (torchrun, one rank per GPU): weak scaling over the natural row sharding -- the torus grows to
(30*N) x 30 x 30, A is broadcast once over NCCL, rank r keeps its product-balanced row block of every
power resident and multiplies it by the replicated A; no data-path collective.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "intermediate_products_per_second"
UNIT = "products/s"
MAX_POWER = 7


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--side", type=int, default=30)
    ap.add_argument("--epn", type=float, default=3.0)
    ap.add_argument("--bits", type=int, default=64, choices=[32, 64])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    # the large configurations of BASELINE.json (not the driver's default line): a fixed-size torus cut over the ranks
    ap.add_argument("--strong", action="store_true", help="strong scaling: the torus stays side^3 whatever --gpus is (e.g. --side 200 --max-power 5)")
    ap.add_argument("--max-power", type=int, default=7)
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (its pinned staging is 12 B per result entry)")
    # BASELINE configs[3]: A^2 of an R-MAT graph built on the device (b200_rmat), rows sharded by product count; always strong
    # scaling (the graph does not grow with --gpus), device-timed only
    ap.add_argument("--chain", default="right", choices=["right", "halo", "auto"],
                    help="N > 1: a rank's steps as block(A^(k-1)) x A ('right', the default), as left multiplies over its block plus a halo "
                         "of neighbouring rows computed redundantly ('halo', no communication either), or whichever is faster in warm-up "
                         "('auto').  Measured on the 30^3-per-GPU chain the halo chain loses (DESIGN.md section 6): kept as an option")
    ap.add_argument("--as-rank", type=int, default=None, help="developer: run rank R of a --gpus N job alone on one GPU (no process group)")
    ap.add_argument("--workload", default="torus", choices=["torus", "rmat"])
    ap.add_argument("--scale", type=int, default=20)
    ap.add_argument("--ef", type=int, default=16)
    ap.add_argument("--abc", type=float, nargs=3, default=[0.45, 0.15, 0.15])
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def workload_config(args, world):
    if args.workload == "rmat":
        return {"workload": f"A^2 of an R-MAT graph, scale {args.scale} ({1 << args.scale} nodes), {args.ef} edges/node, quadrants "
                            f"{args.abc[0]:g}/{args.abc[1]:g}/{args.abc[2]:g}, splitmix64 seed 42 (BASELINE configs[3]; SURVEY.md App. C), built on the device",
                "val_bits": args.bits, "left": "row block of A per GPU", "right": "A (replicated, one NCCL broadcast below the C ABI)",
                "sharding": "contiguous row blocks balanced by intermediate-product count" if world > 1 else "single GPU",
                "l2": "flushed between timed steps (256 MiB write)"}
    dims = [args.side * (1 if args.strong else world), args.side, args.side]
    return {"workload": f"repeated exponentiation A^2..A^{MAX_POWER} on the {dims[0]}x{dims[1]}x{dims[2]} Moore torus, "
                        f"~{args.epn:g} e/n, StdRng([42;32]) thinning (graph_magnus.rs:699-788)",
            "dims": dims, "val_bits": args.bits, "left": "A^(k-1) (resident row block per GPU)", "right": "A (replicated)",
            "sharding": "contiguous row blocks balanced by intermediate-product count" if world > 1 else "single GPU",
            "l2": "flushed between timed steps (256 MiB write); within a step A^(k-1) is L2-warm from the previous multiply, as in the reference loop"}


def build_operand(args, world, ctx=None):
    """The operand of bench_repeated_exponentiation (src/graph_magnus.rs:707-719).  Small instances come from the host builders;
    the large ones (BASELINE configs[4]: 200^3) from the engine's device generators, which give the same bytes (b200_lattice +
    b200_thin, checked against the host builders in tests/) in ~0.1 s instead of ~30 s."""
    from sparse_linear_algebra_tests_b200 import hostgen
    dims = [args.side * (1 if args.strong else world), args.side, args.side]
    if ctx is not None and dims[0] * dims[1] * dims[2] > 500_000:
        full = ctx.lattice(dims, True, args.bits)
        a, _ = ctx.thin(full, args.epn / 26.0, bytes([42] * 32))
        rp, ci, vv = a.download()
        return hostgen.HostCsr(a.rows, a.cols, rp, ci, vv)
    full = hostgen.lattice(dims, True, args.bits)
    density = args.epn / (full.nnz() / full.rows)
    return hostgen.thin(full, density, bytes([42] * 32))


def host_threads():
    """Every host core this process may run on -- NOT omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1, which made
    round 1's CPU arm single-threaded at N > 1.  The oracle's parallel regions take the thread count explicitly."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def source_sha():
    """Hash of the kernel sources: stamps profiles/r2_traffic.json so that a stale ncu figure is never reported."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "sparse_linear_algebra_tests_b200", "csrc")
    for name in sorted(os.listdir(d)):
        if name.endswith((".cu", ".cuh")):
            h.update(open(os.path.join(d, name), "rb").read())
    return h.hexdigest()[:16]


def algorithmic_bytes(nnz_a, nnz_b, nnz_c, rows_a, rows_b, vbytes):
    return (nnz_a + nnz_b + nnz_c) * (4 + vbytes) + (rows_a + rows_b + rows_a + 3) * 8


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (profiling recipe's clocks line)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------ CPU (reference arm / baseline)
def cpu_chain(a_h, iters=3, keep=False):
    """Reference protocol on the CPU restatement of CsrMatrix::matmul_par: per power 1 warm-up (kept as the next
    left operand) + `iters` timed multiplies with the result dropped.  Returns per-power seconds and products (and,
    with `keep`, every power -- the checker of the GPU chain)."""
    from oracle import oracle as O
    nt = host_threads()
    a = O.Csr(a_h.rows, a_h.cols, a_h.row_ptr, a_h.col_idx, a_h.values)
    p, secs, prods, kept = a, [], [], []
    for _k in range(2, MAX_POWER + 1):
        prods.append(int(O.row_products(p, a).sum()))
        nxt = O.matmul_par(p, a, nt)
        secs.append(O.time_matmul(p, a, True, nt, iters))
        p = nxt
        if keep:
            kept.append(p)
    return (secs, prods, nt, kept) if keep else (secs, prods, nt)


def run_reference(args):
    rank, _lr, world = dist_env()
    if rank != 0:
        return
    a_h = build_operand(args, max(1, args.gpus))                    # same job as the GPU arm at this N (the CPU does all rows)
    steps = max(1, min(args.steps, 3))                              # bounded: each step already holds 3 timed multiplies per power
    for _ in range(min(args.warmup, 1)):
        cpu_chain(a_h, 1)
    t_steps, prods = [], None
    for _ in range(steps):
        secs, prods, nt = cpu_chain(a_h, 3)
        t_steps.append(sum(secs))
    t = float(np.mean(t_steps))
    value = sum(prods) / t
    cfg = workload_config(args, max(1, args.gpus))
    cfg["cpu"] = "OpenMP restatement of the reference's rayon matmul_par (Rust toolchain absent: the reference itself cannot be built)"
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
           "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": f"u{args.bits}",
           "data": "synthetic", "config": cfg,
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": nt, "kind": "port",
                            "sample": f"full A^2..A^{MAX_POWER} chain, 3 timed multiplies per power, {steps} repetitions"},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from sparse_linear_algebra_tests_b200 import Context, hostgen

    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback")
    sim = args.as_rank is not None and world == 1                   # one rank of an N-rank job, alone (its steps need nobody else)
    if sim:
        rank, world = args.as_rank, max(1, args.gpus)
    multi = world > 1 and not sim                                   # a process group exists
    lead = rank == 0 or sim                                         # builds the operand, prints the line
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if multi:
        dist.init_process_group("nccl", device_id=dev)
    if args.workload == "rmat":
        if sim:
            raise SystemExit("bench.py: --as-rank covers the torus chain only")
        return run_rmat(args, torch, dist, dev, rank, local_rank, world)
    stream = torch.cuda.Stream(dev)                                 # the engine runs on this (non-default) torch stream,
    torch.cuda.set_stream(stream)                                   # so torch.cuda.Event on it brackets every engine kernel
    ctx = Context(local_rank, stream.cuda_stream)
    vbytes = args.bits // 8
    vdt = torch.int32 if args.bits == 32 else torch.int64          # bit containers for NCCL

    # ---- operand: rank 0 builds A, one NCCL broadcast replicates it (the only collective of the job)
    if lead:
        a_h = build_operand(args, world, ctx)
        meta = torch.tensor([a_h.rows, a_h.cols, a_h.nnz()], dtype=torch.int64, device=dev)
    else:
        a_h, meta = None, torch.zeros(3, dtype=torch.int64, device=dev)
    if multi:
        dist.broadcast(meta, 0)
    n_rows, n_cols, n_nnz = (int(x) for x in meta.tolist())
    if lead:
        d_rp = torch.from_numpy(a_h.row_ptr.view(np.int64)).to(dev)
        d_ci = torch.from_numpy(a_h.col_idx.view(np.int32)).to(dev)
        d_vv = torch.from_numpy(a_h.values.view(np.int32 if args.bits == 32 else np.int64)).to(dev)
    else:
        d_rp = torch.empty(n_rows + 1, dtype=torch.int64, device=dev)
        d_ci = torch.empty(n_nnz, dtype=torch.int32, device=dev)
        d_vv = torch.empty(n_nnz, dtype=vdt, device=dev)
    if multi:
        for t in (d_rp, d_ci, d_vv):
            dist.broadcast(t, 0)
    torch.cuda.synchronize(dev)
    A = ctx.from_device(n_rows, n_cols, n_nnz, d_rp.data_ptr(), d_ci.data_ptr(), d_vv.data_ptr(), args.bits)
    # ---- this rank's row block of the left operand, balanced by intermediate products of A x A
    if world > 1:
        cuts = ctx.shard_rows_by_products(A, A, world)
        r0, r1 = int(cuts[rank]), int(cuts[rank + 1])
        A_blk = ctx.row_block(A, r0, r1)
    else:
        r0, r1, A_blk = 0, n_rows, A

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    # ---- N > 1: the same rows through left multiplies (distributed.HaloPowerChain).  A rank's block of A^(k-1) is no power of
    # A, so `block x A` cannot use the engine's left-multiply kernel; A_k x A^(k-1) over the block plus a halo of neighbouring
    # rows (computed redundantly, nothing exchanged) can.  Rows [r0, r1) of its products are the rank's blocks, bit for bit
    # (checked below against the right multiplies); only the block's products count as work.
    halo = None
    if world > 1 and args.chain != "right":
        from sparse_linear_algebra_tests_b200.distributed import CudaEngine, HaloPowerChain
        a_all = a_h if a_h is not None else hostgen.HostCsr(
            n_rows, n_cols, d_rp.cpu().numpy().view(np.uint64), d_ci.cpu().numpy().view(np.uint32),
            d_vv.cpu().numpy().view(np.uint32 if args.bits == 32 else np.uint64))
        hc = HaloPowerChain(CudaEngine(ctx), a_all, r0, r1, MAX_POWER, a_dev=A)
        if args.chain == "halo" or hc.overhead <= 0.25:                  # (a graph without locality: the halo is everything)
            halo = hc
    use_halo = halo is not None and args.chain == "halo"

    def chain(collect_stats=False, left=None):
        """One step: every power of this rank.  `left`: through the halo chain (default: the mode chosen below)."""
        left = use_halo if left is None else left
        p, stats, keep = (A if left else A_blk), [], []
        for i in range(MAX_POWER - 1):
            x, y = (halo.left[i], p) if left else (p, A)
            if collect_stats:
                c, st = ctx.spgemm(x, y, True)
                stats.append(st.as_dict())
            else:
                c = ctx.spgemm(x, y)
            keep.append(c)
            p = c
        return keep, stats

    def barrier():
        if multi:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        # (shaped like a timed step -- flush, chain, wait, drop -- so that the stream-ordered allocator reaches its steady state
        #  here: at the 200^3 size a step that still grows the pool costs three steady ones)
        flush.fill_(1)
        powers, _ = chain()
        ctx.synchronize()
        del powers
    # one instrumented pass for per-multiply numbers (events inside the engine; not part of the timed steps); the products and
    # entries of the rank's own rows come from the right multiplies in every mode
    powers, st = chain(True, left=False)
    prods = [s["products"] for s in st]
    nnzs = [s["nnz_c"] for s in st]
    halo_note = None
    if halo is not None:
        def dev_bytes(ptr, nbytes):
            class _V:
                __cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}
            return torch.as_tensor(_V(), device=dev)
        hp, hst = chain(True, left=True)
        for k, (g, c) in zip(range(2, MAX_POWER + 1), zip(powers, hp)):   # bit for bit: row_ptr, col_idx, values of the block
            b = halo.block(c)
            ctx.synchronize()
            if b.nnz != g.nnz:
                raise SystemExit(f"bench.py: rank {rank}: A^{k} block through the halo chain has {b.nnz} entries, right multiply {g.nnz}")
            for pa, pb, nb in zip(g.device_ptrs(), b.device_ptrs(), ((g.rows + 1) * 8, g.nnz * 4, g.nnz * vbytes)):
                if nb and not torch.equal(dev_bytes(pa, nb), dev_bytes(pb, nb)):
                    raise SystemExit(f"bench.py: rank {rank}: A^{k} block through the halo chain differs from the right multiply")
            del b
        del hp
        if args.chain == "auto":                                          # whichever is faster on this rank, measured
            t_mode = []
            for left in (False, True):
                best_ms = 1e30
                for _ in range(4):
                    flush.fill_(1)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    pw, _ = chain(left=left)
                    e1.record(stream)
                    e1.synchronize()
                    best_ms = min(best_ms, e0.elapsed_time(e1))
                    del pw
                t_mode.append(best_ms)
            use_halo = t_mode[1] < t_mode[0]
            halo_note = {"right_ms": t_mode[0], "halo_ms": t_mode[1]}
        if use_halo:
            for _ in range(2):
                powers2, _ = chain()
                ctx.synchronize()
                del powers2
            _, st = chain(True)
            for s_, p_, n_ in zip(st, prods, nnzs):                        # the halo rows are redundant work: not counted
                s_["products_with_halo"], s_["products"] = s_["products"], p_
                s_["nnz_with_halo"], s_["nnz_c"] = s_["nnz_c"], n_
    del powers

    ctx.set_timing(False)
    sampler = ClockSampler(local_rank)
    if lead:
        sampler.start()
    barrier()
    launches0 = ctx.kernel_launches()
    step_ms = []
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1)                                               # evict L2 between steps (untimed)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        powers, _ = chain()
        e1.record(stream)
        e1.synchronize()
        step_ms.append(e0.elapsed_time(e1))
        del powers
    barrier()
    wall = time.perf_counter() - wall0
    launches = ctx.kernel_launches() - launches0
    clocks = sampler.stop() if lead else None
    ctx.set_timing(True)

    my_ms = float(np.mean(step_ms))
    # per-rank step time and product count, so that jitter (spread of one rank's steps) and imbalance (spread over ranks)
    # can be told apart in the line
    mine = torch.tensor([my_ms, float(np.min(step_ms)), float(np.max(step_ms)), float(sum(prods))], dtype=torch.float64, device=dev)
    if multi:
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
    else:
        allr = [mine]
    per_rank = [{"rank": i, "ms_mean": float(t[0]), "ms_min": float(t[1]), "ms_max": float(t[2]), "products": int(t[3])} for i, t in enumerate(allr)]
    t_ms = torch.tensor([my_ms], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(sum(prods)), float(launches)], dtype=torch.float64, device=dev)
    if multi:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_per_step = float(t_ms.item())
    total_products, total_launches = float(tot[0].item()), int(tot[1].item())
    value = total_products / (ms_per_step * 1e-3)
    chain_cfg = {}
    if world > 1:
        n_halo = torch.tensor([1.0 if use_halo else 0.0], dtype=torch.float64, device=dev)
        if multi:
            dist.all_reduce(n_halo, op=dist.ReduceOp.SUM)
        chain_cfg["chain"] = {
            "mode": args.chain, "ranks_on_halo_chain": int(n_halo.item()),
            "halo": "A_k x A^(k-1) over the rank's rows plus a halo of neighbouring rows computed redundantly (no communication); "
                    "rows [r0, r1) of every power compared bit for bit with the right multiplies before timing; only the block's products count",
            "right": "block(A^(k-1)) x A",
            "rank0_redundant_work_estimate": halo.overhead if halo is not None else None, "rank0_warmup_ms": halo_note}
        if use_halo:
            chain_cfg["left"] = "A restricted to the rows a rank needs at power k (block + halo), whole-size handle"
            chain_cfg["right"] = "A^(k-1), the rank's rows (block + halo) resident"
        if sim:
            chain_cfg["simulated"] = f"rank {rank} of {world} alone on one GPU: value is this rank's share only"

    # ---- per-multiply detail (rank-local, instrumented pass repeated for a best-of-5)
    best = [dict(s) for s in st]
    for _ in range(4):
        flush.fill_(1)
        powers, st2 = chain(True)
        for b, s in zip(best, st2):
            if s["ms_total"] < b["ms_total"]:
                b.update(s)
        del powers
    if use_halo:
        for b, p_, n_ in zip(best, prods, nnzs):
            b["products"], b["nnz_c"] = p_, n_
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    top = max(best, key=lambda s: s["bytes_algorithmic"])
    achieved = top["bytes_algorithmic"] / (top["ms_total"] * 1e-3) / 1e9
    # DRAM bytes of that multiply from the committed ncu capture (dram__bytes_read.sum + dram__bytes_write.sum over its
    # kernels); reported only when the capture was taken on this workload AND on these kernel sources (hash stamp)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if world == 1 and args.side == 30 and args.bits == 64 and MAX_POWER == 7 and os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("source_sha") == source_sha():
            traffic = float(tj["traffic"])
    kernels = {1: "k_fz_prepass + k_fz_numeric (pre-pass; fused numeric + placement, C written once)",
               3: "pre-pass, row-per-warp count kernels, row_ptr scan, row-per-warp numeric kernels (C written once)",
               4: "k_rw_fused, ONE cooperative launch: count phase, placement, numeric phase (C written once, no host wait)",
               5: "k_dn, ONE launch: dense window accumulators per row, look-back placement over rows",
               6: "k_lm, ONE cooperative launch (operands commute: evaluated as A x A^(k-1), rows of A^(k-1) streamed as sorted lists): "
                  "count phase, placement, numeric phase (C written once, no host wait)"}.get(
                   top.get("pipeline"), "binned pipeline (pre-pass, per-bin numeric, row_ptr scan, compaction)")
    roofline = {"bound": "hbm", "kernel": f"all kernels of the largest multiply A^{MAX_POWER} = A^{MAX_POWER - 1} x A: " + kernels,
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                "algorithmic_bytes": top["bytes_algorithmic"], "ms": top["ms_total"], "traffic": traffic,
                "numeric_only_frac": top["bytes_algorithmic"] / (top["ms_numeric"] * 1e-3) / 1e9 / peak if top["ms_numeric"] else None}
    per_power = [{"power": k, "products": s["products"], "nnz": s["nnz_c"], "ms": s["ms_total"],
                  "gbs": s["bytes_algorithmic"] / (s["ms_total"] * 1e-3) / 1e9, "launches": s["kernel_launches"]}
                 for k, s in zip(range(2, MAX_POWER + 1), best)]

    # ---- end to end through the public API with host buffers (pinned H2D of A, pinned D2H of every power)
    e2e = None
    if args.no_e2e:
        if lead:
            cfg = workload_config(args, world)
            cfg.update({"nodes": n_rows, "nnz_A": n_nnz, "products_per_step": int(total_products), "nnz_per_power_rank0": nnzs,
                        "parallelism": f"row-sharded x{world}" if world > 1 else "1 GPU"})
            cfg.update(chain_cfg)
            print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                              "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None,
                              "dtype": f"u{args.bits}", "data": "synthetic", "config": cfg, "per_power": per_power, "per_rank": per_rank, "gpu_launches": total_launches,
                              "wall_s_timed_region": wall, "clocks": clocks, "e2e": None, "roofline": roofline}), flush=True)
        if multi:
            dist.destroy_process_group()
        return
    a_loc = hostgen.HostCsr(n_rows, n_cols, d_rp.cpu().numpy().view(np.uint64), d_ci.cpu().numpy().view(np.uint32),
                            d_vv.cpu().numpy().view(np.uint32 if args.bits == 32 else np.uint64))
    blk = a_loc.row_block(r0, r1)
    pin = lambda arr: torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8)).pin_memory()
    h_in = [pin(x) for x in (a_loc.row_ptr, a_loc.col_idx, a_loc.values, blk.row_ptr, blk.col_idx, blk.values)]
    h_out = [(torch.empty((r1 - r0 + 1) * 8, dtype=torch.uint8).pin_memory(), torch.empty(max(n, 1) * 4, dtype=torch.uint8).pin_memory(),
              torch.empty(max(n, 1) * vbytes, dtype=torch.uint8).pin_memory()) for n in nnzs]
    vnp = np.uint32 if args.bits == 32 else np.uint64
    h2d = sum(int(t.numel()) for t in (h_in[:3] if world == 1 else h_in))
    d2h = sum(int(a.numel() + b.numel() + c.numel()) for a, b, c in h_out)

    def e2e_step():
        Af = ctx.upload(n_rows, n_cols, h_in[0].numpy().view(np.uint64), h_in[1].numpy().view(np.uint32), h_in[2].numpy().view(vnp))
        p = Af if world == 1 else ctx.upload(r1 - r0, n_cols, h_in[3].numpy().view(np.uint64), h_in[4].numpy().view(np.uint32), h_in[5].numpy().view(vnp))
        keep = []
        for i in range(MAX_POWER - 1):
            c = ctx.spgemm(p, Af)
            c.download_async_into(h_out[i][0].data_ptr(), h_out[i][1].data_ptr(), h_out[i][2].data_ptr())
            keep.append(c)
            p = c
        ctx.synchronize()
        return keep

    for _ in range(2):
        e2e_step()
    barrier()
    e_t = []
    for _ in range(max(3, min(args.steps, 10))):
        flush.fill_(1)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        keep = e2e_step()
        e_t.append(time.perf_counter() - t0)
        del keep
    e_ms = torch.tensor([float(np.mean(e_t)) * 1e3], dtype=torch.float64, device=dev)
    if multi:
        dist.all_reduce(e_ms, op=dist.ReduceOp.MAX)
    # the last power read back must equal the resident result (cheap end-to-end sanity: nnz via row_ptr)
    last_rp = h_out[-1][0].numpy().view(np.uint64)
    assert int(last_rp[-1]) == nnzs[-1], "end-to-end read-back disagrees with the device result"
    e2e = {"value": total_products / (float(e_ms.item()) * 1e-3), "unit": UNIT, "ms_per_step": float(e_ms.item()),
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "timing": "wall clock, synchronize on both sides, max over ranks"}

    if not lead:
        if multi:
            dist.destroy_process_group()
        return
    cfg = workload_config(args, world)
    cfg.update({"nodes": n_rows, "nnz_A": n_nnz, "products_per_step": int(total_products), "nnz_per_power_rank0": nnzs,
                "parallelism": f"row-sharded x{world}" if world > 1 else "1 GPU"})
    cfg.update(chain_cfg)
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
           "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": f"u{args.bits}",
           "data": "synthetic", "config": cfg, "per_power": per_power, "per_rank": per_rank, "gpu_launches": total_launches, "wall_s_timed_region": wall,
           "clocks": clocks, "e2e": e2e, "roofline": roofline}
    if world == 1 and not args.no_cpu_baseline:
        secs, cprods, nt, cpu_powers = cpu_chain(a_loc, 3, keep=True)
        # the GPU chain of this run against the CPU chain it is timed beside: every power, every array, bit for bit
        gpu_powers, _ = chain()
        for k, (g, c) in zip(range(2, MAX_POWER + 1), zip(gpu_powers, cpu_powers)):
            rp, ci, vv = g.download()
            if not (np.array_equal(rp, c.row_ptr) and np.array_equal(ci, c.col_idx) and np.array_equal(vv, c.values)):
                raise SystemExit(f"bench.py: A^{k} from the GPU differs from the CPU restatement -- the timed result is wrong")
        del gpu_powers
        out["parity"] = f"A^2..A^{MAX_POWER} of this run bit-identical to the CPU chain (row_ptr, col_idx, values)"
        out["cpu_baseline"] = {"value": sum(cprods) / sum(secs), "unit": UNIT, "cores": nt, "kind": "port",
                               "sample": f"full A^2..A^{MAX_POWER} chain, reference protocol (1 warm-up + 3 timed multiplies per power)",
                               "ms_per_power": [s * 1e3 for s in secs], "ms_per_step": sum(secs) * 1e3}
    print(json.dumps(out), flush=True)
    if multi:
        dist.destroy_process_group()


def run_rmat(args, torch, dist, dev, rank, local_rank, world):
    """BASELINE configs[3]: rank 0 builds the graph on its GPU (b200_rmat: generator + radix-sort assembly), b200_comm_broadcast_csr
    replicates it, every rank multiplies its product-balanced row block by the replicated operand.  Strong scaling."""
    from sparse_linear_algebra_tests_b200 import Context
    from sparse_linear_algebra_tests_b200.distributed import make_comm
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    ctx = Context(local_rank, stream.cuda_stream)
    t0 = time.perf_counter()
    A = ctx.rmat(args.scale, args.ef, args.abc[0], args.abc[1], args.abc[2], 42, args.bits) if rank == 0 else None
    build_s = time.perf_counter() - t0
    comm = None
    if world > 1:
        comm = make_comm(ctx, rank, world)
        A = comm.broadcast(A, 0)
        cuts = ctx.shard_rows_by_products(A, A, world)
        blk = ctx.row_block(A, int(cuts[rank]), int(cuts[rank + 1]))
    else:
        blk = A
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        c = ctx.spgemm(blk, A)
        ctx.synchronize()
        del c
    c, st = ctx.spgemm(blk, A, True)
    st = st.as_dict()
    del c
    ctx.set_timing(False)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    launches0 = ctx.kernel_launches()
    step_ms = []
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        c = ctx.spgemm(blk, A)
        e1.record(stream)
        e1.synchronize()
        step_ms.append(e0.elapsed_time(e1))
        del c
    barrier()
    wall = time.perf_counter() - wall0
    launches = ctx.kernel_launches() - launches0
    clocks = sampler.stop() if rank == 0 else None
    my_ms = float(np.mean(step_ms))
    mine = torch.tensor([my_ms, float(np.min(step_ms)), float(np.max(step_ms)), float(st["products"]), float(st["nnz_c"]), float(launches)], dtype=torch.float64, device=dev)
    allr = [torch.zeros_like(mine) for _ in range(world)]
    if world > 1:
        dist.all_gather(allr, mine)
    else:
        allr = [mine]
    if rank == 0:
        per_rank = [{"rank": i, "ms_mean": float(t[0]), "ms_min": float(t[1]), "ms_max": float(t[2]), "products": int(t[3]), "nnz": int(t[4])} for i, t in enumerate(allr)]
        ms_per_step = max(p["ms_mean"] for p in per_rank)
        total_products = sum(p["products"] for p in per_rank)
        total_nnz = sum(p["nnz"] for p in per_rank)
        vb = args.bits // 8
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        peak, peak_src = (float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)") if os.path.exists(peaks_path) else (6650.0, "fallback (B200_PROFILING.md)")
        alg = algorithmic_bytes(A.nnz, A.nnz, total_nnz, A.rows, A.rows, vb)                 # whole job: A read once, A (right) once, C written once
        achieved = alg / (ms_per_step * 1e-3) / 1e9
        cfg = workload_config(args, world)
        cfg.update({"nodes": A.rows, "nnz_A": A.nnz, "products_per_step": int(total_products), "nnz_C": int(total_nnz), "max_row_products": st["max_row_products"],
                    "operand_build_s_rank0": build_s, "parallelism": f"row-sharded x{world}" if world > 1 else "1 GPU"})
        out = {"metric": METRIC, "value": total_products / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
               "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": f"u{args.bits}", "data": "synthetic",
               "config": cfg, "per_rank": per_rank, "gpu_launches": int(sum(float(t[5]) for t in allr)), "wall_s_timed_region": wall, "clocks": clocks, "e2e": None,
               "roofline": {"bound": "hbm", "kernel": "all kernels of the multiply, all GPUs (aggregate algorithmic bytes over the max-over-ranks time)", "achieved": achieved,
                            "peak": peak * world, "unit": "GB/s", "frac": achieved / (peak * world), "peak_source": peak_src + f" x {world} GPUs", "algorithmic_bytes": alg,
                            "ms": ms_per_step, "traffic": None}}
        print(json.dumps(out), flush=True)
    if comm is not None:
        comm.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    global MAX_POWER
    args = parse_args()
    MAX_POWER = args.max_power
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
