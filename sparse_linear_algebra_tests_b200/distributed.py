"""Multi-GPU orchestration of the SpGEMM path: natural row sharding, one process per GPU.

The reference has no distributed code; the path shards by rows of the LEFT operand
because row i of C depends only on row i of A and on B (src/graph_csr.rs:433-446):

  * B (the original A of a power chain) is replicated once with a `torch.distributed` broadcast
    (NCCL over NVLink on GPUs, gloo in the CPU tests);
  * the left operand is cut into contiguous row blocks balanced by intermediate-product count,
    not row count (`product_balanced_cuts`);
  * in repeated exponentiation every rank keeps its row block of A^(k-1) resident and multiplies it
    by the replicated A: no communication per step (`ShardedPowerChain`);
  * `HaloPowerChain` is the same chain through left multiplies (the engine's fastest kernel on a power chain): the rank
    computes its block plus a shrinking halo of neighbouring rows redundantly instead of exchanging them;
  * `allgather_csr` optionally assembles the full C from the row blocks (variable sizes).

Everything here is engine-agnostic plumbing: the multiply itself is `engine.spgemm`, which in the
product is the CUDA engine (`CudaEngine`, no CPU fallback).  The gloo tests inject a CPU checker.
"""
from __future__ import annotations

from typing import Protocol

import numpy as np

from . import hostgen


HEAVY_FROM = 8192   # rows with more intermediate products take the heavy-row kernels (b200_hash_cap(7), csrc/common.cuh)


def row_cost(row_products: np.ndarray) -> np.ndarray:
    """shard_cost of csrc/api.cu: products + 1 (empty rows spread too), heavy rows 2.5 x (their kernels' products/s)."""
    p = np.asarray(row_products, dtype=np.uint64)
    return p + np.uint64(1) + np.where(p > np.uint64(HEAVY_FROM), p + p // np.uint64(2), np.uint64(0)).astype(np.uint64)


def product_balanced_cuts(row_products: np.ndarray, nparts: int) -> np.ndarray:
    """cuts[0..nparts]: part k = rows [cuts[k], cuts[k+1]) holds ~1/nparts of sum(cost(P_i)).

    Same rule as b200_shard_rows_by_products (csrc/api.cu): prefix of the row costs, cut k at the first prefix
    >= k * total / nparts."""
    c = row_cost(row_products)
    pre = np.zeros(c.shape[0] + 1, dtype=np.uint64)
    np.cumsum(c, out=pre[1:])
    total = int(pre[-1])
    cuts = np.zeros(nparts + 1, dtype=np.uint64)
    cuts[nparts] = c.shape[0]
    for k in range(1, nparts):
        lo = int(np.searchsorted(pre, np.uint64(total * k // nparts), side="left"))
        cuts[k] = max(min(lo, c.shape[0]), int(cuts[k - 1]))
    return cuts


class Engine(Protocol):
    def upload(self, h: hostgen.HostCsr): ...
    def download(self, m) -> hostgen.HostCsr: ...
    def row_products(self, a, b) -> np.ndarray: ...
    def row_block(self, a, r0: int, r1: int): ...
    def spgemm(self, a, b): ...


class CudaEngine:
    """The product engine: libb200spgemm.so through `_native.Context` (raises without a B200)."""

    def __init__(self, ctx):
        self.ctx = ctx

    def upload(self, h):
        return self.ctx.upload(h.rows, h.cols, h.row_ptr, h.col_idx, h.values)

    def download(self, m):
        rp, ci, vv = m.download()
        return hostgen.HostCsr(m.rows, m.cols, rp, ci, vv)

    def row_products(self, a, b):
        return self.ctx.row_products(a, b)

    def row_block(self, a, r0, r1):
        return self.ctx.row_block(a, r0, r1)

    def spgemm(self, a, b):
        return self.ctx.spgemm(a, b)


def broadcast_host_csr(h: hostgen.HostCsr | None, src: int = 0, group=None, device="cpu") -> hostgen.HostCsr:
    """Replicate a CSR from rank `src` (three tensors + a 4-word header); returns the host copy on every rank."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group)
    meta = torch.zeros(4, dtype=torch.int64, device=device)
    if rank == src:
        meta = torch.tensor([h.rows, h.cols, h.nnz(), h.val_bits], dtype=torch.int64, device=device)
    dist.broadcast(meta, src, group=group)
    rows, cols, nnz, bits = (int(x) for x in meta.tolist())
    vnp = np.int32 if bits == 32 else np.int64
    if rank == src:
        t_rp = torch.from_numpy(h.row_ptr.view(np.int64).copy()).to(device)
        t_ci = torch.from_numpy(h.col_idx.view(np.int32).copy()).to(device)
        t_vv = torch.from_numpy(h.values.view(vnp).copy()).to(device)
    else:
        t_rp = torch.empty(rows + 1, dtype=torch.int64, device=device)
        t_ci = torch.empty(nnz, dtype=torch.int32, device=device)
        t_vv = torch.empty(nnz, dtype=torch.int32 if bits == 32 else torch.int64, device=device)
    for t in (t_rp, t_ci, t_vv):
        dist.broadcast(t, src, group=group)
    return hostgen.HostCsr(rows, cols, t_rp.cpu().numpy().view(np.uint64), t_ci.cpu().numpy().view(np.uint32),
                           t_vv.cpu().numpy().view(np.uint32 if bits == 32 else np.uint64))


class ShardedPowerChain:
    """A^k = A^(k-1) x A with the left operand row-sharded over the ranks and A replicated.

    Rank r owns rows [cuts[r], cuts[r+1]) of every power; `step()` needs no communication."""

    def __init__(self, engine: Engine, a_host: hostgen.HostCsr, rank: int, world: int):
        self.engine, self.rank, self.world = engine, rank, world
        self.a = engine.upload(a_host)
        self.row_products = np.asarray(engine.row_products(self.a, self.a), dtype=np.uint64)
        self.cuts = product_balanced_cuts(self.row_products, world)
        self.r0, self.r1 = int(self.cuts[rank]), int(self.cuts[rank + 1])
        self.block = engine.row_block(self.a, self.r0, self.r1) if world > 1 else self.a
        self.power = 1

    def step(self):
        self.block = self.engine.spgemm(self.block, self.a)
        self.power += 1
        return self.block

    def local_products(self) -> int:
        return int(np.asarray(self.engine.row_products(self.block, self.a), dtype=np.uint64).sum())


def halo_row_sets(a: hostgen.HostCsr, r0: int, r1: int, max_power: int) -> list:
    """need[k] (k = 1 .. max_power, need[0] unused): boolean masks of the rows of A^k a rank must hold so that rows [r0, r1)
    of A^max_power follow from LEFT multiplies alone.  Row i of A^k = A x A^(k-1) reads the rows of A^(k-1) that A[i, :] points
    at (src/graph_csr.rs:433-446 with the operands swapped; powers of one matrix commute), so need[max_power] = [r0, r1) and
    need[k-1] = need[k] | columns(A[need[k], :]).  On a banded or lattice operand the sets grow by one halo per power; on a
    graph without locality they reach every row after a step or two (`halo_overhead` tells the two apart)."""
    lens = np.diff(a.row_ptr.astype(np.int64))
    need = [None] * (max_power + 1)
    m = np.zeros(a.rows, dtype=bool)
    m[r0:r1] = True
    need[max_power] = m
    for k in range(max_power, 1, -1):
        nxt = need[k].copy()
        nxt[a.col_idx[np.repeat(need[k], lens)]] = True
        need[k - 1] = nxt
    return need


def restrict_rows(a: hostgen.HostCsr, mask: np.ndarray) -> hostgen.HostCsr:
    """A with the rows outside `mask` emptied (same shape: the column space and the row numbering stay global)."""
    lens = np.diff(a.row_ptr.astype(np.int64))
    keep = np.repeat(mask, lens)
    rp = np.zeros(a.rows + 1, dtype=np.uint64)
    np.cumsum(np.where(mask, lens, 0), out=rp[1:])
    return hostgen.HostCsr(a.rows, a.cols, rp, a.col_idx[keep].copy(), a.values[keep].copy())


def halo_overhead(need: list, r0: int, r1: int, growth: float = 2.0) -> float:
    """Share of redundant work of a halo chain: rows of need[k] beyond the block, weighted by growth^k (the products of a
    power chain roughly double per power).  0 = none, 1 = as much again as the useful work."""
    own = max(r1 - r0, 1)
    num = den = 0.0
    for k in range(2, len(need)):
        w = growth ** k
        num += w * (int(need[k].sum()) - own) / own
        den += w
    return num / den if den else 0.0


class HaloPowerChain:
    """Rows [r0, r1) of A^2 .. A^max_power on one rank, without communication, through LEFT multiplies.

    `ShardedPowerChain` multiplies the rank's row block of A^(k-1) by the replicated A: the long rows are on the left and the
    engine cannot use its left-multiply kernel (a row block is no power of A, so the operands do not commute).  Here the
    rank evaluates A_k x A^(k-1) instead, where A_k is A with the rows outside need[k] emptied (`halo_row_sets`): whole-size
    handles whose rows outside need[k] are empty, so rows and columns keep their global numbers and rows [r0, r1) of every
    product are bit for bit the rank's block of A^k.  The halo rows are computed redundantly by the neighbours instead of
    being exchanged: communication-avoiding, and worth it while `overhead` stays small (lattices, banded matrices)."""

    def __init__(self, engine: Engine, a_host: hostgen.HostCsr, r0: int, r1: int, max_power: int, a_dev=None):
        self.engine, self.r0, self.r1, self.max_power = engine, r0, r1, max_power
        self.need = halo_row_sets(a_host, r0, r1, max_power)
        self.overhead = halo_overhead(self.need, r0, r1)
        self.a = a_dev if a_dev is not None else engine.upload(a_host)
        self.left = [engine.upload(restrict_rows(a_host, self.need[k])) for k in range(2, max_power + 1)]

    def run(self) -> list:
        """[A^2, .., A^max_power] as whole-size handles; rows outside need[k] are empty."""
        p, out = self.a, []
        for ak in self.left:
            p = self.engine.spgemm(ak, p)
            out.append(p)
        return out

    def block(self, c):
        """The rank's row block of a product of `run`, as its own handle (local row numbers, global columns)."""
        return self.engine.row_block(c, self.r0, self.r1)


def allgather_csr(block: hostgen.HostCsr, group=None, device="cpu") -> hostgen.HostCsr:
    """Assemble the full matrix from per-rank row blocks (rank order = row order).  Blocks have different
    sizes: lengths are gathered first, payloads are padded to the largest block, row_ptr is re-based."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    bits = block.val_bits
    sizes = torch.tensor([block.rows, block.nnz()], dtype=torch.int64, device=device)
    all_sizes = [torch.zeros(2, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    rows_l = [int(s[0]) for s in all_sizes]
    nnz_l = [int(s[1]) for s in all_sizes]
    mr, mn = max(rows_l) + 1, max(max(nnz_l), 1)
    vnp, vt = (np.int32, torch.int32) if bits == 32 else (np.int64, torch.int64)

    def padded(arr, n, dt):
        t = torch.zeros(n, dtype=dt, device=device)
        t[:arr.shape[0]] = torch.from_numpy(arr.copy()).to(device)
        return t

    parts = []
    for arr, n, dt in ((block.row_ptr.view(np.int64), mr, torch.int64), (block.col_idx.view(np.int32), mn, torch.int32),
                       (block.values.view(vnp), mn, vt)):
        mine = padded(arr, n, dt)
        out = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(out, mine, group=group)
        parts.append([o.cpu().numpy() for o in out])
    rp = [np.zeros(1, dtype=np.uint64)]
    base = 0
    for r in range(world):
        local = parts[0][r][:rows_l[r] + 1].view(np.uint64)
        rp.append(local[1:] + np.uint64(base))
        base += nnz_l[r]
    col = np.concatenate([parts[1][r][:nnz_l[r]].view(np.uint32) for r in range(world)])
    val = np.concatenate([parts[2][r][:nnz_l[r]].view(np.uint32 if bits == 32 else np.uint64) for r in range(world)])
    return hostgen.HostCsr(sum(rows_l), block.cols, np.concatenate(rp).astype(np.uint64), col, val)


# ------------------------------------------------------------------------------------------ device collectives (NCCL below the C ABI)
def make_comm(ctx, rank: int, world: int, group=None):
    """A `Comm` (b200_comm) for this rank: rank 0 creates the NCCL id, torch.distributed's object broadcast (any backend)
    carries its 128 bytes -- the only thing torch does for the data path."""
    import torch.distributed as dist
    from ._native import Comm
    box = [Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    return Comm(ctx, world, rank, box[0])


class ShardedSquaring:
    """power_until_stable (src/graph_csr.rs:561-575) over N GPUs: cur <- cur x cur until the pattern stops changing.

    Both operands grow, so unlike the power chain every step ends with an exchange: rank r multiplies its
    product-balanced row block of `cur` by the whole `cur` and `Comm.allgather` (device buffers, NCCL) assembles the next
    `cur` on every rank.  The stabilisation test (equal nnz, row_ptr and col_idx, :567-570) runs on every rank's own copy."""

    def __init__(self, ctx, comm, a_dev):
        self.ctx, self.comm, self.cur = ctx, comm, a_dev

    def step(self):
        cuts = self.ctx.shard_rows_by_products(self.cur, self.cur, self.comm.size)
        blk = self.ctx.row_block(self.cur, int(cuts[self.comm.rank]), int(cuts[self.comm.rank + 1]))
        nxt = self.comm.allgather(self.ctx.spgemm(blk, self.cur))
        stable = nxt.nnz == self.cur.nnz and self.ctx.same_pattern(nxt, self.cur)
        self.cur = nxt
        return stable

    def run(self, max_steps: int = 64):
        for k in range(1, max_steps + 1):
            if self.step():
                return self.cur, k
        raise RuntimeError("power_until_stable did not stabilise")
