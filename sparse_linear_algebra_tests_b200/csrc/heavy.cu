// heavy.cu -- the heaviest rows of C: column-space chunks, B-row segments staged by TMA bulk copies, a dense accumulator
// per chunk.  This is MAGNUS' fine / coarse level (the reference reaches it through
// magnus_spgemm_parallel, /root/reference/src/graph_magnus.rs:225-232) cut for a GPU:
//
//   * the column space is cut into chunks of W columns (W from shared memory: 32 Ki columns with 32-bit sums, 16 Ki with
//     64-bit ones); a chunk's accumulator acc[W] and its bitmap live in shared memory, so a product costs one
//     shared-memory atomic wherever its column falls, and a chunk is emitted by walking its bitmap: ascending columns,
//     no sort, no hash table;
//   * B's rows are sorted, so the part of B[k,:] inside a chunk is one contiguous segment found by binary search (the
//     row's first / last column from the span array settle most rows without a search);
//   * long segments (>= HV_LONG_MIN entries: the hub rows of a power-law graph) are queued and streamed through a
//     double-buffered shared-memory stage by 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx): the copy engine
//     fetches piece i+1 while the CTA accumulates piece i out of shared memory, coalesced and conflict-free;
//     short segments go through a balanced product enumeration straight from L2;
//   * coarse level: a row is cut into work units of consecutive chunks -- several CTAs share the heaviest rows -- by a
//     planning kernel; a count pass (bitmaps only, count chunks as wide as 1 Mi columns) gives every (row, chunk) its
//     length, an in-place scan turns the lengths into offsets inside the row and the numeric pass writes each chunk at
//     its final place.
// Rows take this path when their intermediate products average >= HV_MIN_PER_CHUNK per chunk (below that the per-chunk
// overhead loses to the global-memory table of k_num_heavy); everything else about the multiply (pre-pass, placement,
// report) is api.cu's.
#include "engine.cuh"
#include "devutil.cuh"

#ifndef HV_THREADS
#define HV_THREADS 1024
#endif
#ifndef HV_CTAS
#define HV_CTAS 1                 // CTAs per SM the shared-memory budget is cut for
#endif
#ifndef HV_STAGE_ELEMS
#define HV_STAGE_ELEMS 2048u      // entries per staged piece: 8 KiB of columns (+ 16 KiB of u64 values)
#endif
#define HV_NSTAGE 2
#ifndef HV_LONG_MIN
#define HV_LONG_MIN 256u          // segments at least this long are staged through shared memory by bulk copies
#endif
#ifndef HV_TINY_MAX
#define HV_TINY_MAX 0u            // segments at most this long are walked by the thread that found them
#endif
#define HV_QCAP 256               // long segments queued per tile of HV_THREADS A entries (the rest take the short path)
#define HV_MAX_SW 64              // numeric chunks per count chunk, at most

struct HvCtl { u32 n_count_units, n_num_units, t_count, t_num; };
struct HvDev {
    const u32 *list; const u32 *list_count;       // the pre-pass's heavy list and its length (device)
    const u64 *prod;                              // intermediate products per row (pre-pass)
    u64 pmin; u32 cap_li;                         // rows taken: list index < cap_li and prod >= pmin
    u64 psplit;                                   // products per numeric work unit the planner aims at (coarse split)
    u32 ncols, W, nchunks, sw, nchunks_c;         // count chunk = sw numeric chunks
    u32 *cnt;                                     // [cap_li][nchunks]: length, then offset, of every chunk inside its row
    u64 *units_c, *units_n;                       // work units: li << 32 | first chunk << 16 | chunks
    HvCtl *ctl;
    const uint4 *bspan;                           // B rows: {len, first column, last column, -}
    u64 nnzB;
    u32 *nnz_row;
};

__device__ __forceinline__ bool hv_take(const HvDev &h, u32 li, u32 row) { return li < h.cap_li && h.prod[row] >= h.pmin; }

// ---- mbarrier / bulk-copy primitives (PTX ISA: mbarrier, cp.async.bulk)
__device__ __forceinline__ void mbar_init(u32 bar, u32 count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(u32 bar, u32 bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(u32 bar, u32 parity) {
    u32 ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(u32 dst, const void *src, u32 bytes, u32 bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

static __device__ __forceinline__ u32 hv_block_scan(u32 v, u32 *s_warp /* 33 */, u32 &total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    u32 incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const u32 t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
    if (lane == 31) s_warp[w] = incl;
    __syncthreads();
    if (w == 0) {
        u32 x = lane < nw ? s_warp[lane] : 0u, xi = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const u32 t = __shfl_up_sync(0xFFFFFFFFu, xi, d); if (lane >= d) xi += t; }
        s_warp[lane] = xi - x;
        if (lane == 31) s_warp[32] = xi;
    }
    __syncthreads();
    total = s_warp[32];
    const u32 r = s_warp[w] + incl - v;
    __syncthreads();
    return r;
}

// first index in col[0..n) whose column is >= key.  16-ary: every step reads up to 15 pivots with independent loads, so the
// chain of dependent L2 round trips is log16(n) long instead of log2(n) -- the search is latency, not bandwidth (a row of
// <= 16 entries takes one step, the longest hub rows four).
__device__ __forceinline__ u32 hv_lower_bound(const u32 *__restrict__ col, u32 n, u32 key) {
    u32 base = 0, len = n;
    while (len > 16) {
        const u32 step = (len + 15) >> 4;
        u32 cnt = 0;
#pragma unroll
        for (u32 j = 1; j < 16; j++) { const u32 i = j * step; if (i < len) cnt += col[base + i] < key ? 1u : 0u; }
        base += cnt * step;
        len = min(step, n - base);
    }
    u32 cnt = 0;
#pragma unroll
    for (u32 j = 0; j < 16; j++) if (j < len) cnt += col[base + j] < key ? 1u : 0u;
    return base + cnt;
}

// ---- planning: one thread per row of the heavy list
__global__ void __launch_bounds__(256) k_hv_plan(HvDev h) {
    const u32 count = min(*h.list_count, h.cap_li);
    for (u32 li = blockIdx.x * blockDim.x + threadIdx.x; li < count; li += gridDim.x * blockDim.x) {
        const u32 row = h.list[li];
        if (!hv_take(h, li, row)) continue;
        const u64 P = h.prod[row];
        {   // count units: one per count chunk (bitmaps only: cheap, and the heaviest rows are still shared by several CTAs)
            const u32 base = atomicAdd(&h.ctl->n_count_units, h.nchunks_c);
            for (u32 cc = 0; cc < h.nchunks_c; cc++) h.units_c[base + cc] = ((u64)li << 32) | ((u64)cc << 16) | 1ull;
        }
        // numeric units: G groups of consecutive chunks, ~psplit products each
        u32 G = (u32)min((u64)h.nchunks, max((u64)1, P / h.psplit));
        const u32 per = (h.nchunks + G - 1) / G;
        G = (h.nchunks + per - 1) / per;
        const u32 base = atomicAdd(&h.ctl->n_num_units, G);
        for (u32 g = 0; g < G; g++) {
            const u32 c0 = g * per, n = min(per, h.nchunks - c0);
            h.units_n[base + g] = ((u64)li << 32) | ((u64)c0 << 16) | (u64)n;
        }
    }
}

// ---- lengths -> offsets inside the row; the row's length goes to nnz_row
__global__ void __launch_bounds__(256) k_hv_scan(HvDev h) {
    const u32 count = min(*h.list_count, h.cap_li);
    const int lane = threadIdx.x & 31;
    const u32 wpb = blockDim.x >> 5;
    for (u32 li = blockIdx.x * wpb + (threadIdx.x >> 5); li < count; li += gridDim.x * wpb) {
        const u32 row = h.list[li];
        if (!hv_take(h, li, row)) continue;
        u32 *c = h.cnt + (u64)li * h.nchunks;
        u32 run = 0;
        for (u32 b = 0; b < h.nchunks; b += 32) {
            const u32 v = b + lane < h.nchunks ? c[b + lane] : 0u;
            u32 incl = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const u32 t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
            if (b + lane < h.nchunks) c[b + lane] = run + incl - v;
            run += __shfl_sync(0xFFFFFFFFu, incl, 31);
        }
        if (lane == 0) h.nnz_row[row] = run;
    }
}

template <typename VT>
struct HvSeg { u32 start, len; VT av; };
struct HvPiece { u32 gbase, n, s0, s1; u64 av; };

// One kernel, two roles.  COUNT: units are count chunks (bitmap only); the per-numeric-chunk popcounts go to cnt.
// Otherwise: units are groups of numeric chunks; every chunk is accumulated densely and written at o.base[row] + cnt.
template <typename VT, int MODE, bool COUNT, bool BPAT>
__global__ void __launch_bounds__(HV_THREADS, HV_CTAS) k_hv(HvDev h, NumArgs<VT> a, OutArgs<VT> o, B200Ctrl *ctrl) {
    constexpr bool NEEDV = !COUNT && !BPAT;
    extern __shared__ __align__(128) unsigned char hv_smem[];
    __shared__ __align__(8) u64 s_bar[HV_NSTAGE];
    __shared__ HvPiece s_pd[HV_NSTAGE];
    __shared__ u32 s_warp[33], s_unit, s_qn[2], s_cnt[HV_MAX_SW];
    __shared__ u32 s_pre[HV_THREADS + 1], s_start[HV_THREADS];
    __shared__ u64 s_av[HV_THREADS];
    __shared__ HvSeg<VT> s_q[HV_QCAP];
    __shared__ u32 s_qpre[HV_QCAP + 1];

    const u32 tid = threadIdx.x, nt = HV_THREADS, lane = tid & 31;
    const u32 Wb = COUNT ? h.W * h.sw : h.W;                         // columns of one chunk of this kernel
    const u32 bwords = COUNT ? Wb >> 5 : 0u;                         // (the numeric pass needs no bitmap: a touched accumulator is never zero)
    u32 *bm = reinterpret_cast<u32 *>(hv_smem);
    Acc<MODE> acc;
    unsigned char *after_bm = hv_smem + (size_t)bwords * 4;
    if (!COUNT) acc.bind(after_bm, h.W);
    unsigned char *stage0 = after_bm + (COUNT ? 0 : Acc<MODE>::bytes(h.W));
    const u32 stage_bytes = HV_STAGE_ELEMS * (4u + (NEEDV ? (u32)sizeof(VT) : 0u));
    const u32 sm_stage0 = (u32)__cvta_generic_to_shared(stage0);
    const u32 sm_bar0 = (u32)__cvta_generic_to_shared(&s_bar[0]);

    for (u32 t = tid; t < bwords; t += nt) bm[t] = 0;
    if (!COUNT) for (u32 t = tid; t < h.W; t += nt) acc.clear(t);
    if (tid == 0) {
        for (int s = 0; s < HV_NSTAGE; s++) mbar_init(sm_bar0 + 8u * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    u32 phase_bits = 0;                                              // bit s: parity of the next completion of stage s
    u32 tile_par = 0;
    if (tid < 2) s_qn[tid] = 0;
    u64 vmax = 0;
    const u32 n_units = COUNT ? h.ctl->n_count_units : h.ctl->n_num_units;
    u32 *ticket = COUNT ? &h.ctl->t_count : &h.ctl->t_num;
    const u64 *units = COUNT ? h.units_c : h.units_n;

    while (true) {
        if (tid == 0) s_unit = atomicAdd(ticket, 1u);
        __syncthreads();
        const u32 u = s_unit;
        __syncthreads();
        if (u >= n_units) break;
        const u64 ud = units[u];
        const u32 li = (u32)(ud >> 32), cfirst = (u32)(ud >> 16) & 0xFFFFu, cn = (u32)ud & 0xFFFFu;
        const u32 row = h.list[li];
        const u64 rs = a.rpA[row];
        const u32 lenA = (u32)(a.rpA[row + 1] - rs);
        const u32 *__restrict__ Ac = a.colA + rs;
        const VT *__restrict__ Av = a.valA + rs;
        const u64 obase = COUNT ? 0ull : o.base[row];

        // A row of at most HV_THREADS entries (one tile) keeps its entries in registers over the unit's chunks: B row extent,
        // first / last column, a_ik, and a cursor -- the chunks ascend, so a segment starts where the last one ended and a
        // chunk costs one search per entry instead of the whole chain A.col -> span -> desc -> two searches.
        const bool single = lenA <= nt;
        u32 r_st = 0, r_len = 0, r_first = 0, r_last = 0, r_cur = 0; VT r_av = 0;
        if (single && tid < lenA) {
            const u32 k = Ac[tid];
            const uint4 sp = h.bspan[k];
            r_st = a.bdesc[k].x; r_len = sp.x; r_first = sp.y; r_last = sp.z;
            if (!COUNT) r_av = Av[tid];
            const u64 cs = (u64)cfirst * Wb;                                   // first column of the unit
            if (r_len && (u64)r_first < cs) r_cur = (u64)r_last < cs ? r_len : hv_lower_bound(a.colB + r_st, r_len, (u32)cs);
        }

        for (u32 ch = cfirst; ch < cfirst + cn; ch++) {
            const u64 c0 = (u64)ch * Wb;
            const u64 c1 = min((u64)h.ncols, c0 + Wb);
            u32 chunk_nnz = 1;                                             // numeric: an empty chunk is skipped altogether
            if (!COUNT) {
                const u32 *cr = h.cnt + (u64)li * h.nchunks;
                chunk_nnz = (ch + 1 < h.nchunks ? cr[ch + 1] : h.nnz_row[row]) - cr[ch];
                if (chunk_nnz == 0) {
                    // (the cursors of a register-resident row need not move: the chunk holds none of its entries)
                    continue;
                }
            }
            if (COUNT && tid < HV_MAX_SW) s_cnt[tid] = 0;
            // ---- accumulate: tiles of HV_THREADS entries of the A row
            for (u32 base = 0; base < lenA; base += nt) {
                // (the queue counters alternate: this tile fills s_qn[par], which the last tile but one left at zero)
                const u32 par = tile_par; tile_par ^= 1u;
                const u32 t = base + tid;
                u32 seg_start = 0, seg_len = 0; VT av = 0;
                if (single) {
                    if (r_cur < r_len && (u64)r_first < c1) {
                        const u32 hi = (u64)r_last < c1 ? r_len : r_cur + hv_lower_bound(a.colB + r_st + r_cur, r_len - r_cur, (u32)c1);
                        seg_start = r_st + r_cur; seg_len = hi - r_cur; av = r_av;
                        r_cur = hi;
                    }
                } else if (t < lenA) {
                    const u32 k = Ac[t];
                    const uint4 sp = h.bspan[k];                         // {len, first column, last column}
                    const u32 st = a.bdesc[k].x;
                    if (sp.x && (u64)sp.y < c1 && (u64)sp.z >= c0) {
                        const u32 lo = (u64)sp.y >= c0 ? 0u : hv_lower_bound(a.colB + st, sp.x, (u32)c0);
                        const u32 hi = (u64)sp.z < c1 ? sp.x : lo + hv_lower_bound(a.colB + st + lo, sp.x - lo, (u32)c1);
                        seg_start = st + lo; seg_len = hi - lo;
                        if (!COUNT) av = Av[t];
                    }
                }
                auto one = [&](u32 jb, u64 xav) {
                    const u32 idx = (u32)((u64)a.colB[jb] - c0);
                    if (COUNT) atomicOr(&bm[idx >> 5], 1u << (idx & 31));
                    else if (BPAT) acc.addv(idx, xav);
                    else acc.add(idx, (VT)xav, a.valB[jb]);
                };
                // three ways by segment length: a few entries -- the thread that found them walks them (most segments of a
                // power-law operand: no scan, no search); long -- queued for the bulk-copy stages; the rest -- balanced
                // enumeration (thread q takes products q, q + nt, ... of the tile's line of medium segments)
                if (seg_len >= HV_LONG_MIN) {
                    const u32 slot = atomicAdd(&s_qn[par], 1u);
                    if (slot < HV_QCAP) { s_q[slot].start = seg_start; s_q[slot].len = seg_len; s_q[slot].av = av; seg_len = 0; }
                } else if (seg_len <= HV_TINY_MAX) {
                    for (u32 j = 0; j < seg_len; j++) one(seg_start + j, (u64)av);
                    seg_len = 0;
                }
                if (tid == 0) s_qn[par ^ 1u] = 0;
                if (__syncthreads_or(seg_len != 0)) {
                    u32 total;
                    const u32 ex = hv_block_scan(seg_len, s_warp, total);
                    s_pre[tid] = ex; s_start[tid] = seg_start;
                    if (!COUNT) s_av[tid] = (u64)av;
                    __syncthreads();
                    for (u32 q = tid; q < total; q += nt) {
                        u32 lo = 0, hi = nt;
                        while (hi - lo > 1) { const u32 mid = (lo + hi) >> 1; if (s_pre[mid] <= q) lo = mid; else hi = mid; }
                        one(s_start[lo] + (q - s_pre[lo]), COUNT ? 0ull : s_av[lo]);
                    }
                }
                // long segments: pieces of <= HV_STAGE_ELEMS entries, 16-byte aligned in B's arrays, through the stages
                const u32 qn = min(s_qn[par], (u32)HV_QCAP);              // (final: a barrier lies behind every push)
                if (qn) {
                    u32 np = 0;
                    if (tid < qn) {
                        const u32 as = s_q[tid].start & ~3u;
                        const u32 ae = (s_q[tid].start + s_q[tid].len + 3u) & ~3u;
                        np = (ae - as + HV_STAGE_ELEMS - 1) / HV_STAGE_ELEMS;
                    }
                    u32 tp;
                    const u32 pex = hv_block_scan(np, s_warp, tp);
                    if (tid < qn) s_qpre[tid] = pex;
                    __syncthreads();
                    // producer state (thread 0): next piece to issue
                    u32 pi = 0, pq = 0, ppos = 0;
                    auto issue = [&]() {
                        const HvSeg<VT> sg = s_q[pq];
                        const u32 as = sg.start & ~3u, ae = (sg.start + sg.len + 3u) & ~3u;
                        if (pi == s_qpre[pq]) ppos = as;
                        const u32 n = min(HV_STAGE_ELEMS, ae - ppos);
                        const u32 st = pi % HV_NSTAGE;
                        HvPiece pd; pd.gbase = ppos; pd.n = n; pd.s0 = sg.start; pd.s1 = sg.start + sg.len; pd.av = (u64)sg.av;
                        s_pd[st] = pd;
                        const u32 dst = sm_stage0 + st * stage_bytes, bar = sm_bar0 + 8u * st;
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        mbar_expect_tx(bar, n * (4u + (NEEDV ? (u32)sizeof(VT) : 0u)));
                        bulk_g2s(dst, a.colB + ppos, n * 4u, bar);
                        if (NEEDV) bulk_g2s(dst + HV_STAGE_ELEMS * 4u, a.valB + ppos, n * (u32)sizeof(VT), bar);
                        ppos += n; pi++;
                        if (ppos >= ae) pq++;
                    };
                    if (tid == 0) for (int s = 0; s < HV_NSTAGE && pi < tp; s++) issue();
                    for (u32 i = 0; i < tp; i++) {
                        const u32 st = i % HV_NSTAGE;
                        while (!mbar_try_wait(sm_bar0 + 8u * st, (phase_bits >> st) & 1u)) { }
                        phase_bits ^= 1u << st;
                        const HvPiece pd = s_pd[st];
                        const u32 *scol = reinterpret_cast<const u32 *>(stage0 + (size_t)st * stage_bytes);
                        const VT *sval = reinterpret_cast<const VT *>(stage0 + (size_t)st * stage_bytes + HV_STAGE_ELEMS * 4u);
                        for (u32 e = tid; e < pd.n; e += nt) {
                            const u32 g = pd.gbase + e;
                            if (g >= pd.s0 && g < pd.s1) {
                                const u32 idx = (u32)((u64)scol[e] - c0);
                                if (COUNT) atomicOr(&bm[idx >> 5], 1u << (idx & 31));
                                else if (BPAT) acc.addv(idx, pd.av);
                                else acc.add(idx, (VT)pd.av, sval[e]);
                            }
                        }
                        __syncthreads();                                   // the stage is free again
                        if (tid == 0 && pi < tp) issue();
                    }
                }
                __syncthreads();
            }
            __syncthreads();
            // ---- emit
            if (COUNT) {
                const u32 wpt = (bwords + nt - 1) / nt, w0 = tid * wpt;
                const u32 wpc = h.W >> 5;                                  // bitmap words per numeric chunk
                u32 curj = 0xFFFFFFFFu, run = 0;
                for (u32 i = 0; i < wpt && w0 + i < bwords; i++) {
                    const u32 w = w0 + i, j = w / wpc;
                    if (j != curj) { if (run) atomicAdd(&s_cnt[curj], run); curj = j; run = 0; }
                    run += __popc(bm[w]); bm[w] = 0;
                }
                if (run) atomicAdd(&s_cnt[curj], run);
                __syncthreads();
                if (tid < h.sw) { const u32 nc = ch * h.sw + tid; if (nc < h.nchunks) h.cnt[(u64)li * h.nchunks + nc] = s_cnt[tid]; }
                __syncthreads();
            } else {
                // Every warp owns W / 32 consecutive columns and reads them 32 at a time (lane = column, conflict-free).  The
                // sums of stored values are never zero (no explicit zeros; 32-bit sums are proven, 64-bit ones cannot wrap,
                // saturating ones stick), so "accumulator != 0" is the chunk's pattern: ballots rank the entries, and the
                // lanes holding one write consecutive places of C.
                const u32 cpw = h.W / (HV_THREADS / 32), wbase = (tid >> 5) * cpw;   // columns per warp (a multiple of 32, at most 1024)
                u32 mine = 0, mybal = 0;                                   // lane i keeps the ballot of the warp's i-th group of 32 columns
                for (u32 i = 0, it = 0; i < cpw; i += 32, it++) {
                    const u32 b = __ballot_sync(0xFFFFFFFFu, acc.get(wbase + i + lane) != 0);
                    if (lane == it) mybal = b;
                    mine += __popc(b);
                }
                u32 total;
                const u32 wex = hv_block_scan(lane == 0 ? mine : 0u, s_warp, total);
                u64 pos = obase + h.cnt[(u64)li * h.nchunks + ch] + __shfl_sync(0xFFFFFFFFu, wex, 0);
                for (u32 i = 0, it = 0; i < cpw; i += 32, it++) {
                    const u32 b = __shfl_sync(0xFFFFFFFFu, mybal, it);
                    if (b == 0) continue;
                    if ((b >> lane) & 1u) {
                        const u32 idx = wbase + i + lane;
                        const u64 q = pos + __popc(b & ((1u << lane) - 1u));
                        const VT v = emit_val<VT>(acc.get(idx));
                        acc.clear(idx);
                        o.col[q] = (u32)(c0 + idx); put_val(o, q, v);
                        vmax = vmax > (u64)v ? vmax : (u64)v;
                    }
                    pos += __popc(b);
                }
                __syncthreads();
            }
        }
    }
    if (!COUNT) {
        vmax = warp_max_u64(vmax);
        if (lane == 0 && vmax) atomicMax(&ctrl->max_val_out, (ull)vmax);
    }
}

// ---------------------------------------------------------------------------- host side
struct HvKernel { const void *fn; size_t static_smem; };
static HvKernel g_hv[2][3][2];    // numeric: [value width][mode][pattern-only B]
static HvKernel g_hv_count;

template <typename VT, int MODE, bool COUNT, bool BPAT>
static void hv_register(HvKernel &k, size_t optin) {
    k.fn = (const void *)k_hv<VT, MODE, COUNT, BPAT>;
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, k.fn) == cudaSuccess) k.static_smem = fa.sharedSizeBytes; else { cudaGetLastError(); k.static_smem = 24 * 1024; }
    if (cudaFuncSetAttribute(k.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(optin - k.static_smem)) != cudaSuccess) cudaGetLastError();
}
static size_t g_hv_static = 0;
void hv_setup(b200_ctx *ctx) {
    const size_t o = ctx->smem_optin;
    hv_register<u32, 0, false, false>(g_hv[0][0][0], o); hv_register<u32, 0, false, true>(g_hv[0][0][1], o);
    hv_register<u32, 1, false, false>(g_hv[0][1][0], o); hv_register<u32, 1, false, true>(g_hv[0][1][1], o);
    hv_register<u64, 0, false, false>(g_hv[1][0][0], o); hv_register<u64, 0, false, true>(g_hv[1][0][1], o);
    hv_register<u64, 1, false, false>(g_hv[1][1][0], o); hv_register<u64, 1, false, true>(g_hv[1][1][1], o);
    hv_register<u64, 2, false, false>(g_hv[1][2][0], o); hv_register<u64, 2, false, true>(g_hv[1][2][1], o);
    hv_register<u32, 0, true, true>(g_hv_count, o);
    for (int v = 0; v < 2; v++) for (int m = 0; m < 3; m++) for (int b = 0; b < 2; b++) g_hv_static = std::max(g_hv_static, g_hv[v][m][b].static_smem);
    g_hv_static = std::max(g_hv_static, g_hv_count.static_smem);
}

// Decide whether (and how) the chunked kernels take part in this multiply.  mode: accumulator mode of the multiply.
// Returns plan.on == false when the heavy rows are better served elsewhere (narrow column space: k_num_rank).
int hv_plan(b200_ctx *ctx, const b200_csr *A, const b200_csr *B, int mode, u64 p_bound, HvPlan *plan) {
    memset(plan, 0, sizeof(*plan));
    if (!ctx->cfg.heavy_kernel || p_bound <= (u64)b200_hash_cap(B200_NUM_HASH_BINS - 1) || B->nnz >= 0xFFFFFFF0ull) return B200_OK;
    const u64 ncols = B->cols;
    const size_t accb = mode == 0 ? 4 : 8;
    {   // heavy rows over a narrow column space keep the rank kernel (api.cu, launch_numeric): a whole-row bitmap fits there
        const u64 heavy_cap = std::min<u64>(p_bound, ncols);
        if (heavy_cap < 65536 && (size_t)((ncols + 31) / 32) * 6 + 16 + heavy_cap * (4 + accb) <= ctx->smem_optin - 1024 && ctx->cfg.heavy_chunk_cols <= 0) return B200_OK;
    }
    const size_t budget = (HV_CTAS == 1 ? ctx->smem_optin : (size_t)(228 * 1024) / HV_CTAS - 1024) - g_hv_static - 512;
    const size_t stages = (size_t)HV_NSTAGE * HV_STAGE_ELEMS * (4 + A->val_bits / 8);
    if (budget <= stages + 8192) return B200_OK;
    u64 W = (u64)(((budget - stages) * 8) / (accb * 8 + 1));           // an accumulator and a bitmap bit per column
    if (ctx->cfg.heavy_chunk_cols > 0) W = std::min<u64>(W, (u64)ctx->cfg.heavy_chunk_cols);
    W = std::min<u64>(W, (u64)32 * HV_THREADS);                       // (a warp emits at most 1024 columns of a chunk)
    u64 Wp = 1024; while (Wp * 2 <= W) Wp *= 2;                       // a power of two, at least 1024 columns
    W = Wp;
    if ((size_t)W * accb + W / 8 + stages > budget) return B200_OK;
    const u64 nchunks = (ncols + W - 1) / W;
    if (nchunks > 0xFFFF) return B200_OK;
    u64 sw = std::min<u64>(std::min<u64>(HV_MAX_SW, nchunks), std::max<u64>(1, (1ull << 20) / W));
    while (sw > 1 && (size_t)(sw * W / 8) + (size_t)HV_NSTAGE * HV_STAGE_ELEMS * 4 > budget) sw--;
    plan->W = (u32)W; plan->nchunks = (u32)nchunks; plan->sw = (u32)sw; plan->nchunks_c = (u32)((nchunks + sw - 1) / sw);
    // Rows take the chunked kernels when their products average >= 1024 per chunk (a chunk costs a few barriers and a walk
    // over its accumulators whatever it holds).  Over few chunks that is below the largest hash bins' capacity: those rows
    // (>= 512 products per chunk) then join the heavy list in the pre-pass instead of going through hash + sort.
    const u64 top = (u64)b200_hash_cap(B200_NUM_HASH_BINS - 1) + 1;
    const u64 from = std::min<u64>(top, std::max<u64>(1025, nchunks * 512));
    plan->heavy_from = (u32)from;
    plan->pmin = from < top ? from : std::max<u64>(top, nchunks * 1024);
    if (ctx->cfg.heavy_min_products > 0) { plan->pmin = std::max<u64>((u64)ctx->cfg.heavy_min_products, 33); plan->heavy_from = (u32)std::min<u64>(top, plan->pmin); }
    plan->psplit = ctx->cfg.heavy_unit_products > 0 ? (u64)ctx->cfg.heavy_unit_products : 1ull << 20;
    if (p_bound < plan->pmin) return B200_OK;                          // no row can qualify
    // per-(row, chunk) counters: at most 256 MiB; rows of the heavy list beyond that keep the global-table kernel
    const u64 by_mem = ((u64)256 << 20) / (nchunks * 4);
    plan->cap_li = (u32)std::min<u64>(std::min<u64>(A->rows, by_mem), 0xFFFFFFFFull);
    // the unit lists hold (rows taken) x (chunks) entries at most; rows taken <= bound of all products / pmin
    const u64 by_prod = (u64)std::min<unsigned __int128>((unsigned __int128)A->nnz * B->max_row_len / plan->pmin + 1, (unsigned __int128)A->rows);
    const u64 rows_taken_max = std::min<u64>(plan->cap_li, by_prod);
    const size_t cnt_bytes = ((size_t)plan->cap_li * nchunks * 4 + 255) & ~(size_t)255;
    const size_t unit_bytes_c = ((size_t)rows_taken_max * plan->nchunks_c * 8 + 255) & ~(size_t)255;
    const size_t unit_bytes_n = (size_t)rows_taken_max * nchunks * 8;
    const size_t total = 256 + cnt_bytes + unit_bytes_c + unit_bytes_n;
    if (total > ctx->cap_hv) {
        dfree(ctx, ctx->d_hv); ctx->d_hv = nullptr; ctx->cap_hv = 0;
        TRY(dmalloc(ctx, &ctx->d_hv, total));
        ctx->cap_hv = total;
    }
    unsigned char *base = (unsigned char *)ctx->d_hv;
    plan->ctl = base; plan->cnt = base + 256; plan->units_c = base + 256 + cnt_bytes; plan->units_n = base + 256 + cnt_bytes + unit_bytes_c;
    plan->on = true;
    return B200_OK;
}

static HvDev hv_dev(b200_ctx *ctx, const b200_csr *B, B200Ctrl *ctrl, const HvPlan &p) {
    HvDev h;
    h.list = ctx->d_bin_rows + (u64)B200_BIN_HEAVY * ctx->cap_rows; h.list_count = &ctrl->sym_bin_count[B200_BIN_HEAVY];
    h.prod = ctx->d_prod; h.pmin = p.pmin; h.cap_li = p.cap_li; h.psplit = p.psplit;
    h.ncols = (u32)B->cols; h.W = p.W; h.nchunks = p.nchunks; h.sw = p.sw; h.nchunks_c = p.nchunks_c;
    h.cnt = (u32 *)p.cnt; h.units_c = (u64 *)p.units_c; h.units_n = (u64 *)p.units_n; h.ctl = (HvCtl *)p.ctl;
    h.bspan = B->d_span; h.nnzB = B->nnz; h.nnz_row = ctx->d_nnz_row;
    return h;
}

// Count pass of the rows the chunked kernels take: plan -> per-chunk lengths -> offsets; the rows' lengths land in
// ctx->d_nnz_row like every other count kernel's.
int hv_count(b200_ctx *ctx, const b200_csr *A, const b200_csr *B, B200Ctrl *ctrl, const HvPlan &p, cudaStream_t s) {
    const HvDev h = hv_dev(ctx, B, ctrl, p);
    CUDA_TRY(cudaMemsetAsync(p.ctl, 0, sizeof(HvCtl), s));
    const int pg = (int)std::max<u64>(1, std::min<u64>((A->rows + 255) / 256, (u64)ctx->num_sms * 4));
    k_hv_plan<<<pg, 256, 0, s>>>(h);
    LAUNCH_CHECK(ctx);
    const size_t smem = (size_t)p.W * p.sw / 8 + (size_t)HV_NSTAGE * HV_STAGE_ELEMS * 4;
    NumArgs<u32> na{A->d_rp, A->d_col, nullptr, B->d_desc, B->d_col, nullptr};
    OutArgs<u32> o{nullptr, nullptr, nullptr, nullptr, nullptr, 0u, 0u};
    void *kargs[] = {(void *)&h, (void *)&na, (void *)&o, (void *)&ctrl};
    CUDA_TRY(cudaLaunchKernel(g_hv_count.fn, dim3(ctx->num_sms * HV_CTAS), dim3(HV_THREADS), kargs, smem, s));
    LAUNCH_CHECK(ctx);
    k_hv_scan<<<pg, 256, 0, s>>>(h);
    LAUNCH_CHECK(ctx);
    return B200_OK;
}

template <typename VT>
static int hv_numeric_t(b200_ctx *ctx, const b200_csr *A, const b200_csr *B, B200Ctrl *ctrl, const HvPlan &p, int mode, bool bpat,
                        const u64 *base, u32 *col, void *val, u32 narrow, cudaStream_t s) {
    const HvDev h = hv_dev(ctx, B, ctrl, p);
    const bool v64 = sizeof(VT) == 8;
    const HvKernel &k = g_hv[v64 ? 1 : 0][v64 ? mode : std::min(mode, 1)][bpat ? 1 : 0];
    if (!k.fn) return set_err(B200_ERR_CUDA, "chunked heavy-row kernel variant is not registered");
    const size_t accb = mode == 0 ? 4 : 8;
    const size_t smem = (size_t)p.W / 8 + (size_t)p.W * accb + (size_t)HV_NSTAGE * HV_STAGE_ELEMS * (4 + (bpat ? 0 : sizeof(VT)));
    NumArgs<VT> na{A->d_rp, A->d_col, (const VT *)A->d_val, B->d_desc, B->d_col, (const VT *)B->d_val};
    OutArgs<VT> o{base, col, (VT *)val, nullptr, nullptr, 0u, narrow};
    void *kargs[] = {(void *)&h, (void *)&na, (void *)&o, (void *)&ctrl};
    CUDA_TRY(cudaLaunchKernel(k.fn, dim3(ctx->num_sms * HV_CTAS), dim3(HV_THREADS), kargs, smem, s));
    LAUNCH_CHECK(ctx);
    return B200_OK;
}
int hv_numeric(b200_ctx *ctx, const b200_csr *A, const b200_csr *B, B200Ctrl *ctrl, const HvPlan &p, int mode, bool bpat,
               const u64 *base, u32 *col, void *val, u32 narrow, cudaStream_t s) {
    if (A->val_bits == 32) return hv_numeric_t<u32>(ctx, A, B, ctrl, p, mode, bpat, base, col, val, narrow, s);
    return hv_numeric_t<u64>(ctx, A, B, ctrl, p, mode, bpat, base, col, val, narrow, s);
}
