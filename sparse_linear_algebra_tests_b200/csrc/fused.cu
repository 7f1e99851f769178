// fused.cu -- the fused pipeline of one C = A x B: two launches, no host round trip, C written once.
//
//   k_fz_prepass   per row of A: intermediate-product count P_i, class (empty / tiny / dense / other) and, for dense
//                  rows, the origin of the row's column window; cuts the rows, IN ROW ORDER, into work units (a run of
//                  tiny rows -- one per warp --, one dense row, or a run of rows that only need placing) through a
//                  decoupled look-back scan of the per-CTA unit counts; totals; bin lists of the "other" rows.
//   k_fz_numeric   persistent CTAs take units by ticket, in row order.
//                    dense row : every intermediate product goes straight into a dense accumulator window in shared
//                                memory (acc[col - origin] += a_ik * b_kj, one bit per touched column in a bitmap) -- MAGNUS'
//                                "dense accumulation" category with the window in shared memory instead of L2.  No product
//                                buffer, no hash, no sort: the bitmap is walked in column order on the way out.
//                    tiny rows : one warp per row, products in registers, shuffle bitonic sort, segmented saturating scan.
//                    other rows: their exact length was counted beforehand (hash / heavy lists); only placed here.
//                  Every unit publishes its entry count and finds its place in C by a decoupled look-back over the unit
//                  status words (the hand-written look-back scan of the north star, fused into the numeric kernel), so
//                  rows are written ONCE, at their final offsets, together with row_ptr -- no scratch CSR, no
//                  compaction, no separate row_ptr scan.
// C is allocated from a bound before the kernel runs; its exact size (and the largest value, the longest row, the
// product count) reach the host afterwards through a pinned {word, epoch} mirror that nobody waits for until the
// numbers are needed (resolve_pending).  The buffers between the two kernels clean themselves: the pre-pass zeroes the
// unit status words, the numeric kernel zeroes the pre-pass status words and, last thing, the control block.
//
// Replaces, for the rows it covers, CsrMatrix::matmul_par's three traversals of the product stream + per-row sort
// (/root/reference/src/graph_csr.rs:362-403 symbolic, :412-417 prefix sum, :430-476 numeric).
#include "engine.cuh"
#include "devutil.cuh"

#define FZ_EMPTY 0
#define FZ_TINY 1
#define FZ_DENSE 2
#define FZ_OTHER 3
#define FZ_PRE_MIN_ROWS 32
#define FZ_PRE_ROWS(G) ((256 / (G)) > FZ_PRE_MIN_ROWS ? (256 / (G)) : FZ_PRE_MIN_ROWS)   // rows per pre-pass CTA

struct FzPreArgs {
    u64 rows, ncols;
    const u64 *rpA; const u32 *colA; const uint2 *bdesc;
    u32 wcap;                  // dense window capacity in columns (multiple of 32); 0: no dense class
    long long cs_lo, cs_hi;    // WMODE 1: every entry (k, c) of B has c - k in [cs_lo, cs_hi] (circularly)
    u64 dense_pmax;            // rows with more products than this are never dense (they would pin one CTA for too long)
    u32 tiny_run, other_run;   // longest run of tiny / placed-only rows that forms one unit
    unsigned char *rowclass; u32 *roworg; u32 *units;
    u64 *tile_status, *unit_status;
    B200Ctrl *ctrl; u32 *bin_rows; u32 bin_stride;
    u64 *host_mirror; u32 epoch;
};

// warp-wide decoupled look-back: publish `agg` for tile/unit idx, return the sum over all predecessors
__device__ __forceinline__ u64 fz_lookback(u64 *status, u32 idx, u64 agg, int lane) {
    if (lane == 0) atomicExch((ull *)&status[idx], (ull)((idx == 0 ? SCAN_FLAG_PRE : SCAN_FLAG_AGG) | agg));
    u64 excl = 0;
    if (idx > 0) {
        long long look = (long long)idx - 1;
        while (true) {
            const long long i = look - lane;
            u64 st;
            do { st = i >= 0 ? ld_volatile_u64(&status[i]) : SCAN_FLAG_PRE; } while (__any_sync(0xFFFFFFFFu, (st >> 62) == 0));
            const u32 pre_mask = __ballot_sync(0xFFFFFFFFu, (st >> 62) == 2);
            const int first = pre_mask ? __ffs(pre_mask) - 1 : 32;      // nearest predecessor that knows its full prefix
            excl += warp_sum_u64(lane <= first ? (st & SCAN_VAL_MASK) : 0ull);
            if (pre_mask) break;
            look -= 32;
        }
        if (lane == 0) atomicExch((ull *)&status[idx], (ull)(SCAN_FLAG_PRE | (excl + agg)));
    }
    return excl;
}

// the control block -> pinned host memory as self-validating {word, epoch} chunks (the host polls, never synchronises)
__device__ __forceinline__ void fz_report(const B200Ctrl *ctrl, u64 *host_mirror, u32 epoch) {
    const volatile u32 *src = reinterpret_cast<const volatile u32 *>(ctrl);
    for (u32 i = threadIdx.x; i < sizeof(B200Ctrl) / 4; i += blockDim.x) st_volatile_u64(host_mirror + i, ((u64)epoch << 32) | (u64)src[i]);
}

// =======================================================================================
// 1. pre-pass
// =======================================================================================
// WMODE 0: the dense window is the whole column space (origin 0 for every row).
// WMODE 1 (square B with known offset bounds): the columns of row i of C lie on the arc
//     [min_k + cs_lo, max_k + cs_hi] of the index circle, k over the columns of row i of A; min/max are taken in the frame
//     d(k) = (k - ref + n/2) mod n of the row's first column, so rows that wrap around the end of the index space keep a
//     short arc.  The origin is moved back by < 32 columns so that the wrap point (column n -> 0) falls on a bitmap word
//     boundary: the emit phase then only has to start its walk at that word to produce ascending columns.
template <int G, int WMODE>
__global__ void __launch_bounds__(256) k_fz_prepass(FzPreArgs p) {
    constexpr int RPS = 256 / G;
    constexpr int TILE = FZ_PRE_ROWS(G);
    constexpr int STEPS = TILE / RPS;
    __shared__ u32 s_tile, s_cnt[B200_NBINS], s_base[B200_NBINS], s_binloc[TILE], s_ccnt[4];
    __shared__ unsigned char s_class[TILE];
    __shared__ u32 s_w[8], s_w2[8], s_ubase, s_last;
    __shared__ ull s_sum, s_max, s_bound;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) { s_tile = atomicAdd(&p.ctrl->scan_ticket[1], 1u); s_sum = 0; s_max = 0; s_bound = 0; }
    if (tid < B200_NBINS) s_cnt[tid] = 0;
    if (tid < 4) s_ccnt[tid] = 0;
    __syncthreads();
    const u32 tile = s_tile;
    const u32 sub = tid % G;
    const long long n = (long long)p.ncols, half = n / 2;
    u64 wsum = 0, wmax = 0, wbound = 0;
#pragma unroll
    for (int step = 0; step < STEPS; step++) {
        const u32 lrow = step * RPS + tid / G;
        const u64 row = (u64)tile * TILE + lrow;
        u64 pr = 0; u32 lenA = 0, dmin = 0xFFFFFFFFu, dmax = 0;
        long long ref = 0;
        if (row < p.rows) {
            const u64 s = p.rpA[row];
            lenA = (u32)(p.rpA[row + 1] - s);
            const u32 *Ac = p.colA + s;
            if (WMODE == 1 && lenA) ref = (long long)Ac[0];
            auto see = [&](u32 k) {
                if (WMODE == 1) {
                    long long t = (long long)k - ref + half;
                    if (t < 0) t += n; else if (t >= n) t -= n;
                    dmin = min(dmin, (u32)t); dmax = max(dmax, (u32)t);
                }
            };
            u32 i = sub;
            for (; i + 3 * G < lenA; i += 4 * G) {                           // four independent gathers in flight
                const u32 k0 = Ac[i], k1 = Ac[i + G], k2 = Ac[i + 2 * G], k3 = Ac[i + 3 * G];
                const u32 d0 = p.bdesc[k0].y, d1 = p.bdesc[k1].y, d2 = p.bdesc[k2].y, d3 = p.bdesc[k3].y;
                pr += (u64)d0 + d1 + d2 + d3;
                see(k0); see(k1); see(k2); see(k3);
            }
            for (; i < lenA; i += G) { const u32 k = Ac[i]; pr += p.bdesc[k].y; see(k); }
        }
#pragma unroll
        for (int m = G / 2; m > 0; m >>= 1) {
            pr += shfl_xor_u64(pr, m);
            if (WMODE == 1) {
                dmin = min(dmin, __shfl_xor_sync(0xFFFFFFFFu, dmin, m));
                dmax = max(dmax, __shfl_xor_sync(0xFFFFFFFFu, dmax, m));
            }
        }
        if (sub == 0) {
            u32 cls = 0xFF, binloc = (u32)B200_BIN_NONE << 24;
            if (row < p.rows) {
                cls = FZ_EMPTY;
                if (pr > 0) {
                    if (pr <= 32 && lenA <= 32) cls = FZ_TINY;
                    else {
                        bool fits = false; u32 org = 0;
                        if (WMODE == 0) fits = p.wcap != 0;
                        else if (p.wcap) {
                            const long long lo = (long long)dmin + p.cs_lo, hi = (long long)dmax + p.cs_hi;
                            long long width = hi - lo + 1;
                            long long oc = (ref - half + lo) % n; if (oc < 0) oc += n;
                            const long long delta = ((oc % 32) - (n % 32) + 32) % 32;   // (n - origin) becomes a multiple of 32
                            if (oc >= delta) { oc -= delta; width += delta; }
                            else { width = oc + width <= n ? oc + width : n; oc = 0; }   // origin 0: the window [0, width) needs no rotation
                            fits = width <= (long long)p.wcap && width <= n;
                            org = (u32)oc;
                        }
                        if (fits && pr <= p.dense_pmax) { cls = FZ_DENSE; p.roworg[row] = org; }
                        else {
                            cls = FZ_OTHER;
                            int b = b200_bin_by_size(pr);
                            if (b == B200_BIN_HASH0) b = B200_BIN_HASH0 + 1;           // the two smallest hash bins share a list
                            if (b != B200_BIN_HEAVY) b = B200_BIN_WIDE0 + (b - B200_BIN_HASH0);
                            binloc = ((u32)b << 24) | atomicAdd(&s_cnt[b], 1u);
                        }
                    }
                    wsum += pr; wmax = wmax > pr ? wmax : pr; wbound += pr < p.ncols ? pr : p.ncols;
                }
                p.rowclass[row] = (unsigned char)cls;
                atomicAdd(&s_ccnt[cls], 1u);
            }
            s_class[lrow] = (unsigned char)cls; s_binloc[lrow] = binloc;
        }
    }
    wsum = warp_sum_u64(wsum); wmax = warp_max_u64(wmax); wbound = warp_sum_u64(wbound);
    if (lane == 0) { if (wsum) { atomicAdd(&s_sum, (ull)wsum); atomicAdd(&s_bound, (ull)wbound); } atomicMax(&s_max, (ull)wmax); }
    __syncthreads();
    // ---- units: a row starts one when its family (placed-only / tiny / dense) changes, when it is dense, or when the run
    //      it belongs to has reached the longest run one unit takes
    const u32 cls = tid < TILE ? s_class[tid] : 0xFFu;
    const bool valid = cls != 0xFFu;
    const u32 fam = cls == FZ_TINY ? 1u : cls == FZ_DENSE ? 2u : 0u;
    u32 pfam = 0xFFu;
    if (valid && tid > 0) { const u32 pc = s_class[tid - 1]; pfam = pc == FZ_TINY ? 1u : pc == FZ_DENSE ? 2u : 0u; }
    const bool brk = valid && (tid == 0 || fam != pfam || fam == 2u);
    u32 last = brk ? (u32)tid + 1u : 0u;                                      // inclusive max-scan: last break at or before me
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const u32 t = __shfl_up_sync(0xFFFFFFFFu, last, d); if (lane >= d) last = max(last, t); }
    if (lane == 31) s_w[w] = last;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; i++) if (i < w) last = max(last, s_w[i]);
    const u32 pos = valid ? (u32)tid + 1u - last : 0u;
    const u32 maxrun = fam == 1u ? p.tiny_run : p.other_run;
    const bool start = valid && (brk || pos % maxrun == 0);
    u32 incl = start ? 1u : 0u;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const u32 t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
    if (lane == 31) s_w2[w] = incl;
    if (tid < B200_NBINS) { const u32 c = s_cnt[tid]; s_base[tid] = c ? atomicAdd(&p.ctrl->sym_bin_count[tid], c) : 0u; }
    __syncthreads();
    u32 wbase = 0, agg = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { if (i < w) wbase += s_w2[i]; agg += s_w2[i]; }
    if (w == 0) {
        const u64 excl = fz_lookback(p.tile_status, tile, (u64)agg, lane);
        if (lane == 0) s_ubase = (u32)excl;
    }
    __syncthreads();
    const u64 r = (u64)tile * TILE + tid;
    if (tid < TILE && r < p.rows) {
        if (start) p.units[s_ubase + wbase + incl - 1] = (u32)r;
        const u32 bl = s_binloc[tid], b = bl >> 24;
        if (b != B200_BIN_NONE) p.bin_rows[(u64)b * p.bin_stride + s_base[b] + (bl & 0xFFFFFFu)] = (u32)r;
        p.unit_status[r] = 0;                                                // the numeric kernel's look-back starts clean
    }
    if (tid == 0) {
        if (tile == gridDim.x - 1) { p.units[s_ubase + agg] = (u32)p.rows; p.ctrl->num_units = s_ubase + agg; }
        if (s_sum) { atomicAdd(&p.ctrl->total_products, s_sum); atomicAdd(&p.ctrl->total_bound, s_bound); }
        if (s_max) atomicMax(&p.ctrl->max_row_products, s_max);
    }
    if (tid < 4 && s_ccnt[tid]) atomicAdd(&p.ctrl->class_count[tid], s_ccnt[tid]);
    if (p.host_mirror) {                                                     // the host is waiting for the totals (bound, classes)
        __syncthreads();
        if (tid == 0) { __threadfence(); s_last = atomicAdd(&p.ctrl->pre_done, 1u) == gridDim.x - 1 ? 1u : 0u; }
        __syncthreads();
        if (s_last) { __threadfence(); fz_report(p.ctrl, p.host_mirror, p.epoch); }
    }
}

// =======================================================================================
// 2. numeric + placement
// =======================================================================================
template <typename VT>
struct FzNumArgs {
    NumArgs<VT> a;
    const uint4 *pack;
    const u32 *units; const unsigned char *rowclass; const u32 *roworg; const u32 *nnz_row;
    u64 *unit_status; u64 *tile_status; u64 n_tile_status;
    B200Ctrl *ctrl;
    u64 *rpC; u32 *colC; VT *valC;
    u32 ncols, wcap, nw;       // dense window: wcap columns = nw bitmap words
    int wmode, bpat, lg, narrow, finalize;
    u64 *host_mirror; u32 epoch;
    ull *maxval_dst;           // the product handle's largest-value scalar
};

template <typename VT, int MODE>
__device__ __forceinline__ u64 fz_product(VT av, VT bv, bool bpat) {
    if (MODE == 0) return (u64)((u32)av * (bpat ? 1u : (u32)bv));
    if (bpat) return (u64)av;
    if (MODE == 2) return sat_mul((u64)av, (u64)bv);
    u64 x = (u64)av * (u64)bv;
    if (sizeof(VT) == 4) x = x > 0xFFFFFFFFull ? 0xFFFFFFFFull : x;
    return x;
}

template <typename VT, int MODE, bool PACK>
__global__ void __launch_bounds__(256) k_fz_numeric(FzNumArgs<VT> f) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ u32 s_ticket[2], s_wsum[8], s_wbase[8];
    __shared__ u64 s_excl;
    __shared__ u32 s_total, s_last;
    const u32 nt = blockDim.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nwarps = nt >> 5;
    Acc<MODE> acc; acc.bind(smem_raw, f.wcap);
    u32 *bm = reinterpret_cast<u32 *>(smem_raw + Acc<MODE>::bytes(f.wcap));
    for (u32 t = tid; t < f.wcap; t += nt) acc.clear(t);
    for (u32 t = tid; t < f.nw; t += nt) bm[t] = 0;
    // the pre-pass is done with its look-back words: leave them zeroed for the next multiply
    for (u64 i = (u64)blockIdx.x * nt + tid; i < f.n_tile_status; i += (u64)gridDim.x * nt) f.tile_status[i] = 0;
    if (tid == 0) s_ticket[0] = atomicAdd(&f.ctrl->unit_ticket, 1u);
    __syncthreads();
    const u32 num_units = f.ctrl->num_units;
    const u32 wpt = (f.nw + nt - 1) / nt;                                    // bitmap words per thread in the count / emit walks
    const u32 ncols = f.ncols;
    const bool bpat = f.bpat != 0;
    u64 vmax = 0; u32 maxnnz = 0;
    u32 it = 0;
    while (true) {
        const u32 u = s_ticket[it & 1];
        if (u >= num_units) break;
        u32 next_ticket = 0;
        if (tid == 0) next_ticket = atomicAdd(&f.ctrl->unit_ticket, 1u);    // its latency hides behind this unit (stored at the end)
        it++;
        const u32 first = f.units[u], end = f.units[u + 1];
        const u32 cls = f.rowclass[first];
        u32 cnt = 0;                                                         // my share of the unit's entries (count phase)
        // unit-type specific state that lives across the placement barrier
        u32 org = 0, rot = 0;                                                // dense
        u32 tkey = B200_EMPTY_KEY; VT tval = 0; u32 ttails = 0; u32 trow = 0; bool tactive = false;   // tiny
        if (cls == FZ_DENSE) {
            // ---------------------------------------------------------------- scatter the products into the window
            const u64 rs = f.a.rpA[first];
            const u32 lenA = (u32)(f.a.rpA[first + 1] - rs);
            const u32 *Ac = f.a.colA + rs;
            const VT *Av = f.a.valA + rs;
            org = f.wmode ? f.roworg[first] : 0u;
            auto put = [&](u32 c, VT av, u32 jb) {
                const u64 x = fz_product<VT, MODE>(av, bpat ? (VT)1 : f.a.valB[jb], bpat);
                u32 d = c - org; if (c < org) d += ncols;
                acc.addv(d, x);
                atomicOr(&bm[d >> 5], 1u << (d & 31));
            };
            if constexpr (PACK) {
                auto expand = [&](const PackRec &r, VT av) {
                    const u32 len = r.a.y, st = r.a.x;
                    if (len > 0) put(r.a.z, av, st);
                    if (len > 1) put(r.a.w, av, st + 1);
                    if (len > 2) put(r.b.x, av, st + 2);
                    if (len > 3) put(r.b.y, av, st + 3);
                    if (len > 4) put(r.b.z, av, st + 4);
                    if (len > 5) put(r.b.w, av, st + 5);
                    for (u32 j = B200_PACK_INLINE; j < len; j++) put(f.a.colB[st + j], av, st + j);
                };
                for (u32 t = tid; t < lenA; t += 2 * nt) {
                    const u32 t1 = t + nt;
                    const bool h1 = t1 < lenA;
                    const u32 k0 = Ac[t], k1 = h1 ? Ac[t1] : k0;
                    const VT a0 = Av[t], a1 = h1 ? Av[t1] : (VT)0;
                    const PackRec r0 = load_pack(f.pack, k0);
                    PackRec r1 = load_pack(f.pack, k1);
                    if (!h1) r1.a.y = 0;
                    expand(r0, a0); expand(r1, a1);
                }
            } else {
                const u32 G = 1u << f.lg, sub = tid & (G - 1), grp = tid >> f.lg, ngrp = nt >> f.lg;
                walk_products<VT>(Ac, lenA, f.a.bdesc, grp, ngrp, sub, G, [&](u32 t) { return Av[t]; },
                                  [&](VT av, u32 jb) { put(f.a.colB[jb], av, jb); });
            }
            __syncthreads();
            // ---------------------------------------------------------------- count: my words of the bitmap, in column order
            if (org) { const u32 sw = (ncols - org) >> 5; rot = sw < f.nw ? sw : 0u; }   // (ncols - org) is a multiple of 32 (pre-pass)
            const u32 l0 = tid * wpt;
            for (u32 i = 0; i < wpt; i++) {
                const u32 l = l0 + i;
                if (l < f.nw) { u32 pw = l + rot; if (pw >= f.nw) pw -= f.nw; cnt += __popc(bm[pw]); }
            }
        } else if (cls == FZ_TINY) {
            // ---------------------------------------------------------------- one tiny row per warp, products in registers
            trow = first + w;
            tactive = trow < end;
            u32 dA = 0, k = 0; VT av = 0;
            if (tactive) {
                const u64 s = f.a.rpA[trow];
                dA = (u32)(f.a.rpA[trow + 1] - s);
                if (lane < dA) { k = f.a.colA[s + lane]; av = f.a.valA[s + lane]; }
            }
            tiny_gather<VT, true>(dA, k, av, f.a.bdesc, f.a.colB, f.a.valB, (int)lane, tkey, tval, bpat);
            if (f.narrow) {                                                  // columns < 2^27: sort one packed word per lane
                u32 packed = tkey == B200_EMPTY_KEY ? B200_EMPTY_KEY : (tkey << 5) | lane, none = 0;
                warp_bitonic<u32, false>(packed, none, (int)lane);
                tkey = packed == B200_EMPTY_KEY ? B200_EMPTY_KEY : packed >> 5;
                tval = shfl_any(tval, (int)(packed & 31u));
            } else {
                warp_bitonic<VT, true>(tkey, tval, (int)lane);
            }
            if (MODE == 0) {                                                 // row sums proven < 2^32: scan 32-bit words
                u32 v = (u32)tval;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const u32 pk = __shfl_up_sync(0xFFFFFFFFu, tkey, d);
                    const u32 pv = __shfl_up_sync(0xFFFFFFFFu, v, d);
                    if (lane >= (u32)d && pk == tkey) v += pv;
                }
                tval = (VT)v;
            } else {
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const u32 pk = __shfl_up_sync(0xFFFFFFFFu, tkey, d);
                    const VT pv = shfl_up_any(tval, d);
                    if (lane >= (u32)d && pk == tkey) tval = sat_add(tval, pv);
                }
            }
            const u32 nk = __shfl_down_sync(0xFFFFFFFFu, tkey, 1);
            const bool tail = tkey != B200_EMPTY_KEY && (lane == 31 || nk != tkey);
            ttails = __ballot_sync(0xFFFFFFFFu, tail);
            cnt = lane == 0 ? __popc(ttails) : 0u;
        } else {
            // ---------------------------------------------------------------- rows that are only placed (empty / counted beforehand)
            const u32 r = first + tid;
            if (r < end && f.rowclass[r] == FZ_OTHER) cnt = f.nnz_row[r];
        }
        // -------------------------------------------------------------------- place the unit: block scan + look-back
        u32 incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const u32 t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= (u32)d) incl += t; }
        if (lane == 31) s_wsum[w] = incl;
        __syncthreads();
        if (w == 0) {
            const u32 x = lane < nwarps ? s_wsum[lane] : 0u;
            u32 xi = x;
#pragma unroll
            for (int d = 1; d < 8; d <<= 1) { const u32 t = __shfl_up_sync(0xFFFFFFFFu, xi, d); if (lane >= (u32)d) xi += t; }
            if (lane < 8) s_wbase[lane] = xi - x;
            const u32 total = __shfl_sync(0xFFFFFFFFu, xi, 7);
            const u64 excl = fz_lookback(f.unit_status, u, (u64)total, (int)lane);
            if (lane == 0) { s_excl = excl; s_total = total; }
        }
        __syncthreads();
        const u64 ubase = s_excl;
        const u64 wb = ubase + s_wbase[w];                                   // where my warp's entries start
        u64 run = wb + incl - cnt;                                           // where mine start
        if (u == 0 && tid == 0) f.rpC[0] = 0;
        if (u + 1 == num_units && tid == 0) f.ctrl->total_nnz = ubase + s_total;
        if (cls == FZ_DENSE) {
            // ---------------------------------------------------------------- emit in column order, clearing as we go
            const u32 l0 = tid * wpt;
            for (u32 i = 0; i < wpt; i++) {
                const u32 l = l0 + i;
                if (l >= f.nw) break;
                u32 pw = l + rot; if (pw >= f.nw) pw -= f.nw;
                u32 word = bm[pw];
                if (word) bm[pw] = 0;
                while (word) {
                    const u32 b = __ffs(word) - 1; word &= word - 1;
                    const u32 d = (pw << 5) + b;
                    u32 c = org + d; if (c >= ncols) c -= ncols;
                    const VT v = emit_val<VT>(acc.get(d));
                    acc.clear(d);
                    f.colC[run] = c; f.valC[run] = v;
                    vmax = vmax > (u64)v ? vmax : (u64)v;
                    run++;
                }
            }
            if (tid == 0) { f.rpC[first + 1] = ubase + s_total; maxnnz = max(maxnnz, s_total); }
        } else if (cls == FZ_TINY) {
            if (tactive) {
                const bool tail = (ttails >> lane) & 1u;
                if (tail) {
                    const u64 pos = wb + __popc(ttails & ((1u << lane) - 1u));
                    f.colC[pos] = tkey; f.valC[pos] = tval;
                    vmax = vmax > (u64)tval ? vmax : (u64)tval;
                }
                if (lane == 0) { f.rpC[trow + 1] = wb + __popc(ttails); maxnnz = max(maxnnz, (u32)__popc(ttails)); }
            }
        } else {
            const u32 r = first + tid;
            if (r < end) { f.rpC[r + 1] = run + cnt; maxnnz = max(maxnnz, cnt); }
        }
        if (tid == 0) s_ticket[it & 1] = next_ticket;
        __syncthreads();                                                     // window clean, scan scratch free, next ticket visible
    }
    // ---- per-CTA maxima, then the last CTA out reports and leaves the control block zeroed for the next multiply
    vmax = warp_max_u64(vmax);
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) maxnnz = max(maxnnz, __shfl_xor_sync(0xFFFFFFFFu, maxnnz, m));
    if (lane == 0) { if (vmax) atomicMax(&f.ctrl->max_val_out, (ull)vmax); if (maxnnz) atomicMax(&f.ctrl->max_row_nnz, (ull)maxnnz); }
    if (f.finalize) {
        __syncthreads();
        if (tid == 0) { __threadfence(); s_last = atomicAdd(&f.ctrl->fused_done, 1u) == gridDim.x - 1 ? 1u : 0u; }
        __syncthreads();
        if (s_last) {
            __threadfence();
            if (f.host_mirror) fz_report(f.ctrl, f.host_mirror, f.epoch);
            if (tid == 0) *f.maxval_dst = *reinterpret_cast<volatile ull *>(&f.ctrl->max_val_out);
            __syncthreads();
            u32 *cw = reinterpret_cast<u32 *>(f.ctrl);
            for (u32 i = tid; i < sizeof(B200Ctrl) / 4; i += nt) cw[i] = 0;
        }
    }
}

// report + clean-up as a kernel of its own: multiplies whose "other" rows are computed after the placement kernel
__global__ void __launch_bounds__(64) k_fz_finalize(B200Ctrl *ctrl, u64 *host_mirror, u32 epoch, ull *maxval_dst) {
    if (host_mirror) fz_report(ctrl, host_mirror, epoch);
    if (threadIdx.x == 0) *maxval_dst = *reinterpret_cast<volatile ull *>(&ctrl->max_val_out);
    __syncthreads();
    u32 *cw = reinterpret_cast<u32 *>(ctrl);
    for (u32 i = threadIdx.x; i < sizeof(B200Ctrl) / 4; i += blockDim.x) cw[i] = 0;
}

// longest circular column span of any row (handles that did not come out of a multiply): thread per row
__global__ void __launch_bounds__(256) k_fz_row_span(u64 rows, u64 ncols, const u64 *__restrict__ rp, const u32 *__restrict__ col, u32 *out) {
    const long long n = (long long)ncols, half = n / 2;
    u32 m = 0;
    for (u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (u64)gridDim.x * blockDim.x) {
        const u64 s = rp[r], e = rp[r + 1];
        if (e <= s) continue;
        const long long ref = (long long)col[s];
        u32 dmin = 0xFFFFFFFFu, dmax = 0;
        for (u64 j = s; j < e; j++) {
            long long t = (long long)col[j] - ref + half;
            if (t < 0) t += n; else if (t >= n) t -= n;
            dmin = min(dmin, (u32)t); dmax = max(dmax, (u32)t);
        }
        m = max(m, dmax - dmin);
    }
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) m = max(m, __shfl_xor_sync(0xFFFFFFFFu, m, k));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

// =======================================================================================
// 3. host side
// =======================================================================================
int fz_row_span(b200_ctx *ctx, b200_csr *m) {
    // result lands in ctx->d_flag[26]; the caller reads it back with its own synchronising copy
    if (!m->rows || !m->nnz) return B200_OK;
    const u64 g = std::min<u64>((m->rows + 255) / 256, (u64)ctx->num_sms * 8);
    k_fz_row_span<<<(unsigned)g, 256, 0, ctx->stream>>>(m->rows, m->cols, m->d_rp, m->d_col, ctx->d_flag + 26);
    LAUNCH_CHECK(ctx);
    return B200_OK;
}

static int fz_ensure_scratch(b200_ctx *ctx, u64 rows) {
    const u64 tiles = rows / FZ_PRE_MIN_ROWS + 2;
    if (!ctx->d_fz || tiles > ctx->cap_ftile || rows + 1 > ctx->cap_fustat) {
        dfree(ctx, ctx->d_fz); ctx->d_fz = nullptr;
        const u64 ct = tiles + tiles / 8 + 64, cu = rows + rows / 8 + 1024;
        TRY(dmalloc(ctx, (void **)&ctx->d_fz, 512 + (ct + cu) * 8));
        ctx->d_fctrl = (B200Ctrl *)ctx->d_fz;
        ctx->d_ftile = (u64 *)(ctx->d_fz + 512);
        ctx->d_fustat = ctx->d_ftile + ct;
        ctx->cap_ftile = ct; ctx->cap_fustat = cu;
        ctx->f_dirty = true;
    }
    if (rows + 1 > ctx->cap_frows) {
        dfree(ctx, ctx->d_units); dfree(ctx, ctx->d_roworg); dfree(ctx, ctx->d_rowclass);
        ctx->d_units = nullptr; ctx->d_roworg = nullptr; ctx->d_rowclass = nullptr; ctx->cap_frows = 0;
        const u64 cap = rows + rows / 8 + 1024;
        TRY(dmalloc(ctx, (void **)&ctx->d_units, (cap + 1) * 4));
        TRY(dmalloc(ctx, (void **)&ctx->d_roworg, cap * 4));
        TRY(dmalloc(ctx, (void **)&ctx->d_rowclass, cap));
        ctx->cap_frows = cap;
    }
    if (ctx->f_dirty) {
        CUDA_TRY(cudaMemsetAsync(ctx->d_fz, 0, 512 + (ctx->cap_ftile + ctx->cap_fustat) * 8, ctx->stream));
        ctx->f_dirty = false;
    }
    return B200_OK;
}

// wait until every chunk of report slot `slot` carries `epoch`, reassemble the control block
static int fz_wait_report(b200_ctx *ctx, int slot, u32 epoch, B200Ctrl *out) {
    const u32 n = sizeof(B200Ctrl) / 4;
    volatile u64 *chunk = ctx->h_freport + (size_t)slot * n;
    u32 *o = reinterpret_cast<u32 *>(out);
    u64 spins = 0;
    for (u32 i = 0; i < n; i++) {
        u64 c;
        while ((u32)((c = chunk[i]) >> 32) != epoch) {
            if ((++spins & 0xFFF) == 0) {
                const cudaError_t q = cudaStreamQuery(ctx->stream);
                if (q == cudaErrorNotReady) { cudaGetLastError(); continue; }
                if (q != cudaSuccess) { ctx->f_dirty = true; return set_err(B200_ERR_CUDA, "multiply failed on the device: %s", cudaGetErrorString(q)); }
                if ((u32)(chunk[i] >> 32) != epoch) { ctx->f_dirty = true; return set_err(B200_ERR_CUDA, "the multiply finished without reporting (slot %d, epoch %u)", slot, epoch); }
            }
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        }
        o[i] = (u32)c;
    }
    __sync_synchronize();
    return B200_OK;
}

int resolve_pending(b200_ctx *ctx, const b200_csr *cm) {
    b200_csr *m = const_cast<b200_csr *>(cm);
    if (m->pending_slot < 0) return B200_OK;
    const int slot = m->pending_slot;
    B200Ctrl hc;
    const int r = fz_wait_report(ctx, slot, m->pending_epoch, &hc);
    m->pending_slot = -1;
    if (ctx->slot_owner[slot] == m) ctx->slot_owner[slot] = nullptr;
    if (r != B200_OK) return r;
    if (hc.error_flag) { ctx->f_dirty = true; return set_err(B200_ERR_CUDA, "a kernel of the multiply reported an impossible state (flag %u)", hc.error_flag); }
    m->nnz = hc.total_nnz; m->max_row_len = hc.max_row_nnz; m->h_maxval = hc.max_val_out; m->h_maxval_known = true;
    if (m->stats) {
        b200_stats *st = m->stats;
        st->nnz_c = m->nnz; st->products = hc.total_products; st->max_row_products = hc.max_row_products; st->max_row_nnz = hc.max_row_nnz;
        st->bytes_algorithmic += m->nnz * (u64)(4 + m->val_bits / 8);        // operands and row pointers were counted at launch
        st->sym_bin_rows[0] = hc.class_count[FZ_TINY]; st->sym_bin_rows[1] = hc.class_count[FZ_DENSE];
        st->sym_bin_rows[2] = hc.class_count[FZ_OTHER]; st->sym_bin_rows[3] = hc.class_count[FZ_EMPTY];
        for (int i = 0; i < 6; i++) st->sym_bin_rows[10 + i] = hc.sym_bin_count[B200_BIN_WIDE0 + i];
        st->sym_bin_rows[9] = hc.sym_bin_count[B200_BIN_HEAVY];
        if (ctx->timing && m->stats_timed) {
            if (cudaEventSynchronize(ctx->f_ev[slot][2]) == cudaSuccess) {
                cudaEventElapsedTime(&st->ms_symbolic, ctx->f_ev[slot][0], ctx->f_ev[slot][1]);
                cudaEventElapsedTime(&st->ms_numeric, ctx->f_ev[slot][1], ctx->f_ev[slot][2]);
                cudaEventElapsedTime(&st->ms_total, ctx->f_ev[slot][0], ctx->f_ev[slot][2]);
            } else cudaGetLastError();
        }
    }
    return B200_OK;
}

// one entry per instantiation of k_fz_numeric: [value width][accumulator mode][packed B]
struct FzKernel { const void *fn; int regs; size_t static_smem; };
static FzKernel g_fzk[2][3][2];
template <typename VT, int MODE, bool PACK>
static void fz_register(size_t optin) {
    FzKernel &k = g_fzk[sizeof(VT) == 8][MODE][PACK];
    k.fn = (const void *)k_fz_numeric<VT, MODE, PACK>;
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, k.fn) == cudaSuccess) { k.regs = fa.numRegs; k.static_smem = fa.sharedSizeBytes; } else { cudaGetLastError(); k.regs = 64; k.static_smem = 1024; }
    if (cudaFuncSetAttribute(k.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(optin - k.static_smem)) != cudaSuccess) cudaGetLastError();
}
void fz_setup(b200_ctx *ctx) {
    const size_t o = ctx->smem_optin;
    fz_register<u32, 0, false>(o); fz_register<u32, 0, true>(o); fz_register<u32, 1, false>(o); fz_register<u32, 1, true>(o);
    fz_register<u64, 0, false>(o); fz_register<u64, 0, true>(o); fz_register<u64, 1, false>(o); fz_register<u64, 1, true>(o);
    fz_register<u64, 2, false>(o); fz_register<u64, 2, true>(o);
}
// resident CTAs per SM (228 KB of shared memory per SM, 1 KB reserved per CTA; 64 K registers; 2048 threads)
static int fz_ctas_per_sm(const FzKernel &k, int threads, size_t smem) {
    const int by_smem = (int)((size_t)(228 * 1024) / (smem + k.static_smem + 1024));
    const int regs_per_cta = ((k.regs * 32 + 511) / 512 * 512) * (threads / 32);   // warp allocations are rounded to 512 registers
    const int by_regs = regs_per_cta ? 65536 / regs_per_cta : 32;
    return std::max(1, std::min(std::min(by_smem, by_regs), std::min(2048 / threads, 32)));
}

// C = A x B through the fused pipeline.  *handled = false (and nothing launched) when the multiply is better served by
// the binned pipeline of api.cu; the caller then runs that one.
template <typename VT>
int spgemm_fused(b200_ctx *ctx, const b200_csr *A, const b200_csr *B, b200_csr **out, b200_stats *st_out, bool *handled) {
    *handled = false;
    const b200_config &cfg = ctx->cfg;
    if (cfg.pipeline == 2) return B200_OK;
    const u64 rows = A->rows, ncols = B->cols;
    if (rows == 0 || A->nnz == 0 || B->nnz == 0 || B->nnz >= 0xFFFFFFFFull) return B200_OK;
    cudaStream_t s = ctx->stream;
    const u64 launches0 = ctx->launches;
    // ---- operand caches and host-known bounds
    TRY(ensure_desc(ctx, B));
    const bool packed = want_pack(ctx, B);
    if (packed) TRY(ensure_pack(ctx, B));
    TRY(host_maxval(ctx, A));
    TRY(host_maxval(ctx, B));
    const u64 maxA = A->h_maxval, maxB = B->h_maxval;
    const bool bpat = maxB == 1;
    const u64 p_bound = A->max_row_len * B->max_row_len;
    const int mode = pick_mode_bits(ctx, (int)sizeof(VT) * 8, p_bound, maxA, maxB);
    const size_t accb = mode == 0 ? 4 : 8;
    const int lg = pick_lg(ctx, B, 5);
    // ---- the dense window: per-row arcs when B is square with known offset bounds, else the whole column space
    const size_t smem_max = ctx->smem_optin - 2048;
    const u64 full = (ncols + 31) / 32 * 32;
    const u64 cap_cols_max = (u64)(smem_max / (accb * 8 + 1) * 8) / 32 * 32;       // acc + one bitmap bit per column
    u64 cap_cols_pref = (u64)((smem_max / 4 - 1024) / (accb * 8 + 1) * 8) / 32 * 32;   // four CTAs per SM
    if (cfg.fused_window_cols > 0) cap_cols_pref = std::min<u64>(cap_cols_max, ((u64)cfg.fused_window_cols + 31) / 32 * 32);
    int wmode = 0; u64 wcap = 0; u64 wneed = ~0ull;
    long long cs_lo = 0, cs_hi = 0;
    if (B->rows == B->cols && cfg.arc_window) {
        TRY(ensure_cs_bounds(ctx, B));
        if (B->cs_state == 1 && A->max_row_span < ncols) {
            cs_lo = B->cs_lo; cs_hi = B->cs_hi;
            wneed = (A->max_row_span + (u64)(cs_hi - cs_lo) + 1 + 31 + 31) / 32 * 32;
        }
    }
    if (wneed < full && wneed <= cap_cols_max) { wmode = 1; wcap = wneed; }          // every row's arc fits
    else if (full <= cap_cols_max && (full <= cap_cols_pref || wneed == ~0ull || wneed >= full)) { wmode = 0; wcap = full; }
    else if (wneed != ~0ull) { wmode = 1; wcap = std::min(cap_cols_pref, full); }     // rows are tested one by one
    else if (full <= cap_cols_max) { wmode = 0; wcap = full; }
    else { wmode = 0; wcap = 0; }                                                     // no dense class: tiny + counted rows only
    const u64 dense_pmax = cfg.fused_dense_pmax > 0 ? (u64)cfg.fused_dense_pmax : 65536;
    const bool all_fit = wcap != 0 && (wmode == 0 || wneed <= wcap);
    const bool guaranteed = all_fit && p_bound <= dense_pmax;                          // no "other" rows can occur
    const double meanP = ((double)A->nnz / (double)rows) * ((double)B->nnz / (double)B->rows);
    if (cfg.pipeline == 0 && !guaranteed && !(p_bound <= 32 || meanP <= 24.0)) return B200_OK;   // mostly hash / heavy rows: binned pipeline
    // ---- C's capacity from a host-known bound when that is cheap, else from the pre-pass's exact sum of min(P_i, cols)
    const size_t esz = 4 + sizeof(VT);
    unsigned __int128 hb128 = (unsigned __int128)A->nnz * B->max_row_len;
    const unsigned __int128 dense128 = (unsigned __int128)rows * ncols;
    if (dense128 < hb128) hb128 = dense128;
    const bool cheap_bound = hb128 * esz <= (unsigned __int128)(ctx->total_mem / 16);
    const bool need_pre_report = !guaranteed || !cheap_bound;

    TRY(ensure_row_scratch(ctx, rows));
    TRY(fz_ensure_scratch(ctx, rows));
    // report slot of this multiply; a slot still owned by an unread product is read first
    const u32 epoch = ++ctx->fepoch ? ctx->fepoch : ++ctx->fepoch;
    const int slot = (int)(epoch % B200_REPORT_SLOTS);
    if (ctx->slot_owner[slot]) TRY(resolve_pending(ctx, ctx->slot_owner[slot]));
    u64 *mirror = ctx->h_freport + (size_t)slot * (sizeof(B200Ctrl) / 4);
    const bool timing = ctx->timing;
    if (timing) cudaEventRecord(ctx->f_ev[slot][0], s);

    const int threads = cfg.fused_threads > 0 ? cfg.fused_threads : 256;
    const int nwarps = threads / 32;
    // ---- pre-pass
    {
        const double avgA = (double)A->nnz / (double)rows;
        const int G = avgA <= 2.0 ? 1 : avgA <= 6.0 ? 4 : avgA <= 24.0 ? 8 : 32;
        const u64 tile_rows = FZ_PRE_ROWS(G);
        const u64 tiles = (rows + tile_rows - 1) / tile_rows;
        FzPreArgs pa;
        pa.rows = rows; pa.ncols = ncols; pa.rpA = A->d_rp; pa.colA = A->d_col; pa.bdesc = B->d_desc;
        pa.wcap = (u32)wcap; pa.cs_lo = cs_lo; pa.cs_hi = cs_hi; pa.dense_pmax = dense_pmax;
        pa.tiny_run = (u32)nwarps; pa.other_run = (u32)threads;
        pa.rowclass = ctx->d_rowclass; pa.roworg = ctx->d_roworg; pa.units = ctx->d_units;
        pa.tile_status = ctx->d_ftile; pa.unit_status = ctx->d_fustat;
        pa.ctrl = ctx->d_fctrl; pa.bin_rows = ctx->d_bin_rows; pa.bin_stride = (u32)ctx->cap_rows;
        pa.host_mirror = need_pre_report ? mirror : nullptr; pa.epoch = epoch;
#define FZ_PRE(GG) do { if (wmode == 1) k_fz_prepass<GG, 1><<<(unsigned)tiles, 256, 0, s>>>(pa); else k_fz_prepass<GG, 0><<<(unsigned)tiles, 256, 0, s>>>(pa); } while (0)
        if (G == 1) FZ_PRE(1); else if (G == 4) FZ_PRE(4); else if (G == 8) FZ_PRE(8); else FZ_PRE(32);
#undef FZ_PRE
        ctx->f_dirty = true;                                                  // until the numeric kernel has cleaned up behind us
        LAUNCH_CHECK(ctx);
        if (timing) cudaEventRecord(ctx->f_ev[slot][1], s);
        // ---- everything below needs the pre-pass's tile count to zero its status words
        u64 cap_entries = (u64)hb128;
        B200Ctrl hc; memset(&hc, 0, sizeof(hc));
        bool others = false;
        if (need_pre_report) {
            TRY(fz_wait_report(ctx, slot, epoch, &hc));
            cap_entries = hc.total_bound;
            others = hc.class_count[FZ_OTHER] != 0;
        }
        b200_csr *C = nullptr;
        TRY(csr_alloc(ctx, rows, ncols, 0, A->val_bits, false, &C));
        C->cap_entries = std::max<u64>(cap_entries, 1);
        int r = alloc_entries(ctx, C);
        if (r != B200_OK) { b200_csr_free(ctx, C); return r; }
        // a product's columns stay inside the arc (column range of A) + (offsets of B), and its rows' spans grow by B's
        C->cr_start = 0; C->cr_len = ncols;
        {
            const u64 sb = A->max_row_span + (u64)(cs_hi - cs_lo);            // a row's arc grows by B's offset range
            C->max_row_span = wneed != ~0ull && sb <= ncols / 2 ? sb : ncols;
        }
        Fan fan(ctx);
        if (others) {
            r = legacy_counts(ctx, A, B, ctx->d_fctrl, p_bound, lg, fan);
            fan.join();
            if (r != B200_OK) { b200_csr_free(ctx, C); return r; }
        }
        // ---- numeric + placement
        FzNumArgs<VT> fa;
        fa.a = NumArgs<VT>{A->d_rp, A->d_col, (const VT *)A->d_val, B->d_desc, B->d_col, (const VT *)B->d_val};
        fa.pack = B->d_pack; fa.units = ctx->d_units; fa.rowclass = ctx->d_rowclass; fa.roworg = ctx->d_roworg; fa.nnz_row = ctx->d_nnz_row;
        fa.unit_status = ctx->d_fustat; fa.tile_status = ctx->d_ftile; fa.n_tile_status = tiles;
        fa.ctrl = ctx->d_fctrl; fa.rpC = C->d_rp; fa.colC = C->d_col; fa.valC = (VT *)C->d_val;
        fa.ncols = (u32)ncols; fa.wcap = (u32)wcap; fa.nw = (u32)(wcap / 32);
        fa.wmode = wmode; fa.bpat = bpat ? 1 : 0; fa.lg = std::min(lg, 5); fa.narrow = ncols < (1ull << 27) ? 1 : 0;
        fa.finalize = others ? 0 : 1; fa.maxval_dst = C->d_maxval; fa.host_mirror = mirror; fa.epoch = need_pre_report ? epoch + 0x80000000u : epoch;
        const u32 final_epoch = fa.epoch;
        const size_t smem = (size_t)wcap * accb + (size_t)wcap / 8 + 16;
        const int m2 = sizeof(VT) == 8 ? mode : std::min(mode, 1);
        const FzKernel &fk = g_fzk[sizeof(VT) == 8][m2][packed ? 1 : 0];
        const int grid = (int)std::min<u64>(rows, (u64)ctx->num_sms * fz_ctas_per_sm(fk, threads, smem));
        void *kargs[] = {(void *)&fa};
        const cudaError_t le = cudaLaunchKernel(fk.fn, dim3(grid), dim3(threads), kargs, smem, s);
        ctx->launches++;
        if (le != cudaSuccess) { b200_csr_free(ctx, C); return set_err(B200_ERR_CUDA, "fused numeric kernel launch failed: %s", cudaGetErrorString(le)); }
        if (others) {
            r = legacy_numeric(ctx, A, B, ctx->d_fctrl, C, p_bound, std::min<u64>(p_bound, ncols), mode, packed, bpat, lg, fan);
            fan.join();
            if (r != B200_OK) { b200_csr_free(ctx, C); return r; }
            k_fz_finalize<<<1, 64, 0, s>>>(ctx->d_fctrl, mirror, final_epoch, C->d_maxval);
            LAUNCH_CHECK(ctx);
        }
        ctx->f_dirty = false;                                                 // the kernels queued above leave the buffers clean
        if (timing) cudaEventRecord(ctx->f_ev[slot][2], s);
        // ---- hand the product out; its size follows through the report
        C->pending_slot = slot; C->pending_epoch = final_epoch; ctx->slot_owner[slot] = C;
        C->est_nnz = std::max<u64>(1, (u64)(meanP * (double)rows / 1.6));
        C->nnz = 0; C->max_row_len = std::min<u64>(p_bound, ncols);
        b200_stats *st = new b200_stats();
        memset(st, 0, sizeof(*st));
        st->rows = rows; st->cols = ncols; st->nnz_a = A->nnz; st->nnz_b = B->nnz; st->acc_mode = mode; st->pipeline = 1;
        st->kernel_launches = (int32_t)(ctx->launches - launches0);
        st->bytes_algorithmic = (A->nnz + B->nnz) * (u64)esz + (A->rows + B->rows + rows + 3) * 8;
        C->stats = st; C->stats_timed = timing;
        *out = C; *handled = true;
        if (st_out) { TRY(resolve_pending(ctx, C)); *st_out = *C->stats; }
    }
    return B200_OK;
}

template int spgemm_fused<u32>(b200_ctx *, const b200_csr *, const b200_csr *, b200_csr **, b200_stats *, bool *);
template int spgemm_fused<u64>(b200_ctx *, const b200_csr *, const b200_csr *, b200_csr **, b200_stats *, bool *);
