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

struct FzScan { u64 *st, *gagg, *gpre; };   // two-level look-back state (see fz_lookback)

struct FzPreArgs {
    u64 rows, ncols;
    const u64 *rpA; const u32 *colA; const uint2 *bdesc;
    u32 wcap;                  // dense window capacity in columns (multiple of 32); 0: no dense class
    long long cs_lo, cs_hi;    // WMODE 1: every entry (k, c) of B has c - k in [cs_lo, cs_hi] (circularly)
    u64 dense_pmax;            // rows with more products than this are never dense (they would pin one CTA for too long)
    u32 tiny_run, other_run;   // longest run of tiny / placed-only rows that forms one unit
    unsigned char *rowclass; uint4 *rowwin; u32 *units;    // rowwin = {origin column, bitmap words, products} of a dense row
    FzScan tile_scan, unit_scan;
    B200Ctrl *ctrl; u32 *bin_rows; u32 bin_stride;
    u64 *host_mirror; u32 epoch;
};

// Two-level decoupled look-back.  With several hundred units in flight a one-level look-back (32 predecessors per step,
// each step an L2 round trip) advances the chain of known prefixes by 32 units per step and the placement serialises
// (measured: 1.5 ms for the 25 696 rows of A^7).  Here units publish their count (a) in their own status word and (b)
// into the accumulator of their group of 32 units (one 64-bit atomicAdd: count of publishers in the top 16 bits, sum
// below); a unit then needs the <= 31 status words before it in its own group plus a look-back over GROUP words, which
// moves 32 groups = 1024 units per step.  Whoever finishes a look-back publishes the inclusive prefix of the group before
// its own, which is where later look-backs stop.
#define FZ_GROUP_LG 5
#define FZ_GROUP (1u << FZ_GROUP_LG)
#define FZ_GSUM_BITS 48
__device__ __forceinline__ void fz_publish(const FzScan &sc, u32 idx, u64 agg) {          // one thread
    atomicExch((ull *)&sc.st[idx], (ull)(SCAN_FLAG_AGG | agg));
    atomicAdd((ull *)&sc.gagg[idx >> FZ_GROUP_LG], (ull)((1ull << FZ_GSUM_BITS) + agg));
}
// one warp, one attempt: sum of everything published before idx -> excl; false when a needed word is not there yet
__device__ __forceinline__ bool fz_try(const FzScan &sc, u32 idx, int lane, u64 &excl) {
    const u32 g = idx >> FZ_GROUP_LG, j = idx & (FZ_GROUP - 1);
    // predecessors inside my group
    u64 st = SCAN_FLAG_AGG;
    if ((u32)lane < j) st = ld_volatile_u64(&sc.st[(g << FZ_GROUP_LG) + lane]);
    if (__any_sync(0xFFFFFFFFu, (st >> 62) == 0)) return false;
    const u64 s1 = warp_sum_u64((u32)lane < j ? (st & SCAN_VAL_MASK) : 0ull);
    // groups before mine
    u64 s2 = 0;
    if (g > 0) {
        long long look = (long long)g - 1;
        while (true) {
            const long long i = look - lane;
            u64 val = 0; int kind = 2;                                   // 2: inclusive prefix, 1: complete group sum, 0: not there yet
            if (i >= 0) {
                const u64 p = ld_volatile_u64(&sc.gpre[i]);
                if (p >> 62) val = p & SCAN_VAL_MASK;
                else {
                    const u64 a = ld_volatile_u64(&sc.gagg[i]);
                    if ((a >> FZ_GSUM_BITS) == FZ_GROUP) { val = a & ((1ull << FZ_GSUM_BITS) - 1); kind = 1; }
                    else kind = 0;
                }
            }
            const u32 pre_mask = __ballot_sync(0xFFFFFFFFu, kind == 2);
            const u32 miss_mask = __ballot_sync(0xFFFFFFFFu, kind == 0);
            const int first = pre_mask ? __ffs(pre_mask) - 1 : 32;      // nearest group whose inclusive prefix is known
            const u32 need = first >= 31 ? 0xFFFFFFFFu : ((2u << first) - 1u);
            if (miss_mask & need) return false;
            s2 += warp_sum_u64(lane <= first ? val : 0ull);
            if (pre_mask) break;
            look -= 32;
        }
        if (lane == 0) atomicExch((ull *)&sc.gpre[g - 1], (ull)(SCAN_FLAG_PRE | s2));
    }
    excl = s1 + s2;
    return true;
}
// one warp: the same, waiting (with a back-off, so that the polling does not eat the issue slots of working warps)
__device__ __forceinline__ u64 fz_wait(const FzScan &sc, u32 idx, int lane) {
    u64 excl = 0;
    while (!fz_try(sc, idx, lane, excl)) __nanosleep(100);
    return excl;
}
__device__ __forceinline__ u64 fz_lookback(const FzScan &sc, u32 idx, u64 agg, int lane) {
    if (lane == 0) fz_publish(sc, idx, agg);
    return fz_wait(sc, idx, lane);
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
//     short arc.  The origin is moved back by < 128 columns so that the wrap point (column n -> 0) falls on the edge of a group of four bitmap
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
                        bool fits = false; u32 org = 0, words = (u32)(((p.ncols + 127) >> 7) << 2);   // whole column space unless measured
                        if (WMODE == 0) fits = p.wcap != 0;
                        else if (p.wcap) {
                            const long long lo = (long long)dmin + p.cs_lo, hi = (long long)dmax + p.cs_hi;
                            long long width = hi - lo + 1;
                            long long oc = (ref - half + lo) % n; if (oc < 0) oc += n;
                            const long long delta = ((oc % 128) - (n % 128) + 128) % 128;   // (n - origin) becomes a multiple of 128
                            if (oc >= delta) { oc -= delta; width += delta; }
                            else { width += oc; oc = 0; }                               // origin 0: the window [0, width) needs no rotation
                            if (width >= n) { oc = 0; width = n; }                      // the arc closes on itself: the whole circle
                            org = (u32)oc; words = (u32)(((width + 127) >> 7) << 2);    // windows are cut in groups of 128 columns
                            fits = (u64)words * 32 <= (u64)p.wcap;
                        }
                        if (fits && pr <= p.dense_pmax) { cls = FZ_DENSE; p.rowwin[row] = make_uint4(org, words, (u32)pr, 0u); }
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
        const u64 excl = fz_lookback(p.tile_scan, tile, (u64)agg, lane);
        if (lane == 0) s_ubase = (u32)excl;
    }
    __syncthreads();
    const u64 r = (u64)tile * TILE + tid;
    if (tid < TILE && r < p.rows) {
        if (start) p.units[s_ubase + wbase + incl - 1] = (u32)r;
        const u32 bl = s_binloc[tid], b = bl >> 24;
        if (b != B200_BIN_NONE) p.bin_rows[(u64)b * p.bin_stride + s_base[b] + (bl & 0xFFFFFFu)] = (u32)r;
        p.unit_scan.st[r] = 0;                                               // the numeric kernel's look-back starts clean
        if ((r & (FZ_GROUP - 1)) == 0) { p.unit_scan.gagg[r >> FZ_GROUP_LG] = 0; p.unit_scan.gpre[r >> FZ_GROUP_LG] = 0; }
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
    const u32 *units; const unsigned char *rowclass; const uint4 *rowwin; const u32 *nnz_row;
    FzScan unit_scan, tile_scan; u64 n_tile_status;
    B200Ctrl *ctrl;
    u64 *rpC; u32 *colC; VT *valC;
    u32 ncols, nw;             // bitmap capacity in 32-column words (a dense row's window never has more)
    u32 ncap;                  // accumulator / column ring slots in shared memory (power of two)
    u32 pcap;                  // product buffer entries in shared memory
    u64 *spill_acc; u32 *spill_col; u64 spill_stride;   // per-CTA global slots for ranks >= ncap (zeroed, left zeroed)
    int wmode, bpat, lg, narrow, finalize;
    u64 *host_mirror; u32 epoch;
    ull *maxval_dst;           // the product handle's largest-value scalar
    ull *prof;                 // developer probe (B200_FZ_PROF=1): cycles of thread 0 per phase, summed over CTAs; else null
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
// accumulate into a global (spill) slot: plain 64-bit add while no overflow is possible, saturating CAS in mode 2
template <int MODE>
__device__ __forceinline__ void fz_spill_add(u64 *slot, u64 x) {
    if (MODE != 2) { atomicAdd((ull *)slot, (ull)x); return; }
    ull old = *reinterpret_cast<volatile ull *>(slot), assumed;
    do {
        assumed = old;
        ull sum = assumed + x; if (sum < assumed) sum = ~0ull;
        if (sum == assumed) break;
        old = atomicCAS((ull *)slot, assumed, sum);
    } while (old != assumed);
}

// storage type of one product's value in the product buffer
template <int MODE, typename VT> struct PVal { typedef u64 type; };
template <typename VT> struct PVal<0, VT> { typedef u32 type; };
template <> struct PVal<1, u32> { typedef u32 type; };

#define FZ_KEEP 2   // A entries per thread whose value and B record stay in registers between the two product passes
#define FZ_QUEUE 8  // finished units a CTA may hold back while their offsets are on the way (power of two)

// tiny_gather (devutil.cuh) with the B row descriptors already in registers (they are fetched one unit ahead here)
template <typename VT>
__device__ __forceinline__ u32 fz_tiny_gather(u32 dA, u32 bstart, u32 deg, VT a, const u32 *colB, const VT *valB, int lane, u32 &key, VT &val, bool bpat) {
    u32 incl = deg;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { u32 t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
    const u32 excl = incl - deg;
    const u32 P = __shfl_sync(0xFFFFFFFFu, incl, 31);
    int lo = 0;                                                             // lane p: last entry e with excl_e <= p
#pragma unroll
    for (int step = 16; step > 0; step >>= 1) {
        const int cand = lo + step;
        const u32 t = __shfl_sync(0xFFFFFFFFu, excl, cand & 31);
        if (cand < (int)dA && t <= (u32)lane) lo = cand;
    }
    const u32 e_excl = __shfl_sync(0xFFFFFFFFu, excl, lo);
    const u32 e_bstart = __shfl_sync(0xFFFFFFFFu, bstart, lo);
    const VT e_a = shfl_any(a, lo);
    key = B200_EMPTY_KEY; val = 0;
    if ((u32)lane < P) {
        const u32 j = e_bstart + ((u32)lane - e_excl);
        key = colB[j];
        val = bpat ? e_a : sat_mul(e_a, valB[j]);
    }
    return P;
}

// Dense rows: mark -> rank -> accumulate at rank -> (later) emit.
//   mark   every intermediate product sets its column's bit in the row's window bitmap (shared memory, relative to the
//          window origin the pre-pass chose);
//   rank   exclusive prefix popcount over the window's words, taken in ascending-column order (the walk starts at the
//          word where the window wraps from column n-1 to 0; windows are cut in 128-column groups, so this is four words
//          per 128-bit access); the total is the row's length and is published to the look-back straight away;
//   accum  the products are generated a second time -- from the A values and B records still in registers -- and each is
//          added into acc[rank(column)]: compact accumulators of the row's LENGTH, not of its window; the column is
//          dropped into cols[rank];
//   emit   thread per output entry, coalesced stores, sorted by construction.
// Emission is DECOUPLED from the rest: rows have to be placed in row order, and a CTA that waited for its row's offset
// before going on would run at the pace of the slowest row in flight (measured: 45 % of all issued instructions were
// look-back polls, the barrier behind them the top stall).  The accumulators therefore form a RING: a finished row stays
// where it was accumulated, the CTA goes on with its next unit behind it in the ring, and pending rows are drained -- in
// order, one look-back attempt each, its loads issued a phase before their result is needed -- whenever their offset has
// arrived; a CTA only waits when the ring or its queue of pending units is full, and at the very end.
// The two product passes are not equal: the first runs one A entry per lane (the lanes of a warp see B rows of different
// lengths, about half of the slots are idle), so it also writes every product as a {window offset, value} pair into a
// shared-memory product buffer, and the second pass -- the expensive one: rank lookup, accumulate, column store -- runs
// one PRODUCT per lane from that buffer.  Rows with more products than the buffer holds generate their products twice.
template <typename VT, int MODE, bool PACK>
__global__ void __launch_bounds__(256) k_fz_numeric(FzNumArgs<VT> f) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ u32 s_ticket[4], s_wsum[8], s_P;
    __shared__ u64 s_excl;
    __shared__ u32 s_ok, s_last, s_ok1;
    __shared__ u64 s_excl1;
    __shared__ u32 q_unit[FZ_QUEUE], q_first[FZ_QUEUE], q_rows[FZ_QUEUE], q_base[FZ_QUEUE], q_nnz[FZ_QUEUE], q_kind[FZ_QUEUE], q_cnt[FZ_QUEUE][8];
    const u32 nt = blockDim.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nwarps = nt >> 5;
    // layout: window bits | window prefixes | accumulator ring | column ring | product buffer (offsets, then values)
    u32 *bits = reinterpret_cast<u32 *>(smem_raw);
    u32 *pre = bits + f.nw;
    uint4 *bits4 = reinterpret_cast<uint4 *>(bits), *pre4 = reinterpret_cast<uint4 *>(pre);
    const u32 R = f.ncap, rmask = R - 1;
    Acc<MODE> acc; acc.bind(smem_raw + (size_t)f.nw * 8, R);
    u32 *cols = reinterpret_cast<u32 *>(smem_raw + (size_t)f.nw * 8 + Acc<MODE>::bytes(R));
    typedef typename PVal<MODE, VT>::type PV;                              // a product's value: 32 bits when the sums are (mode 0, u32 values)
    u32 *pd = cols + R;
    PV *px = reinterpret_cast<PV *>(pd + f.pcap);
    const u32 pcap = f.pcap;
    u64 *sp_acc = f.spill_acc + (u64)blockIdx.x * f.spill_stride;
    u32 *sp_col = f.spill_col + (u64)blockIdx.x * f.spill_stride;
    for (u32 t = tid; t < f.nw; t += nt) bits[t] = 0;
    for (u32 t = tid; t < R; t += nt) acc.clear(t);
    // the pre-pass is done with its look-back words: leave them zeroed for the next multiply
    for (u64 i = (u64)blockIdx.x * nt + tid; i < f.n_tile_status; i += (u64)gridDim.x * nt) {
        f.tile_scan.st[i] = 0;
        if ((i & (FZ_GROUP - 1)) == 0) { f.tile_scan.gagg[i >> FZ_GROUP_LG] = 0; f.tile_scan.gpre[i >> FZ_GROUP_LG] = 0; }
    }
    if (tid < 3) s_ticket[tid] = atomicAdd(&f.ctrl->unit_ticket, 1u);       // (each CTA's three tickets need not be consecutive)
    __syncthreads();
    if (tid == 0) {                                                         // ... but they are worked on in ascending order
        u32 a = s_ticket[0], b = s_ticket[1], c = s_ticket[2];
        if (a > b) { const u32 t = a; a = b; b = t; }
        if (b > c) { const u32 t = b; b = c; c = t; }
        if (a > b) { const u32 t = a; a = b; b = t; }
        s_ticket[0] = a; s_ticket[1] = b; s_ticket[2] = c; s_P = 0;
    }
    __syncthreads();
    const u32 num_units = f.ctrl->num_units;
    const u32 ncols = f.ncols;
    const bool bpat = f.bpat != 0;
    u64 vmax = 0; u32 maxnnz = 0;
    u32 qh = 0, qn = 0, ring_next = 0, ring_used = 0;                       // the same on every thread
    long long pacc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; long long tlast = f.prof ? clock64() : 0;
#define FZ_TICK(i) do { if (f.prof && tid == 0) { const long long now_ = clock64(); pacc[i] += now_ - tlast; tlast = now_; } } while (0)

    // ---- emission of the oldest pending unit at offset `off` (its look-back has succeeded)
    auto emit_oldest = [&](u64 off) {
        const u32 qi = qh & (FZ_QUEUE - 1);
        const u32 du = q_unit[qi], kind = q_kind[qi], dfirst = q_first[qi], dn = q_nnz[qi], dbase = q_base[qi], drows = q_rows[qi];
        if (kind != FZ_OTHER) {
            for (u32 t = tid; t < dn; t += nt) {
                const u32 idx = (dbase + t) & rmask;
                const u32 c = cols[idx];
                const VT v = emit_val<VT>(acc.get(idx));
                acc.clear(idx);
                f.colC[off + t] = c; f.valC[off + t] = v;
                vmax = vmax > (u64)v ? vmax : (u64)v;
            }
            ring_used -= dn;
        }
        if (kind == FZ_DENSE) {
            if (tid == 0) { f.rpC[dfirst + 1] = off + dn; maxnnz = max(maxnnz, dn); }
        } else if (kind == FZ_TINY) {
            if (tid < drows) {
                u32 p = 0;
                for (u32 i = 0; i <= tid; i++) p += q_cnt[qi][i];
                f.rpC[dfirst + tid + 1] = off + p; maxnnz = max(maxnnz, q_cnt[qi][tid]);
            }
        } else {                                                             // rows counted beforehand (or empty): only placed
            const u32 r = dfirst + tid;
            u32 cnt = 0;
            if (tid < drows && f.rowclass[r] == FZ_OTHER) cnt = f.nnz_row[r];
            u32 incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const u32 t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= (u32)d) incl += t; }
            if (lane == 31) s_wsum[w] = incl;
            __syncthreads();
            u32 wbase = 0;
            for (u32 i = 0; i < w; i++) wbase += s_wsum[i];
            if (tid < drows) { f.rpC[r + 1] = off + wbase + incl; maxnnz = max(maxnnz, cnt); }
        }
        if (tid == 0) { if (du == 0) f.rpC[0] = 0; if (du + 1 == num_units) f.ctrl->total_nnz = off + dn; }
        qh++; qn--;
    };
    // blocking (or one-shot) placement of the oldest pending unit: the slow path (ring / queue full, kernel end)
    auto drain_one = [&](bool blocking) -> bool {
        const u32 du = q_unit[qh & (FZ_QUEUE - 1)];
        if (w == 0) {
            u64 excl = 0; bool ok = true;
            if (blocking) excl = fz_wait(f.unit_scan, du, (int)lane); else ok = fz_try(f.unit_scan, du, (int)lane, excl);
            if (lane == 0) { s_excl = excl; s_ok = ok ? 1u : 0u; }
        }
        __syncthreads();
        const bool ok = s_ok != 0;
        if (ok) emit_oldest(s_excl);
        __syncthreads();
        return ok;
    };
    auto push = [&](u32 u, u32 first, u32 nrows, u32 nnz, u32 kind, const u32 *rowcnt) {   // thread 0 fills the slot; every thread advances the counters
        if (tid == 0) {
            const u32 qi = (qh + qn) & (FZ_QUEUE - 1);
            q_unit[qi] = u; q_first[qi] = first; q_rows[qi] = nrows; q_base[qi] = ring_next; q_nnz[qi] = nnz; q_kind[qi] = kind;
            if (kind == FZ_TINY) {
#pragma unroll
                for (u32 i = 0; i < 8; i++) q_cnt[qi][i] = rowcnt[i];
            }
        }
        qn++;
        if (kind != FZ_OTHER) { ring_next = (ring_next + nnz) & rmask; ring_used += nnz; }
    };

    // ---- unit pipeline registers: c_* the unit being processed, n_* the next one, nn_* the one after
    u32 c_first = 0, c_end = 0, c_cls = FZ_EMPTY, c_len = 0; uint4 c_win = make_uint4(0u, 0u, 0u, 0u); u64 c_rs = 0;
    u32 c_k[FZ_KEEP]; VT c_av[FZ_KEEP]; PackRec c_rec[FZ_KEEP];
    u32 n_first = 0, n_end = 0, n_cls = FZ_EMPTY, n_len = 0; uint4 n_win = make_uint4(0u, 0u, 0u, 0u); u64 n_rs = 0;
    u64 n_rsD = 0, n_reD = 0, n_rsT = 0, n_reT = 0;
    u32 n_k[FZ_KEEP]; VT n_av[FZ_KEEP]; PackRec n_rec[FZ_KEEP];
    u32 nn_first = 0, nn_end = 0;
#pragma unroll
    for (int e = 0; e < FZ_KEEP; e++) { c_k[e] = 0; c_av[e] = 0; c_rec[e].a = make_uint4(0, 0, 0, 0); c_rec[e].b = c_rec[e].a; n_k[e] = 0; n_av[e] = 0; n_rec[e] = c_rec[e]; }
    // level 1: the unit list entry
    auto lvl1 = [&](u32 u, u32 &first, u32 &end) { first = 0; end = 0; if (u < num_units) { first = f.units[u]; end = f.units[u + 1]; } };
    // level 2: class, window, row pointers of the unit's first row (dense) and of this warp's row (tiny)
    auto lvl2 = [&]() {
        n_cls = FZ_EMPTY; n_rsD = n_reD = n_rsT = n_reT = 0; n_win = make_uint4(0u, 0u, 0u, 0u);
        if (n_first < n_end) {
            n_cls = f.rowclass[n_first];
            n_win = f.rowwin[n_first];                                       // (only meaningful for dense rows)
            n_rsD = f.a.rpA[n_first]; n_reD = f.a.rpA[n_first + 1];
            if (n_first + w < n_end) { n_rsT = f.a.rpA[n_first + w]; n_reT = f.a.rpA[n_first + w + 1]; }
        }
    };
    // level 3: the A entries this thread keeps (dense) / this warp's tiny row
    auto lvl3 = [&]() {
#pragma unroll
        for (int e = 0; e < FZ_KEEP; e++) { n_k[e] = 0; n_av[e] = 0; }
        if (n_cls == FZ_DENSE) {
            n_rs = n_rsD; n_len = (u32)(n_reD - n_rsD);
#pragma unroll
            for (int e = 0; e < FZ_KEEP; e++) {
                const u32 t = tid + e * nt;
                if (t < n_len) { n_k[e] = f.a.colA[n_rs + t]; n_av[e] = f.a.valA[n_rs + t]; }
            }
        } else if (n_cls == FZ_TINY) {
            n_rs = n_rsT; n_len = (u32)(n_reT - n_rsT);
            if (lane < n_len) { n_k[0] = f.a.colA[n_rs + lane]; n_av[0] = f.a.valA[n_rs + lane]; }
        } else { n_rs = 0; n_len = 0; }
    };
    // level 4: B records (dense, packed B) / B row descriptors (tiny)
    auto lvl4 = [&]() {
#pragma unroll
        for (int e = 0; e < FZ_KEEP; e++) { n_rec[e].a = make_uint4(0, 0, 0, 0); n_rec[e].b = n_rec[e].a; }
        if (n_cls == FZ_DENSE) {
            if constexpr (PACK) {
#pragma unroll
                for (int e = 0; e < FZ_KEEP; e++) if (tid + e * nt < n_len) n_rec[e] = load_pack(f.pack, n_k[e]);
            }
        } else if (n_cls == FZ_TINY) {
            if (lane < n_len) { const uint2 d = f.a.bdesc[n_k[0]]; n_rec[0].a.x = d.x; n_rec[0].a.y = d.y; }
        }
    };
    auto rotate = [&]() {
        c_first = n_first; c_end = n_end; c_cls = n_cls; c_len = n_len; c_win = n_win; c_rs = n_rs;
#pragma unroll
        for (int e = 0; e < FZ_KEEP; e++) { c_k[e] = n_k[e]; c_av[e] = n_av[e]; c_rec[e] = n_rec[e]; }
        n_first = nn_first; n_end = nn_end;
    };
    // prologue: unit 0 of this CTA fully loaded, unit 1's list entry
    lvl1(s_ticket[0], n_first, n_end); lvl2(); lvl3(); lvl4();
    lvl1(s_ticket[1], nn_first, nn_end);
    rotate();

    // ---- look-back attempt for the oldest pending unit, split in two so that its loads fly during a compute phase:
    //      issue (warp 0) ... evaluate (warp 0) -> s_ok / s_excl, read by everybody after the next barrier
    u64 t_st = 0, t_pre = 0, t_agg = 0; u32 t_unit = 0; bool t_armed = false;
    auto try_issue = [&]() {
        t_armed = qn > 0 && q_kind[qh & (FZ_QUEUE - 1)] != FZ_OTHER;         // (placed-only units go through the slow path: their emit has a barrier)
        if (t_armed && w == 0) {
            t_unit = q_unit[qh & (FZ_QUEUE - 1)];
            const u32 g = t_unit >> FZ_GROUP_LG, j = t_unit & (FZ_GROUP - 1);
            t_st = SCAN_FLAG_AGG; t_pre = SCAN_FLAG_PRE; t_agg = 0;
            if (lane < j) t_st = ld_volatile_u64(&f.unit_scan.st[(g << FZ_GROUP_LG) + lane]);
            if (g > lane) { t_pre = ld_volatile_u64(&f.unit_scan.gpre[g - 1 - lane]); t_agg = ld_volatile_u64(&f.unit_scan.gagg[g - 1 - lane]); }
        }
    };
    auto try_eval = [&]() {                                                  // one step of the look-back from the loaded words; no retry
        if (t_armed && w == 0) {
            const u32 g = t_unit >> FZ_GROUP_LG, j = t_unit & (FZ_GROUP - 1);
            bool ok = !__any_sync(0xFFFFFFFFu, (t_st >> 62) == 0);
            const u64 s1 = warp_sum_u64(lane < j ? (t_st & SCAN_VAL_MASK) : 0ull);
            u64 s2 = 0;
            if (g > 0) {
                int kind = 2; u64 val = 0;                                   // lanes beyond group 0 count as "prefix 0"
                if (g > lane) {
                    if (t_pre >> 62) val = t_pre & SCAN_VAL_MASK;
                    else if ((t_agg >> FZ_GSUM_BITS) == FZ_GROUP) { val = t_agg & ((1ull << FZ_GSUM_BITS) - 1); kind = 1; }
                    else kind = 0;
                }
                const u32 pre_mask = __ballot_sync(0xFFFFFFFFu, kind == 2);
                const u32 miss_mask = __ballot_sync(0xFFFFFFFFu, kind == 0);
                const int first = pre_mask ? __ffs(pre_mask) - 1 : 32;
                const u32 need = first >= 31 ? 0xFFFFFFFFu : ((2u << first) - 1u);
                if (!pre_mask || (miss_mask & need)) ok = false;             // (no prefix within 32 groups: leave it to the slow path)
                s2 = warp_sum_u64(lane <= (u32)first ? val : 0ull);
                if (ok && lane == 0) atomicExch((ull *)&f.unit_scan.gpre[g - 1], (ull)(SCAN_FLAG_PRE | s2));
            }
            if (lane == 0) { s_excl1 = s1 + s2; s_ok1 = ok ? 1u : 0u; }
        } else if (tid == 0) s_ok1 = 0;
    };

    u32 it = 0;
    while (true) {
        const u32 u = s_ticket[it & 3];
        if (u >= num_units) break;
        u32 next_ticket = 0;
        if (tid == 0) next_ticket = atomicAdd(&f.ctrl->unit_ticket, 1u);    // three units ahead; stored at the end of this turn
        // prefetch: list entry of the unit after next, level 2 of the next unit, and the oldest pending unit's look-back words
        lvl1(s_ticket[(it + 2) & 3], nn_first, nn_end);
        lvl2();
        try_issue();
        FZ_TICK(0);
        if (c_cls == FZ_DENSE) {
            const u32 lenA = c_len;
            const u32 *Ac = f.a.colA + c_rs;
            const VT *Av = f.a.valA + c_rs;
            const u32 org = c_win.x, nwr = c_win.y;                          // {window origin, bitmap words, products} from the pre-pass
            const bool buffered = c_win.z <= pcap;                           // the row's products fit the product buffer
            const bool wraps = org != 0 && (u64)org + ((u64)nwr << 5) > (u64)ncols;
            auto dcol = [&](u32 c) -> u32 { u32 d = c - org; if (wraps && c < org) d += ncols; return d; };
            // every product of the row -> g(column, a value, index of the B entry, valid)
            auto products = [&](auto g) {
                if constexpr (PACK) {
                    auto expand = [&](const PackRec &r, VT av) {
                        const u32 len = r.a.y, st = r.a.x;
                        g(r.a.z, av, st, len > 0); g(r.a.w, av, st + 1, len > 1); g(r.b.x, av, st + 2, len > 2);
                        g(r.b.y, av, st + 3, len > 3); g(r.b.z, av, st + 4, len > 4); g(r.b.w, av, st + 5, len > 5);
                        if (len > B200_PACK_INLINE) for (u32 j = B200_PACK_INLINE; j < len; j++) g(f.a.colB[st + j], av, st + j, true);
                    };
#pragma unroll
                    for (int e = 0; e < FZ_KEEP; e++) expand(c_rec[e], c_av[e]);
                    for (u32 t = tid + FZ_KEEP * nt; t < lenA; t += nt) expand(load_pack(f.pack, Ac[t]), Av[t]);
                } else {
                    const u32 G = 1u << f.lg, sub = tid & (G - 1), grp = tid >> f.lg, ngrp = nt >> f.lg;
                    walk_products<VT>(Ac, lenA, f.a.bdesc, grp, ngrp, sub, G, [&](u32 t) { return Av[t]; },
                                      [&](VT av, u32 jb) { g(f.a.colB[jb], av, jb, true); });
                }
            };
            // ---------------------------------------------------------------- mark (+ fill the product buffer)
            if (buffered) {
                if constexpr (PACK) {
                    // a warp's 32 entries get a contiguous piece of the buffer: warp scan of the B-row lengths, one shared atomic
                    auto stage = [&](const PackRec &r, VT av) {
                        const u32 len = r.a.y, st = r.a.x;
                        u32 incl = len;
#pragma unroll
                        for (int dd = 1; dd < 32; dd <<= 1) { const u32 t = __shfl_up_sync(0xFFFFFFFFu, incl, dd); if (lane >= (u32)dd) incl += t; }
                        u32 pbase = 0;
                        if (lane == 31) pbase = atomicAdd(&s_P, incl);
                        pbase = __shfl_sync(0xFFFFFFFFu, pbase, 31);
                        const u32 dst = pbase + incl - len;
                        auto one = [&](u32 c, u32 j) {
                            const u32 d = dcol(c);
                            pd[dst + j] = d; px[dst + j] = (PV)fz_product<VT, MODE>(av, bpat ? (VT)1 : f.a.valB[st + j], bpat);
                            atomicOr(&bits[d >> 5], 1u << (d & 31));
                        };
                        if (len > 0) one(r.a.z, 0); if (len > 1) one(r.a.w, 1); if (len > 2) one(r.b.x, 2);
                        if (len > 3) one(r.b.y, 3); if (len > 4) one(r.b.z, 4); if (len > 5) one(r.b.w, 5);
                        for (u32 j = B200_PACK_INLINE; j < len; j++) one(f.a.colB[st + j], j);
                    };
                    const u32 lenA_w = (lenA + 31u) & ~31u;                  // whole warps take part in the scans
#pragma unroll
                    for (int e = 0; e < FZ_KEEP; e++) if (tid - lane + e * nt < lenA_w) stage(c_rec[e], c_av[e]);
                    for (u32 t = tid + FZ_KEEP * nt; t - lane < lenA_w; t += nt) {
                        PackRec r; r.a = make_uint4(0, 0, 0, 0); r.b = r.a; VT av = 0;
                        if (t < lenA) { av = Av[t]; r = load_pack(f.pack, Ac[t]); }
                        stage(r, av);
                    }
                } else {
                    // B rows of any length: G lanes per A entry; a product takes the next free slot (order is irrelevant)
                    products([&](u32 c, VT av, u32 jb, bool) {
                        const u32 d = dcol(c);
                        const u32 dst = atomicAdd(&s_P, 1u);
                        pd[dst] = d; px[dst] = (PV)fz_product<VT, MODE>(av, bpat ? (VT)1 : f.a.valB[jb], bpat);
                        atomicOr(&bits[d >> 5], 1u << (d & 31));
                    });
                }
            } else {
                products([&](u32 c, VT, u32, bool valid) { const u32 d = dcol(c); if (valid) atomicOr(&bits[d >> 5], 1u << (d & 31)); });
            }
            FZ_TICK(1);
            try_eval();
            __syncthreads();
            FZ_TICK(2);
            if (s_ok1) emit_oldest(s_excl1);                                 // (dense / tiny kinds only: no barrier inside)
            lvl3();
            FZ_TICK(3);
            // ---------------------------------------------------------------- rank: groups of four words in ascending-column order
            const u32 ngr = nwr >> 2;
            u32 rot = 0;
            if (wraps) { const u32 sw = (ncols - org) >> 7; rot = sw < ngr ? sw : 0u; }   // (ncols - org) is a multiple of 128 (pre-pass)
            const u32 gpt = (ngr + nt - 1) / nt;
            const u32 g0 = tid * gpt;
            u32 pg0 = g0 + rot; if (pg0 >= ngr) pg0 -= ngr;                  // my first group; the following ones wrap at ngr
            uint4 wr0 = make_uint4(0, 0, 0, 0), wr1 = wr0;                   // my words, read once (rows with up to two groups per thread)
            u32 cnt = 0;
            {
                u32 pg = pg0;
                if (gpt <= 2) {
                    if (g0 < ngr) { wr0 = bits4[pg]; cnt += __popc(wr0.x) + __popc(wr0.y) + __popc(wr0.z) + __popc(wr0.w); pg++; if (pg == ngr) pg = 0; }
                    if (gpt == 2 && g0 + 1 < ngr) { wr1 = bits4[pg]; cnt += __popc(wr1.x) + __popc(wr1.y) + __popc(wr1.z) + __popc(wr1.w); }
                } else {
                    for (u32 i = 0; i < gpt && g0 + i < ngr; i++) { const uint4 b4 = bits4[pg]; cnt += __popc(b4.x) + __popc(b4.y) + __popc(b4.z) + __popc(b4.w); pg++; if (pg == ngr) pg = 0; }
                }
            }
            u32 incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const u32 t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= (u32)d) incl += t; }
            if (lane == 31) s_wsum[w] = incl;
            FZ_TICK(4);
            __syncthreads();
            FZ_TICK(5);
            u32 wbase = 0, total = 0;
#pragma unroll
            for (u32 i = 0; i < 8; i++) { const u32 x = i < nwarps ? s_wsum[i] : 0u; if (i < w) wbase += x; total += x; }
            if (tid == 0) fz_publish(f.unit_scan, u, (u64)total);           // as early as possible: successors place themselves meanwhile
            {
                u32 run = wbase + incl - cnt, pg = pg0;
                auto put4 = [&](const uint4 &b4) {
                    const u32 p1 = run + __popc(b4.x), p2 = p1 + __popc(b4.y), p3 = p2 + __popc(b4.z);
                    pre4[pg] = make_uint4(run, p1, p2, p3); run = p3 + __popc(b4.w);
                    pg++; if (pg == ngr) pg = 0;
                };
                if (gpt <= 2) { if (g0 < ngr) put4(wr0); if (gpt == 2 && g0 + 1 < ngr) put4(wr1); }
                else for (u32 i = 0; i < gpt && g0 + i < ngr; i++) put4(bits4[pg]);
            }
            // ---------------------------------------------------------------- room in the ring (or the global slots for a row longer than the ring)
            lvl4();
            FZ_TICK(6);
            const bool big = total > R;
            if (big) { while (qn > 0) drain_one(true); }
            else { while (qn == FZ_QUEUE || ring_used + total > R) drain_one(true); }
            FZ_TICK(7);
            __syncthreads();
            FZ_TICK(8);
            // ---------------------------------------------------------------- accumulate at the column's rank
            if (!big && buffered) {
                const u32 rb = ring_next;
                const u32 P = s_P;
                for (u32 p = tid; p < P; p += nt) {                          // one product per lane
                    const u32 d = pd[p];
                    const u32 idx = (rb + pre[d >> 5] + __popc(bits[d >> 5] & ((1u << (d & 31)) - 1u))) & rmask;
                    u32 c = org + d; if (c >= ncols) c -= ncols;
                    acc.addv(idx, (u64)px[p]); cols[idx] = c;
                }
                FZ_TICK(9);
                __syncthreads();
                FZ_TICK(10);
                for (u32 t = tid; t < nwr; t += nt) bits[t] = 0;
                if (tid == 0) s_P = 0;
                push(u, c_first, 1u, total, FZ_DENSE, nullptr);
            } else if (!big) {
                const u32 rb = ring_next;
                products([&](u32 c, VT av, u32 jb, bool valid) {
                    const u32 d = valid ? dcol(c) : 0u;
                    const u32 wd = bits[d >> 5], wp = pre[d >> 5];
                    const u32 idx = (rb + wp + __popc(wd & ((1u << (d & 31)) - 1u))) & rmask;
                    if (valid) { acc.addv(idx, fz_product<VT, MODE>(av, bpat ? (VT)1 : f.a.valB[jb], bpat)); cols[idx] = c; }
                });
                __syncthreads();
                for (u32 t = tid; t < nwr; t += nt) bits[t] = 0;
                push(u, c_first, 1u, total, FZ_DENSE, nullptr);
            } else {
                products([&](u32 c, VT av, u32 jb, bool valid) {
                    if (!valid) return;
                    const u32 d = dcol(c);
                    const u32 rank = pre[d >> 5] + __popc(bits[d >> 5] & ((1u << (d & 31)) - 1u));
                    fz_spill_add<MODE>(&sp_acc[rank], fz_product<VT, MODE>(av, bpat ? (VT)1 : f.a.valB[jb], bpat)); sp_col[rank] = c;
                });
                if (w == 0) { const u64 excl = fz_wait(f.unit_scan, u, (int)lane); if (lane == 0) s_excl = excl; }
                __syncthreads();
                const u64 base = s_excl;
                for (u32 t = tid; t < total; t += nt) {
                    const u32 c = sp_col[t];
                    const VT v = emit_val<VT>(__ldcg(&sp_acc[t]));
                    sp_acc[t] = 0;
                    f.colC[base + t] = c; f.valC[base + t] = v;
                    vmax = vmax > (u64)v ? vmax : (u64)v;
                }
                for (u32 t = tid; t < nwr; t += nt) bits[t] = 0;
                if (tid == 0) {
                    s_P = 0;
                    f.rpC[c_first + 1] = base + total; maxnnz = max(maxnnz, total);
                    if (u == 0) f.rpC[0] = 0;
                    if (u + 1 == num_units) f.ctrl->total_nnz = base + total;
                }
            }
        } else if (c_cls == FZ_TINY) {
            // ---------------------------------------------------------------- one tiny row per warp, products in registers
            const bool tactive = c_first + w < c_end;
            const u32 dA = tactive ? c_len : 0u;
            u32 tkey; VT tval;
            fz_tiny_gather<VT>(dA, c_rec[0].a.x, lane < dA ? c_rec[0].a.y : 0u, c_av[0], f.a.colB, f.a.valB, (int)lane, tkey, tval, bpat);
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
            const u32 ttails = __ballot_sync(0xFFFFFFFFu, tail);
            if (lane == 0) s_wsum[w] = __popc(ttails);
            try_eval();
            __syncthreads();
            u32 wbase = 0, total = 0, rowcnt[8];                             // (kept in registers: a drain below may reuse s_wsum)
#pragma unroll
            for (u32 i = 0; i < 8; i++) { rowcnt[i] = i < nwarps ? s_wsum[i] : 0u; if (i < w) wbase += rowcnt[i]; total += rowcnt[i]; }
            if (tid == 0) fz_publish(f.unit_scan, u, (u64)total);
            if (s_ok1) { emit_oldest(s_excl1); __syncthreads(); }            // (the ring space it frees may be reused right below)
            lvl3();
            while (qn == FZ_QUEUE || ring_used + total > R) drain_one(true);  // (a tiny unit has at most 32 entries per warp <= R)
            if (tail) {
                const u32 idx = (ring_next + wbase + __popc(ttails & ((1u << lane) - 1u))) & rmask;
                cols[idx] = tkey; acc.set(idx, (u64)tval);
            }
            lvl4();
            push(u, c_first, c_end - c_first, total, FZ_TINY, rowcnt);
        } else {
            // ---------------------------------------------------------------- rows that are only placed (empty / counted beforehand)
            const u32 r = c_first + tid;
            u32 cnt = 0;
            if (r < c_end && f.rowclass[r] == FZ_OTHER) cnt = f.nnz_row[r];
            cnt = warp_sum_u32(cnt);
            if (lane == 0) s_wsum[w] = cnt;
            try_eval();
            __syncthreads();
            u32 total = 0;
#pragma unroll
            for (u32 i = 0; i < 8; i++) total += i < nwarps ? s_wsum[i] : 0u;
            if (tid == 0) fz_publish(f.unit_scan, u, (u64)total);
            if (s_ok1) emit_oldest(s_excl1);
            lvl3();
            while (qn == FZ_QUEUE) drain_one(true);
            lvl4();
            push(u, c_first, c_end - c_first, total, FZ_OTHER, nullptr);
        }
        rotate();
        if (tid == 0) s_ticket[(it + 3) & 3] = next_ticket;
        it++;
        __syncthreads();                                                     // queue slot, ring and bitmap writes visible; next ticket visible
        FZ_TICK(11);
    }
    while (qn > 0) drain_one(true);
    if (f.prof && tid == 0) { for (int i = 0; i < 12; i++) atomicAdd(&f.prof[i], (ull)pacc[i]); atomicAdd(&f.prof[12], (ull)it); }
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

// layout of the self-cleaning buffer: control block | tile scan {status, group sums, group prefixes} | unit scan {...}
static size_t fz_scan_words(u64 n) { return (size_t)(n + 2 * (n / FZ_GROUP + 2)); }
static FzScan fz_scan_at(u64 *base, u64 n) { FzScan sc; sc.st = base; sc.gagg = base + n; sc.gpre = sc.gagg + (n / FZ_GROUP + 2); return sc; }
static int fz_ensure_scratch(b200_ctx *ctx, u64 rows) {
    const u64 tiles = rows / FZ_PRE_MIN_ROWS + 2;
    if (!ctx->d_fz || tiles > ctx->cap_ftile || rows + 1 > ctx->cap_fustat) {
        dfree(ctx, ctx->d_fz); ctx->d_fz = nullptr;
        const u64 ct = tiles + tiles / 8 + 64, cu = rows + rows / 8 + 1024;
        TRY(dmalloc(ctx, (void **)&ctx->d_fz, 512 + (fz_scan_words(ct) + fz_scan_words(cu)) * 8));
        ctx->d_fctrl = (B200Ctrl *)ctx->d_fz;
        ctx->d_ftile = (u64 *)(ctx->d_fz + 512);
        ctx->d_fustat = ctx->d_ftile + fz_scan_words(ct);
        ctx->cap_ftile = ct; ctx->cap_fustat = cu;
        ctx->f_dirty = true;
    }
    if (rows + 1 > ctx->cap_frows) {
        dfree(ctx, ctx->d_units); dfree(ctx, ctx->d_rowwin); dfree(ctx, ctx->d_rowclass);
        ctx->d_units = nullptr; ctx->d_rowwin = nullptr; ctx->d_rowclass = nullptr; ctx->cap_frows = 0;
        const u64 cap = rows + rows / 8 + 1024;
        TRY(dmalloc(ctx, (void **)&ctx->d_units, (cap + 1) * 4));
        TRY(dmalloc(ctx, (void **)&ctx->d_rowwin, cap * sizeof(uint4)));
        TRY(dmalloc(ctx, (void **)&ctx->d_rowclass, cap));
        ctx->cap_frows = cap;
    }
    if (ctx->f_dirty) {
        CUDA_TRY(cudaMemsetAsync(ctx->d_fz, 0, 512 + (fz_scan_words(ctx->cap_ftile) + fz_scan_words(ctx->cap_fustat)) * 8, ctx->stream));
        ctx->f_dirty = false;
    }
    return B200_OK;
}

static int fz_ensure_spill(b200_ctx *ctx, u64 slots) {
    if (slots <= ctx->cap_spill) return B200_OK;
    dfree(ctx, ctx->d_spill_acc); dfree(ctx, ctx->d_spill_col); ctx->d_spill_acc = nullptr; ctx->d_spill_col = nullptr; ctx->cap_spill = 0;
    const u64 cap = slots + slots / 4;
    TRY(dmalloc(ctx, (void **)&ctx->d_spill_acc, cap * 8));
    TRY(dmalloc(ctx, (void **)&ctx->d_spill_col, cap * 4));
    CUDA_TRY(cudaMemsetAsync(ctx->d_spill_acc, 0, cap * 8, ctx->stream));
    ctx->cap_spill = cap;
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
                if (q != cudaSuccess) { ctx->f_dirty = true; ctx->lm_tot_dirty = true; return set_err(B200_ERR_CUDA, "multiply failed on the device: %s", cudaGetErrorString(q)); }
                if ((u32)(chunk[i] >> 32) != epoch) { ctx->f_dirty = true; ctx->lm_tot_dirty = true; return set_err(B200_ERR_CUDA, "the multiply finished without reporting (slot %d, epoch %u)", slot, epoch); }
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

// this multiply's slot in the pinned report ring; a slot still owned by an unread product is read first
int claim_report_slot(b200_ctx *ctx, u32 *epoch, int *slot, u64 **mirror) {
    const u32 e = ++ctx->fepoch ? ctx->fepoch : ++ctx->fepoch;              // never 0: a zeroed chunk is "no report"
    const int sl = (int)(e % B200_REPORT_SLOTS);
    if (ctx->slot_owner[sl]) TRY(resolve_pending(ctx, ctx->slot_owner[sl]));
    *epoch = e; *slot = sl; *mirror = ctx->h_freport + (size_t)sl * (sizeof(B200Ctrl) / 4);
    return B200_OK;
}
// hand a product out before its size is known: what the report will fill in, and what is known now
void mark_pending(b200_ctx *ctx, b200_csr *C, int slot, u32 epoch, const b200_csr *A, const b200_csr *B, int mode, int pipeline,
                  int32_t launches, bool timed, u64 max_row_len_bound) {
    C->pending_slot = slot; C->pending_epoch = epoch; ctx->slot_owner[slot] = C;
    C->nnz = 0; C->max_row_len = max_row_len_bound; C->h_maxval_known = false;
    b200_stats *st = new b200_stats();
    memset(st, 0, sizeof(*st));
    st->rows = A->rows; st->cols = B->cols; st->nnz_a = A->nnz; st->nnz_b = B->nnz; st->acc_mode = mode; st->pipeline = (uint32_t)pipeline;
    st->kernel_launches = launches;
    st->bytes_algorithmic = (A->nnz + B->nnz) * (u64)(4 + A->val_bits / 8) + (A->rows + B->rows + A->rows + 3) * 8;   // + nnz(C) entries at resolve
    delete C->stats;
    C->stats = st; C->stats_timed = timed;
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
    if (hc.error_flag) { ctx->f_dirty = true; ctx->lm_tot_dirty = true; return set_err(B200_ERR_CUDA, "a kernel of the multiply reported an impossible state (flag %u)", hc.error_flag); }
    m->nnz = hc.total_nnz; m->max_row_len = hc.max_row_nnz; m->h_maxval = hc.max_val_out; m->h_maxval_known = true;
    if (m->stats) {
        b200_stats *st = m->stats;
        st->nnz_c = m->nnz; st->products = hc.total_products; st->max_row_products = hc.max_row_products; st->max_row_nnz = hc.max_row_nnz;
        st->bytes_algorithmic += m->nnz * (u64)(4 + m->val_bits / 8);        // operands and row pointers were counted at launch
        if (st->pipeline == 1) {
            st->sym_bin_rows[0] = hc.class_count[FZ_TINY]; st->sym_bin_rows[1] = hc.class_count[FZ_DENSE];
            st->sym_bin_rows[2] = hc.class_count[FZ_OTHER]; st->sym_bin_rows[3] = hc.class_count[FZ_EMPTY];
            for (int i = 0; i < 6; i++) st->sym_bin_rows[10 + i] = hc.sym_bin_count[B200_BIN_WIDE0 + i];
            st->sym_bin_rows[9] = hc.sym_bin_count[B200_BIN_HEAVY];
        } else {
            for (int i = 0; i < B200_STAT_BINS; i++) st->sym_bin_rows[i] = hc.sym_bin_count[i];
        }
        if (ctx->timing && m->stats_timed) {
            if (cudaEventSynchronize(ctx->f_ev[slot][2]) == cudaSuccess) {
                // fused: pre-pass | numeric + placement.  binned: pre-pass + numeric kernels + row_ptr scan | compaction
                float first = 0, second = 0;
                cudaEventElapsedTime(&first, ctx->f_ev[slot][0], ctx->f_ev[slot][1]);
                cudaEventElapsedTime(&second, ctx->f_ev[slot][1], ctx->f_ev[slot][2]);
                if (st->pipeline != 2) { st->ms_symbolic = first; st->ms_numeric = second; } else { st->ms_numeric = first; st->ms_symbolic = second; }
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
    // pipeline 0 (auto) takes the binned pipeline: measured on the 30^3 chain (B200, round 2) the binned kernels are faster on
    // every power (55 / 72 / 102 / 143 / 208 / 316 us against 80 / 192 / 262 / 248 / 296 / 388 us here); see DESIGN.md
    if (cfg.pipeline != 1) return B200_OK;
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
    // ---- the window bitmap of a dense row ({bits, prefix} per 32 columns): per-row arcs of the index circle when B is
    //      square with known offset bounds (the pre-pass measures every row), else the whole column space
    const double meanP = ((double)A->nnz / (double)rows) * ((double)B->nnz / (double)B->rows);
    const int threads = cfg.fused_threads > 0 ? std::min(256, std::max(32, (int)cfg.fused_threads / 32 * 32)) : 128;
    const int nwarps = threads / 32;
    const u64 dense_pmax = cfg.fused_dense_pmax > 0 ? (u64)cfg.fused_dense_pmax : 16384;
    const u64 full = (ncols + 127) / 128 * 128;
    // the ring of accumulator + column slots in shared memory: room for a few finished rows (~3x the mean row's products;
    // nnz <= products), a power of two; a row longer than the whole ring goes through per-CTA global slots instead
    u64 ncap = 256; while (ncap < 8192 && (double)ncap < 3.0 * meanP) ncap <<= 1;
    if (cfg.fused_ring_slots > 0) { ncap = 256; while (ncap < 16384 && ncap < (u64)cfg.fused_ring_slots) ncap <<= 1; }
    // the product buffer (window offset + value per intermediate product): twice the mean row, a power of two
    u64 pcap = 256; while (pcap < 8192 && (double)pcap < 2.0 * meanP) pcap <<= 1;
    if (cfg.fused_product_slots > 0) { pcap = 64; while (pcap < 16384 && pcap < (u64)cfg.fused_product_slots) pcap <<= 1; }
    const size_t pvb = (mode == 0 || sizeof(VT) == 4) ? 4 : 8;             // PVal<MODE, VT>
    const size_t smem_budget = 96 * 1024;                                  // beyond this occupancy collapses: rows go to the counted lists
    u64 cap_cols = (u64)((smem_budget - ncap * (accb + 4) - pcap * (4 + pvb)) / 8) * 32 / 128 * 128;   // 8 bytes (bits + prefix) per 32 columns
    if (cfg.fused_window_cols > 0) cap_cols = std::min<u64>(cap_cols, ((u64)cfg.fused_window_cols + 127) / 128 * 128);
    int wmode = 0; u64 wcap = 0; u64 wneed = ~0ull;
    long long cs_lo = 0, cs_hi = 0;
    if (B->rows == B->cols && cfg.arc_window) {
        TRY(ensure_cs_bounds(ctx, B));
        if (B->cs_state == 1) {
            wmode = 1; cs_lo = B->cs_lo; cs_hi = B->cs_hi;
            if (A->max_row_span < ncols) wneed = (A->max_row_span + (u64)(cs_hi - cs_lo) + 1 + 127 + 127) / 128 * 128;
        }
    }
    if (wmode == 1) wcap = std::min(cap_cols, std::min(full, wneed));      // every row is measured; wcap is the longest window taken
    else wcap = full <= cap_cols ? full : 0;                               // whole column space or no dense class at all
    const bool all_fit = wcap != 0 && (wcap >= full || (wmode == 1 && wneed <= wcap));
    const bool guaranteed = all_fit && p_bound <= dense_pmax;              // no "other" rows can occur
    // ---- C's capacity from a host-known bound when that is cheap, else from the pre-pass's exact sum of min(P_i, cols)
    const size_t esz = 4 + sizeof(VT);
    unsigned __int128 hb128 = (unsigned __int128)A->nnz * B->max_row_len;
    const unsigned __int128 dense128 = (unsigned __int128)rows * ncols;
    if (dense128 < hb128) hb128 = dense128;
    const bool cheap_bound = hb128 * esz <= (unsigned __int128)(ctx->total_mem / 16);
    const bool need_pre_report = !guaranteed || !cheap_bound;

    TRY(ensure_row_scratch(ctx, rows));
    TRY(fz_ensure_scratch(ctx, rows));
    u32 epoch = 0; int slot = 0; u64 *mirror = nullptr;
    TRY(claim_report_slot(ctx, &epoch, &slot, &mirror));
    const bool timing = ctx->timing;
    if (timing) cudaEventRecord(ctx->f_ev[slot][0], s);

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
        pa.rowclass = ctx->d_rowclass; pa.rowwin = ctx->d_rowwin; pa.units = ctx->d_units;
        pa.tile_scan = fz_scan_at(ctx->d_ftile, ctx->cap_ftile); pa.unit_scan = fz_scan_at(ctx->d_fustat, ctx->cap_fustat);
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
        fa.pack = B->d_pack; fa.units = ctx->d_units; fa.rowclass = ctx->d_rowclass; fa.rowwin = ctx->d_rowwin; fa.nnz_row = ctx->d_nnz_row;
        fa.unit_scan = fz_scan_at(ctx->d_fustat, ctx->cap_fustat); fa.tile_scan = fz_scan_at(ctx->d_ftile, ctx->cap_ftile); fa.n_tile_status = tiles;
        fa.ctrl = ctx->d_fctrl; fa.rpC = C->d_rp; fa.colC = C->d_col; fa.valC = (VT *)C->d_val;
        fa.ncols = (u32)ncols; fa.nw = (u32)(wcap / 32); fa.ncap = (u32)ncap; fa.pcap = (u32)pcap;
        fa.wmode = wmode; fa.bpat = bpat ? 1 : 0; fa.lg = std::min(lg, 5); fa.narrow = ncols < (1ull << 27) ? 1 : 0;
        fa.finalize = others ? 0 : 1; fa.maxval_dst = C->d_maxval; fa.host_mirror = mirror; fa.epoch = need_pre_report ? epoch + 0x80000000u : epoch;
        const u32 final_epoch = fa.epoch;
        const size_t smem = (size_t)(wcap / 32) * 8 + (size_t)ncap * (accb + 4) + (size_t)pcap * (4 + pvb) + 16;
        const int m2 = sizeof(VT) == 8 ? mode : std::min(mode, 1);
        const FzKernel &fk = g_fzk[sizeof(VT) == 8][m2][packed ? 1 : 0];
        const int grid = (int)std::min<u64>(rows, (u64)ctx->num_sms * fz_ctas_per_sm(fk, threads, smem));
        // global slots for the ranks of rows longer than ncap (zeroed when allocated, left zeroed by the kernel)
        {
            const u64 longest = std::min<u64>(std::min<u64>(p_bound, dense_pmax), wcap);
            const u64 stride = longest > ncap ? longest : 0;
            r = fz_ensure_spill(ctx, (u64)grid * stride);
            if (r != B200_OK) { b200_csr_free(ctx, C); return r; }
            fa.spill_acc = ctx->d_spill_acc; fa.spill_col = ctx->d_spill_col; fa.spill_stride = stride;
        }
        fa.prof = nullptr;
        static ull *d_prof = nullptr; static int prof_on = -1;
        if (prof_on < 0) { const char *v = getenv("B200_FZ_PROF"); prof_on = v && *v == '1'; }
        if (prof_on) { if (!d_prof) cudaMalloc((void **)&d_prof, 16 * 8); cudaMemsetAsync(d_prof, 0, 16 * 8, s); fa.prof = d_prof; }
        if (ctx->trace) fprintf(stderr, "[b200 trace] fused: rows %llu wmode %d wcap %llu (wneed %llu, span %llu) ncap %llu pcap %llu spill %llu mode %d packed %d bpat %d threads %d grid %d (%d CTAs/SM, %d regs) smem %zu guaranteed %d cap_entries %llu\n",
                                (ull)rows, wmode, (ull)wcap, (ull)wneed, (ull)A->max_row_span, (ull)ncap, (ull)pcap, (ull)fa.spill_stride, mode, (int)packed, (int)bpat, threads, grid, fz_ctas_per_sm(fk, threads, smem), fk.regs, smem, (int)guaranteed, (ull)C->cap_entries);
        void *kargs[] = {(void *)&fa};
        const cudaError_t le = cudaLaunchKernel(fk.fn, dim3(grid), dim3(threads), kargs, smem, s);
        ctx->launches++;
        if (le != cudaSuccess) { b200_csr_free(ctx, C); return set_err(B200_ERR_CUDA, "fused numeric kernel launch failed: %s", cudaGetErrorString(le)); }
        if (prof_on) {
            ull hp[16]; cudaStreamSynchronize(s); cudaMemcpy(hp, d_prof, sizeof(hp), cudaMemcpyDeviceToHost);
            static const char *nm[12] = {"top(unit fetch)", "mark", "B1 wait", "fast emit", "rank", "B2 wait", "prefix+publish", "room (drains)", "B3 wait", "accumulate", "B4 wait", "tail"};
            double tot = 0; for (int i = 0; i < 12; i++) tot += (double)hp[i];
            fprintf(stderr, "[b200 prof] grid %d units/CTA %.1f cycles/unit %.0f:", grid, (double)hp[12] / grid, tot / (double)(hp[12] ? hp[12] : 1));
            for (int i = 0; i < 12; i++) fprintf(stderr, " %s %.1f%%", nm[i], 100.0 * (double)hp[i] / tot);
            fprintf(stderr, "\n");
        }
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
        mark_pending(ctx, C, slot, final_epoch, A, B, mode, 1, (int32_t)(ctx->launches - launches0), timing, std::min<u64>(p_bound, ncols));
        C->est_nnz = std::max<u64>(1, (u64)(meanP * (double)rows / 1.6));
        *out = C; *handled = true;
        if (st_out) { TRY(resolve_pending(ctx, C)); *st_out = *C->stats; }
    }
    return B200_OK;
}

template int spgemm_fused<u32>(b200_ctx *, const b200_csr *, const b200_csr *, b200_csr **, b200_stats *, bool *);
template int spgemm_fused<u64>(b200_ctx *, const b200_csr *, const b200_csr *, b200_csr **, b200_stats *, bool *);
