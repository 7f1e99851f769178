// rowwarp.cu -- the row-per-warp kernels of the exact placement: one warp owns one row of C from its first product to its
// last stored entry, nothing waits for another row, no CTA barrier anywhere.
//
//   k_rw<.., COUNT = true>    distinct output columns of a row: every intermediate product sets its column's bit in the warp's
//                             window bitmap (shared memory), the row's length is the popcount.  Feeds the row_ptr scan.
//   k_rw<.., COUNT = false>   values.  mark (as above) -> rank (exclusive prefix popcount per bitmap word, kept beside the
//                             word) -> accumulate (each product is added into acc[rank(column)]: compact accumulators of
//                             the row's LENGTH, not of its window, so a warp needs ~8 bytes per output entry and a bit per
//                             window column) -> emit (lane per output entry, coalesced stores at row_ptr_C[row]; the rank
//                             is the column's place in the sorted row, so there is no sort).
// Rows come from the pre-pass's bin lists (largest bin first, dealt round-robin over the resident warps); a row's column
// window {origin, groups} comes from the pre-pass too (whole column space, an operand-level arc, or a per-row plain /
// circular window).  Rows with more entries than the warp has accumulator slots are produced in several passes over
// their products (rank ranges), so the slot count is a tuning figure, not a limit.
//
// This is the MAGNUS "dense accumulation" category cut for a GPU: the accumulator is dense in RANK space,
// the bitmap is the only thing that scales with the window.  It replaces, for rows of 33..4096 intermediate products,
// the symbolic and numeric passes of CsrMatrix::matmul_par (/root/reference/src/graph_csr.rs:362-403, :430-476).
#include <type_traits>
#include <cooperative_groups.h>
#include "engine.cuh"
#include "devutil.cuh"

#define RW_WARPS 4
#define RW_THREADS (RW_WARPS * 32)
#ifndef RW_MIN_CTAS
#define RW_MIN_CTAS 4
#endif

template <typename VT>
struct RwArgs {
    NumArgs<VT> a;
    const uint4 *pack;
    const u32 *bin_rows; const u32 *bin_cnt; u32 bin_stride;
    int first_bin, nbins;        // lists first_bin .. first_bin + nbins - 1, walked from the last (longest rows) to the first
    const uint4 *win; u32 ncols;
    u32 nw;                      // bitmap words per warp (multiple of 4); no listed row's window has more
    u32 cap;                     // accumulator slots per warp
    u32 nsm;                     // SMs of the device (CTA -> slice mapping)
    u32 *nnz_row;                // COUNT: the row's length goes here
    const u64 *rpC; u32 *colC; VT *valC;
    B200Ctrl *ctrl;
};

__device__ __forceinline__ PackRec rw_load_pack(const uint4 *__restrict__ pack, u32 k) {
    PackRec r;
    asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(r.a.x), "=r"(r.a.y), "=r"(r.a.z), "=r"(r.a.w), "=r"(r.b.x), "=r"(r.b.y), "=r"(r.b.z), "=r"(r.b.w)
        : "l"(pack + 2 * (u64)k));
    return r;
}

template <int MODE, typename VT>
__device__ __forceinline__ u64 rw_product(VT a, VT b) {
    if (MODE == 0) return (u64)((u32)a * (u32)b);
    if (MODE == 2) return sat_mul((u64)a, (u64)b);
    u64 x = (u64)a * (u64)b;
    if (sizeof(VT) == 4) x = x > 0xFFFFFFFFull ? 0xFFFFFFFFull : x;
    return x;
}

// All intermediate products of one A row, by one warp.
// PACK (low-degree B, one 32-byte record per B row): lane per A entry; f6(columns[6], n, first index into B's arrays, a_ik)
// gets the record's inline columns at once (its slots are straight-line predicated code whose shared-memory reads overlap),
// f1(column, index, a_ik) the rest of a longer row.  Two entry streams (even / odd blocks of 32 entries) are kept in
// flight, each with its A column and value two blocks ahead and its record one block ahead, so a step waits for neither
// and no register is moved between the stages.
// Otherwise: lane per A entry for short B rows, the whole warp striding over the long ones (coalesced), all through f1.
// XT: what is kept of a_ik (u32 when 32-bit sums are proven, else the value type).
template <typename VT, typename XT, bool PACK, bool NEEDV, typename F6, typename F1>
__device__ __forceinline__ void rw_enumerate(const NumArgs<VT> &a, const uint4 *__restrict__ pack, u64 rs, u32 lenA, int lane, F6 f6, F1 f1) {
    const u32 *__restrict__ Ac = a.colA + rs;
    const VT *__restrict__ Av = a.valA + rs;
    if constexpr (PACK) {
        u32 kA = 0, kB = 0; XT xA = 0, xB = 0, xAn = 0, xBn = 0;
        PackRec rA, rB; rA.a.y = 0; rB.a.y = 0;
        {
            const u32 t0 = (u32)lane, t1 = t0 + 32, t2 = t0 + 64, t3 = t0 + 96;
            u32 k0 = 0, k1 = 0;
            if (t0 < lenA) { k0 = Ac[t0]; if (NEEDV) xA = (XT)Av[t0]; }
            if (t1 < lenA) { k1 = Ac[t1]; if (NEEDV) xB = (XT)Av[t1]; }
            if (t2 < lenA) { kA = Ac[t2]; if (NEEDV) xAn = (XT)Av[t2]; }
            if (t3 < lenA) { kB = Ac[t3]; if (NEEDV) xBn = (XT)Av[t3]; }
            if (t0 < lenA) rA = rw_load_pack(pack, k0);
            if (t1 < lenA) rB = rw_load_pack(pack, k1);
        }
        auto step = [&](PackRec &rec, XT &x, XT &xn, u32 &kn, u32 t) {
            // t: this lane's entry of the block being consumed; its stream's next entry is t + 64, the one after t + 128
            const u32 len = rec.a.y, st = rec.a.x;
            const u32 c[B200_PACK_INLINE] = {rec.a.z, rec.a.w, rec.b.x, rec.b.y, rec.b.z, rec.b.w};
            const XT xc = x;
            rec.a.y = 0;
            if (t + 64 < lenA) rec = rw_load_pack(pack, kn);
            x = xn;
            if (t + 128 < lenA) { kn = Ac[t + 128]; if (NEEDV) xn = (XT)Av[t + 128]; }
            if (len) {
                // a record's unused inline slots repeat its last column (k_build_pack), so all six slots run unpredicated:
                // marking a column twice is harmless, and f6 adds zero for the slots >= n
                f6(c, len < B200_PACK_INLINE ? len : (u32)B200_PACK_INLINE, st, xc);
                for (u32 j = B200_PACK_INLINE; j < len; j++) f1(a.colB[st + j], st + j, xc);
            }
        };
        for (u32 base = 0; base < lenA; base += 64) {
            step(rA, xA, xAn, kA, base + lane);
            if (base + 32 < lenA) step(rB, xB, xBn, kB, base + 32 + lane);
        }
    } else {
        for (u32 base = 0; base < lenA; base += 32) {
            const u32 t = base + lane;
            const bool valid = t < lenA;
            u32 st = 0, len = 0; XT x = 0;
            if (valid) { const uint2 d = a.bdesc[Ac[t]]; st = d.x; len = d.y; if (NEEDV) x = (XT)Av[t]; }
            u32 longm = __ballot_sync(0xFFFFFFFFu, len >= 16);
            while (longm) {
                const int src = __ffs(longm) - 1;
                longm &= longm - 1;
                const u32 s = __shfl_sync(0xFFFFFFFFu, st, src), l = __shfl_sync(0xFFFFFFFFu, len, src);
                const XT xs = shfl_any(x, src);
                for (u32 j = lane; j < l; j += 32) f1(a.colB[s + j], s + j, xs);
            }
            if (len < 16) for (u32 j = 0; j < len; j++) f1(a.colB[st + j], st + j, x);
        }
    }
}

// The rows a warp takes: every list (longest rows first) is cut into one slice per CTA, CTAs resident on the same SM
// take adjacent slices, and the warps of a CTA interleave inside their slice -- so the rows in flight on an SM are
// neighbours and share most of their B records (lattice-like operands), which then come from L1 instead of L2.
struct RwCursor { int b; u32 r, hi; };
template <typename VT>
__device__ __forceinline__ bool rw_next(const RwArgs<VT> &p, RwCursor &cur, u32 slice, u32 nslices, u32 wid, u32 &row) {
    while (true) {
        if (cur.r < cur.hi) { row = p.bin_rows[(u64)cur.b * p.bin_stride + cur.r]; cur.r += RW_WARPS; return true; }
        if (--cur.b < p.first_bin) return false;
        const u64 cnt = p.bin_cnt[cur.b];
        cur.r = (u32)(cnt * slice / nslices) + wid; cur.hi = (u32)(cnt * (slice + 1) / nslices);
    }
}

// shared-memory accesses by 32-bit shared address (one register, immediate offsets; no generic-address arithmetic)
__device__ __forceinline__ void sm_red_or(u32 addr, u32 v) { asm volatile("red.shared.or.b32 [%0], %1;" :: "r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void sm_red_add(u32 addr, u32 v) { asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void sm_st_u16(u32 addr, u32 v) { asm volatile("st.shared.u16 [%0], %1;" :: "r"(addr), "h"((unsigned short)v) : "memory"); }
__device__ __forceinline__ uint2 sm_ld_v2(u32 addr) {
    uint2 r;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(addr) : "memory");
    return r;
}

// WRAP: the window's origin is not column 0 (bit d is column (org + d) mod ncols); else d = c.
template <bool WRAP>
__device__ __forceinline__ u32 rw_dcol(u32 c, u32 org, u32 ncols) {
    if (!WRAP) return c;
    const u32 d = c - org;
    return c >= org ? d : d + ncols;
}

// One warp's share of shared memory and the per-row phases that work on it.
//   count layout  : u32 bitmap[nw]
//   numeric layout: {bits, prefix} uint2[nw] | acc[cap] | window offsets u16[cap]          (nw % 32 == 0)
// Both start at the same address and both leave every word they touched zeroed, so one kernel may run count_row over its
// rows first and numeric_row afterwards (k_rw_fused) on the same zero-initialised region.
template <typename VT, int MODE, bool PACK, bool BPAT>
struct RwWarp {
    typedef typename std::conditional<MODE == 0, u32, VT>::type XT;          // a_ik as the accumulate phase keeps it
    const NumArgs<VT> &a; const uint4 *pack;
    u32 *bm; uint2 *bw; Acc<MODE> acc; unsigned short *offs;
    u32 sm_bits, sm_acc, sm_offs, nw, cap, ncols; int lane;

    __device__ __forceinline__ RwWarp(const NumArgs<VT> &a_, const uint4 *pack_, unsigned char *base, u32 nw_, u32 cap_, u32 ncols_, int lane_)
        : a(a_), pack(pack_), nw(nw_), cap(cap_), ncols(ncols_), lane(lane_) {
        bm = reinterpret_cast<u32 *>(base); bw = reinterpret_cast<uint2 *>(base);
        acc.bind(base + (size_t)nw * 8, cap);
        offs = reinterpret_cast<unsigned short *>(base + (size_t)nw * 8 + Acc<MODE>::bytes(cap));
        sm_bits = (u32)__cvta_generic_to_shared(base); sm_acc = sm_bits + nw * 8; sm_offs = sm_acc + (u32)Acc<MODE>::bytes(cap);
    }
    static __host__ __device__ size_t bytes(bool count_only, u32 nw, u32 cap) {
        return count_only ? (size_t)nw * 4 : (size_t)nw * 8 + Acc<MODE>::bytes(cap) + (size_t)cap * 2;
    }
    __device__ __forceinline__ void zero(bool count_only) {
        if (count_only) { for (u32 t = lane; t < nw; t += 32) bm[t] = 0; }
        else { for (u32 t = lane; t < nw; t += 32) bw[t] = make_uint2(0u, 0u); for (u32 t = lane; t < cap; t += 32) acc.clear(t); }
        __syncwarp();
    }

    // every product sets its column's bit (bit d of the window is column (org + d) mod ncols); STRIDE: bytes between bitmap words.
    // wlo / whi: lowest and highest bitmap word the warp touched (the later phases walk only those).
    template <u32 STRIDE, bool SUMP>
    __device__ __forceinline__ u32 mark(u64 rs, u32 lenA, u32 org, u32 &wlo, u32 &whi) {
        u32 psum = 0, lo = 0xFFFFFFFFu, hi = 0;
        auto go = [&](auto wrap) {
            constexpr bool WRAP = decltype(wrap)::value;
            rw_enumerate<VT, u32, PACK, false>(a, pack, rs, lenA, lane,
                [&](const u32 (&c)[B200_PACK_INLINE], u32 n, u32, u32) {
                    if (SUMP) psum += n;
#pragma unroll
                    for (int j = 0; j < B200_PACK_INLINE; j++) {
                        const u32 d = rw_dcol<WRAP>(c[j], org, ncols);
                        sm_red_or(sm_bits + (d >> 5) * STRIDE, __funnelshift_l(0u, 1u, d));
                        // a record's columns ascend: without a wrap the first and the last (repeated) slot bound them
                        if (WRAP || j == 0) lo = min(lo, d);
                        if (WRAP || j == B200_PACK_INLINE - 1) hi = max(hi, d);
                    }
                },
                [&](u32 c, u32, u32) {
                    if (SUMP) psum++;
                    const u32 d = rw_dcol<WRAP>(c, org, ncols);
                    sm_red_or(sm_bits + (d >> 5) * STRIDE, __funnelshift_l(0u, 1u, d));
                    lo = min(lo, d); hi = max(hi, d);
                });
        };
        if (org) go(std::true_type{}); else go(std::false_type{});
        wlo = __reduce_min_sync(0xFFFFFFFFu, lo) >> 5; whi = __reduce_max_sync(0xFFFFFFFFu, hi) >> 5;
        return psum;
    }

    // distinct columns of the row; `mid` runs between the mark and the popcount (the callers fetch their next row's header there).
    // P (SUMP): intermediate products of the row, summed over the warp.
    template <bool SUMP, typename MID>
    __device__ __forceinline__ u32 count_row(u64 rs, u32 lenA, u32 org, u32 words, u32 &P, MID mid) {
        u32 wlo, whi;
        const u32 psum = mark<4u, SUMP>(rs, lenA, org, wlo, whi);
        mid();
        __syncwarp();
        u32 mine = 0;
        if (wlo <= whi) {                                                    // (a row without products touches nothing)
            const u32 wpl = ((whi - wlo + 32u) >> 5) | 1u, w0 = wlo + (u32)lane * wpl;   // (odd: the lanes fall on different banks)
            for (u32 i = 0; i < wpl; i++) if (w0 + i <= whi) { mine += __popc(bm[w0 + i]); bm[w0 + i] = 0; }
        }
        if (SUMP) P = warp_sum_u32(psum);
        const u32 nnz = warp_sum_u32(mine);
        __syncwarp();
        return nnz;
    }

    // values of the row into colp / valp (the row's place in C); returns its length
    template <typename MID>
    __device__ __forceinline__ u32 numeric_row(u64 rs, u32 lenA, u32 org, u32 words, u32 *colp, VT *valp, const VT *valB, u64 &vmax, MID mid) {
        u32 wlo, whi;
        mark<8u, false>(rs, lenA, org, wlo, whi);
        mid();
        __syncwarp();
        if (wlo > whi) return 0u;                                            // no products (uniform over the warp)
        // ---- rank: consecutive words per lane over the touched span, warp scan of the lanes' popcounts
        const u32 wpl = ((whi - wlo + 32u) >> 5) | 1u, w0 = wlo + (u32)lane * wpl;   // (odd: the lanes fall on different banks)
        u32 mine = 0;
        for (u32 i = 0; i < wpl; i++) if (w0 + i <= whi) mine += __popc(bw[w0 + i].x);
        u32 incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const u32 t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
        const u32 nnz = __shfl_sync(0xFFFFFFFFu, incl, 31);
        u32 run = incl - mine;
        for (u32 i = 0; i < wpl; i++) if (w0 + i <= whi) { const u32 b = bw[w0 + i].x; bw[w0 + i].y = run; run += __popc(b); }
        __syncwarp();
        // Ranks are in d order; when the window starts at a column org > 0 the entries whose column lies below org
        // (d >= ncols - org) belong in FRONT of the others: the row is written rotated by r0 = entries with d < ncols - org.
        u32 r0 = nnz;
        if (org) {
            const u32 split = ncols - org, sw = split >> 5;
            if (sw < wlo) r0 = 0;
            else if (sw <= whi) { const uint2 s = bw[sw]; r0 = s.y + __popc(s.x & (__funnelshift_l(0u, 1u, split) - 1u)); }
        }
        const u32 shift_hi = nnz - r0;
        // ---- accumulate at the column's rank, then emit; rows longer than the accumulator take several passes over their
        //      products (ranks pass .. pass + cap - 1 in each)
        auto accumulate = [&](auto wrap, auto multi, u32 pass) {
            constexpr bool WRAP = decltype(wrap)::value, MULTI = decltype(multi)::value;
            auto acc1 = [&](u32 c, u32 jb, XT av) {
                const u32 d = rw_dcol<WRAP>(c, org, ncols);
                const uint2 s = bw[d >> 5];
                const u32 pos = s.y + __popc(s.x & (__funnelshift_l(0u, 1u, d) - 1u)) - pass;
                if (!MULTI || pos < cap) {
                    offs[pos] = (unsigned short)d;
                    acc.addv(pos, BPAT ? (u64)av : rw_product<MODE, VT>((VT)av, valB[jb]));
                }
            };
            rw_enumerate<VT, XT, PACK, true>(a, pack, rs, lenA, lane,
                [&](const u32 (&c)[B200_PACK_INLINE], u32 n, u32 st, XT av) {
                    // all six bitmap words first, then the six accumulations: the reads do not wait for the writes.  Slots >= n
                    // repeat the record's last column: they add zero to its accumulator.
                    u32 d[B200_PACK_INLINE], pos[B200_PACK_INLINE]; uint2 s[B200_PACK_INLINE];
#pragma unroll
                    for (int j = 0; j < B200_PACK_INLINE; j++) { d[j] = rw_dcol<WRAP>(c[j], org, ncols); s[j] = sm_ld_v2(sm_bits + (d[j] >> 5) * 8u); }
#pragma unroll
                    for (int j = 0; j < B200_PACK_INLINE; j++) pos[j] = s[j].y + __popc(s[j].x & (__funnelshift_l(0u, 1u, d[j]) - 1u)) - pass;
#pragma unroll
                    for (int j = 0; j < B200_PACK_INLINE; j++) {
                        const bool live = (u32)j < n;
                        if (MODE == 0 && BPAT && !MULTI) {
                            sm_st_u16(sm_offs + pos[j] * 2u, d[j]);
                            sm_red_add(sm_acc + pos[j] * 4u, live ? (u32)av : 0u);
                        } else if (live && (!MULTI || pos[j] < cap)) {
                            offs[pos[j]] = (unsigned short)d[j];
                            acc.addv(pos[j], BPAT ? (u64)av : rw_product<MODE, VT>((VT)av, valB[st + j]));
                        }
                    }
                }, acc1);
        };
        auto emit = [&](u32 pass) {
            const u32 m = min(cap, nnz - pass);
            for (u32 t = lane; t < m; t += 32) {
                u32 c = org + (u32)offs[t]; if (c >= ncols) c -= ncols;
                const VT v = emit_val<VT>(acc.get(t));
                acc.clear(t);
                const u32 g = pass + t;
                const u32 q = g >= r0 ? g - r0 : g + shift_hi;
                colp[q] = c; valp[q] = v;
                vmax = vmax > (u64)v ? vmax : (u64)v;
            }
        };
        if (nnz <= cap) {
            if (org) accumulate(std::true_type{}, std::false_type{}, 0u); else accumulate(std::false_type{}, std::false_type{}, 0u);
            __syncwarp();
            emit(0u);
            __syncwarp();
        } else {
            for (u32 pass = 0; pass < nnz; pass += cap) {
                accumulate(std::true_type{}, std::true_type{}, pass);
                __syncwarp();
                emit(pass);
                __syncwarp();
            }
        }
        for (u32 i = 0; i < wpl; i++) if (w0 + i <= whi) bw[w0 + i] = make_uint2(0u, 0u);
        __syncwarp();
        return nnz;
    }
};

__device__ __forceinline__ u32 rw_origin(const uint4 &wn, u32 ncols) { const u64 t = (u64)wn.x + wn.z; return (u32)(t >= ncols ? t - ncols : t); }

template <typename VT, int MODE, bool PACK, bool BPAT, bool COUNT>
__global__ void __launch_bounds__(RW_THREADS, COUNT ? 10 : RW_MIN_CTAS) k_rw(RwArgs<VT> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const u32 nslices = gridDim.x;
    const u32 per_sm = p.nsm && gridDim.x % p.nsm == 0 ? gridDim.x / p.nsm : 0u;
    const u32 slice = per_sm ? (blockIdx.x % p.nsm) * per_sm + blockIdx.x / p.nsm : blockIdx.x;
    RwCursor cur; cur.b = p.first_bin + p.nbins; cur.r = 0; cur.hi = 0;
    u32 row;
    bool have = rw_next(p, cur, slice, nslices, (u32)wid, row);
    if (!have) return;
    typedef RwWarp<VT, MODE, PACK, BPAT> W;
    W w(p.a, p.pack, smem_raw + (size_t)wid * W::bytes(COUNT, p.nw, p.cap), p.nw, p.cap, p.ncols, lane);
    w.zero(COUNT);
    u64 vmax = 0;
    // row header one row ahead: id, A row extent, window, output base
    u64 rs = p.a.rpA[row];
    u32 lenA = (u32)(p.a.rpA[row + 1] - rs);
    uint4 wn = p.win[row];
    u64 obase = COUNT ? 0ull : p.rpC[row];
    while (have) {
        u32 row_n = 0;
        const bool has_next = rw_next(p, cur, slice, nslices, (u32)wid, row_n);
        u64 rs_n = 0; u32 lenA_n = 0; uint4 wn_n = make_uint4(0, 0, 0, 0); u64 obase_n = 0;
        auto mid = [&]() {                                                   // next row's header (its id has arrived by now)
            if (has_next) {
                rs_n = p.a.rpA[row_n]; lenA_n = (u32)(p.a.rpA[row_n + 1] - rs_n); wn_n = p.win[row_n];
                if (!COUNT) obase_n = p.rpC[row_n];
            }
        };
        const u32 org = rw_origin(wn, p.ncols);
        if constexpr (COUNT) {
            u32 P;
            const u32 nnz = w.template count_row<false>(rs, lenA, org, wn.y * 4u, P, mid);
            if (lane == 0) p.nnz_row[row] = nnz;
        } else {
            w.numeric_row(rs, lenA, org, wn.y * 4u, p.colC + obase, p.valC + obase, p.a.valB, vmax, mid);
        }
        have = has_next; row = row_n; rs = rs_n; lenA = lenA_n; wn = wn_n; obase = obase_n;
    }
    if (!COUNT) {
        vmax = warp_max_u64(vmax);
        if (lane == 0 && vmax) atomicMax(&p.ctrl->max_val_out, (ull)vmax);
    }
}

// =======================================================================================
// The whole multiply as ONE cooperative kernel (pipeline 4): every row's window is the same (the whole column space or the
// operand-level arc), so nothing has to be classified beforehand.
//   phase A  every CTA owns a contiguous range of rows (CTAs of one SM own adjacent ranges), its warps interleave over
//            them: count_row -> nnz_row[row], intermediate products; the CTA's total goes to cta_tot[slice]
//   grid sync
//   phase B  the CTA's base = sum of the totals of the slices before it (no look-back chain: everything is there);
//            row_ptr of its rows by a block scan over nnz_row
//   phase C  numeric_row for every row, at row_ptr
//   the last CTA to finish reports the control block to the pinned ring, hands the largest value to the product's handle
//   and leaves the control block zeroed.
// Replaces pre-pass, count kernels, row_ptr scan, numeric kernels and the finishing kernel of the exact placement: one
// launch per multiply, which is what the small powers of a chain are bound by.
// =======================================================================================
template <typename VT>
struct RwFusedArgs {
    NumArgs<VT> a; const uint4 *pack;
    u64 rows; u32 ncols, org, words;  // one window for every row: bit d is column (org + d) mod ncols, `words` bitmap words
    u32 nw, cap, nsm;
    u32 *nnz_row; u64 *cta_tot;
    u64 *rpC; u32 *colC; VT *valC;
    B200Ctrl *ctrl; u64 *host_mirror; u32 epoch; ull *maxval_dst;
};

template <typename VT, int MODE, bool PACK, bool BPAT>
__global__ void __launch_bounds__(RW_THREADS, RW_MIN_CTAS) k_rw_fused(RwFusedArgs<VT> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ u64 s_base, s_tot; __shared__ ull s_P, s_maxP; __shared__ u32 s_maxN, s_last, s_scan[RW_WARPS + 1];
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const u32 nslices = gridDim.x;
    const u32 per_sm = p.nsm && gridDim.x % p.nsm == 0 ? gridDim.x / p.nsm : 0u;
    const u32 slice = per_sm ? (blockIdx.x % p.nsm) * per_sm + blockIdx.x / p.nsm : blockIdx.x;
    const u64 r_lo = p.rows * slice / nslices, r_hi = p.rows * (slice + 1) / nslices;
    typedef RwWarp<VT, MODE, PACK, BPAT> W;
    W w(p.a, p.pack, smem_raw + (size_t)wid * W::bytes(false, p.nw, p.cap), p.nw, p.cap, p.ncols, lane);
    w.zero(false);
    if (tid == 0) { s_tot = 0; s_P = 0; s_maxP = 0; s_maxN = 0; }
    __syncthreads();
    // ---- phase A: lengths
    {
        u64 tot = 0, Psum = 0; u32 maxP = 0, maxN = 0;
        u64 row = r_lo + wid;
        u64 rs = 0; u32 lenA = 0;
        if (row < r_hi) { rs = p.a.rpA[row]; lenA = (u32)(p.a.rpA[row + 1] - rs); }
        for (; row < r_hi; row += RW_WARPS) {
            const u64 row_n = row + RW_WARPS;
            u64 rs_n = 0; u32 lenA_n = 0;
            auto mid = [&]() { if (row_n < r_hi) { rs_n = p.a.rpA[row_n]; lenA_n = (u32)(p.a.rpA[row_n + 1] - rs_n); } };
            u32 P = 0, nnz = 0;
            if (lenA) nnz = w.template count_row<true>(rs, lenA, p.org, p.words, P, mid); else mid();
            if (lane == 0) p.nnz_row[row] = nnz;
            tot += nnz; Psum += P; maxP = max(maxP, P); maxN = max(maxN, nnz);
            rs = rs_n; lenA = lenA_n;
        }
        if (lane == 0) { atomicAdd((ull *)&s_tot, (ull)tot); atomicAdd(&s_P, (ull)Psum); atomicMax(&s_maxP, (ull)maxP); atomicMax(&s_maxN, maxN); }
    }
    __syncthreads();
    if (tid == 0) {
        p.cta_tot[slice] = s_tot;
        if (s_P) atomicAdd(&p.ctrl->total_products, s_P);
        atomicMax(&p.ctrl->max_row_products, s_maxP);
        atomicMax(&p.ctrl->max_row_nnz, (ull)s_maxN);
        __threadfence();
    }
    grid.sync();
    // ---- phase B: this CTA's first entry, row_ptr of its rows
    if (wid == 0) {
        u64 sum = 0;
        for (u32 i = lane; i < slice; i += 32) sum += __ldcg(&p.cta_tot[i]);
        sum = warp_sum_u64(sum);
        if (lane == 0) s_base = sum;
    }
    __syncthreads();
    {
        u64 run = s_base;
        for (u64 r0 = r_lo; r0 < r_hi; r0 += RW_THREADS) {
            const u64 r = r0 + tid;
            const u32 v = r < r_hi ? p.nnz_row[r] : 0u;
            u32 incl = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const u32 t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
            if (lane == 31) s_scan[wid] = incl;
            __syncthreads();
            u32 wbase = 0, total = 0;
#pragma unroll
            for (int i = 0; i < RW_WARPS; i++) { if (i < wid) wbase += s_scan[i]; total += s_scan[i]; }
            if (r < r_hi) p.rpC[r] = run + wbase + incl - v;
            run += total;
            __syncthreads();
        }
        if (slice == nslices - 1 && tid == 0) { p.rpC[p.rows] = run; p.ctrl->total_nnz = run; }
    }
    __syncthreads();
    // ---- phase C: values
    {
        u64 vmax = 0;
        u64 row = r_lo + wid;
        u64 rs = 0, obase = 0; u32 lenA = 0;
        if (row < r_hi) { rs = p.a.rpA[row]; lenA = (u32)(p.a.rpA[row + 1] - rs); obase = p.rpC[row]; }
        for (; row < r_hi; row += RW_WARPS) {
            const u64 row_n = row + RW_WARPS;
            u64 rs_n = 0, obase_n = 0; u32 lenA_n = 0;
            auto mid = [&]() { if (row_n < r_hi) { rs_n = p.a.rpA[row_n]; lenA_n = (u32)(p.a.rpA[row_n + 1] - rs_n); obase_n = p.rpC[row_n]; } };
            if (lenA) w.numeric_row(rs, lenA, p.org, p.words, p.colC + obase, p.valC + obase, p.a.valB, vmax, mid); else mid();
            rs = rs_n; lenA = lenA_n; obase = obase_n;
        }
        vmax = warp_max_u64(vmax);
        if (lane == 0 && vmax) atomicMax(&p.ctrl->max_val_out, (ull)vmax);
    }
    // ---- report: the last CTA to get here
    __syncthreads();
    if (tid == 0) { __threadfence(); s_last = atomicAdd(&p.ctrl->fused_done, 1u) == gridDim.x - 1 ? 1u : 0u; }
    __syncthreads();
    if (s_last) {
        __threadfence();
        const volatile u32 *src = reinterpret_cast<const volatile u32 *>(p.ctrl);
        for (u32 i = tid; i < sizeof(B200Ctrl) / 4; i += blockDim.x) st_volatile_u64(p.host_mirror + i, ((u64)p.epoch << 32) | (u64)src[i]);
        if (tid == 0) *p.maxval_dst = *reinterpret_cast<volatile ull *>(&p.ctrl->max_val_out);
        __syncthreads();
        u32 *cw = reinterpret_cast<u32 *>(p.ctrl);
        for (u32 i = tid; i < sizeof(B200Ctrl) / 4; i += blockDim.x) cw[i] = 0;
    }
}

// ---------------------------------------------------------------------------- host side
struct RwKernel { const void *fn; int regs; size_t static_smem; };
// [value width][mode][packed][pattern-only B][count]
static RwKernel g_rw[2][3][2][2][2];
template <typename VT, int MODE, bool PACK, bool BPAT, bool COUNT>
static void rw_register(size_t optin) {
    RwKernel &k = g_rw[sizeof(VT) == 8][MODE][PACK][BPAT][COUNT];
    k.fn = (const void *)k_rw<VT, MODE, PACK, BPAT, COUNT>;
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, k.fn) == cudaSuccess) { k.regs = fa.numRegs; k.static_smem = fa.sharedSizeBytes; } else { cudaGetLastError(); k.regs = 64; k.static_smem = 0; }
    if (cudaFuncSetAttribute(k.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(optin - k.static_smem)) != cudaSuccess) cudaGetLastError();
}
template <typename VT, int MODE>
static void rw_register_mode(size_t o) {
    rw_register<VT, MODE, false, false, false>(o); rw_register<VT, MODE, false, true, false>(o);
    rw_register<VT, MODE, true, false, false>(o); rw_register<VT, MODE, true, true, false>(o);
}
void rw_setup(b200_ctx *ctx) {
    const size_t o = ctx->smem_optin;
    rw_register_mode<u32, 0>(o); rw_register_mode<u32, 1>(o);
    rw_register_mode<u64, 0>(o); rw_register_mode<u64, 1>(o); rw_register_mode<u64, 2>(o);
    // the count kernels touch no values: one pair (packed or not) serves every width and mode
    rw_register<u32, 0, false, true, true>(o); rw_register<u32, 0, true, true, true>(o);
}

size_t rw_smem_per_warp(bool count, int mode, u32 nw, u32 cap);
// fused (cooperative) kernel variants: [value width][mode][packed][pattern-only B]
static RwKernel g_rwf[2][3][2][2];
template <typename VT, int MODE, bool PACK, bool BPAT>
static void rwf_register(size_t optin) {
    RwKernel &k = g_rwf[sizeof(VT) == 8][MODE][PACK][BPAT];
    k.fn = (const void *)k_rw_fused<VT, MODE, PACK, BPAT>;
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, k.fn) == cudaSuccess) { k.regs = fa.numRegs; k.static_smem = fa.sharedSizeBytes; } else { cudaGetLastError(); k.regs = 64; k.static_smem = 0; }
    if (cudaFuncSetAttribute(k.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(optin - k.static_smem)) != cudaSuccess) cudaGetLastError();
}
template <typename VT, int MODE>
static void rwf_register_mode(size_t o) {
    rwf_register<VT, MODE, false, false>(o); rwf_register<VT, MODE, false, true>(o); rwf_register<VT, MODE, true, false>(o); rwf_register<VT, MODE, true, true>(o);
}
void rwf_setup(b200_ctx *ctx) {
    const size_t o = ctx->smem_optin;
    rwf_register_mode<u32, 0>(o); rwf_register_mode<u32, 1>(o);
    rwf_register_mode<u64, 0>(o); rwf_register_mode<u64, 1>(o); rwf_register_mode<u64, 2>(o);
}

template <typename VT>
static cudaError_t rwf_go(b200_ctx *ctx, const b200_csr *A, const b200_csr *B, b200_csr *C, B200Ctrl *ctrl, u32 org, u32 words, u32 nw, u32 cap,
                          u64 *mirror, u32 epoch, const void *fn, int grid, size_t smem, cudaStream_t s) {
    RwFusedArgs<VT> p;
    p.a = NumArgs<VT>{A->d_rp, A->d_col, (const VT *)A->d_val, B->d_desc, B->d_col, (const VT *)B->d_val};
    p.pack = B->d_pack; p.rows = A->rows; p.ncols = (u32)B->cols; p.org = org; p.words = words; p.nw = nw; p.cap = cap; p.nsm = (u32)ctx->num_sms;
    p.nnz_row = ctx->d_nnz_row; p.cta_tot = ctx->d_cta_tot; p.rpC = C->d_rp; p.colC = C->d_col; p.valC = (VT *)C->d_val;
    p.ctrl = ctrl; p.host_mirror = mirror; p.epoch = epoch; p.maxval_dst = C->d_maxval;
    void *kargs[] = {(void *)&p};
    return cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(RW_THREADS), kargs, smem, s);
}

// The whole multiply in one cooperative launch (every row shares the window {org, words}); C's arrays are already allocated.
int rw_fused_launch(b200_ctx *ctx, const b200_csr *A, const b200_csr *B, b200_csr *C, B200Ctrl *ctrl, int mode, bool packed, bool bpat,
                    u32 org, u32 words, u32 nw, u32 cap, u64 *mirror, u32 epoch, cudaStream_t s) {
    const bool v64 = A->val_bits == 64;
    const RwKernel &k = g_rwf[v64 ? 1 : 0][v64 ? mode : std::min(mode, 1)][packed ? 1 : 0][bpat ? 1 : 0];
    if (!k.fn) return set_err(B200_ERR_CUDA, "fused row-per-warp kernel variant is not registered");
    const size_t smem = rw_smem_per_warp(false, mode, nw, cap) * RW_WARPS;
    if (smem + k.static_smem > ctx->smem_optin) return set_err(B200_ERR_CUDA, "fused row-per-warp kernel needs %zu B of shared memory", smem);
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k.fn, RW_THREADS, smem) != cudaSuccess || per_sm < 1) { cudaGetLastError(); return set_err(B200_ERR_CUDA, "fused row-per-warp kernel does not fit an SM"); }
    const u64 want = (A->rows + RW_WARPS - 1) / RW_WARPS;
    int grid = (int)std::max<u64>(1, std::min<u64>(want, (u64)ctx->num_sms * per_sm));
    if (grid > ctx->num_sms) grid = grid / ctx->num_sms * ctx->num_sms;
    if ((u64)grid > ctx->cap_cta_tot) return set_err(B200_ERR_CUDA, "internal: %d CTAs exceed the per-CTA scratch", grid);
    const cudaError_t le = v64 ? rwf_go<u64>(ctx, A, B, C, ctrl, org, words, nw, cap, mirror, epoch, k.fn, grid, smem, s)
                               : rwf_go<u32>(ctx, A, B, C, ctrl, org, words, nw, cap, mirror, epoch, k.fn, grid, smem, s);
    ctx->launches++;
    if (ctx->trace) { fprintf(stderr, "[b200 trace] fused rw: grid %d (%d/SM) regs %d smem %zu nw %u cap %u org %u words %u\n", grid, per_sm, k.regs, smem, nw, cap, org, words); trace_mark(ctx, __LINE__); }
    if (le != cudaSuccess) return set_err(B200_ERR_CUDA, "fused row-per-warp kernel launch failed: %s", cudaGetErrorString(le));
    return B200_OK;
}

size_t rw_smem_per_warp(bool count, int mode, u32 nw, u32 cap) {
    return count ? (size_t)nw * 4 : (size_t)nw * 8 + (size_t)cap * ((mode == 0 ? 4 : 8) + 2);
}

template <typename VT>
static cudaError_t rw_go(b200_ctx *ctx, const b200_csr *A, const b200_csr *B, B200Ctrl *ctrl, int first_bin, int nbins, u32 nw, u32 cap,
                         b200_csr *C, const void *fn, int grid, size_t smem, cudaStream_t s) {
    RwArgs<VT> p;
    p.a = NumArgs<VT>{A->d_rp, A->d_col, (const VT *)A->d_val, B->d_desc, B->d_col, (const VT *)B->d_val};
    p.pack = B->d_pack; p.bin_rows = ctx->d_bin_rows; p.bin_cnt = ctrl->sym_bin_count; p.bin_stride = (u32)ctx->cap_rows;
    p.first_bin = first_bin; p.nbins = nbins; p.win = ctx->d_win; p.ncols = (u32)B->cols; p.nw = nw; p.cap = cap; p.nsm = (u32)ctx->num_sms;
    p.nnz_row = ctx->d_nnz_row; p.rpC = C ? C->d_rp : nullptr; p.colC = C ? C->d_col : nullptr; p.valC = C ? (VT *)C->d_val : nullptr; p.ctrl = ctrl;
    void *kargs[] = {(void *)&p};
    return cudaLaunchKernel(fn, dim3(grid), dim3(RW_THREADS), kargs, smem, s);
}

// Launch the row-per-warp kernel over the pre-pass lists first_bin .. first_bin + nbins - 1.  count: lengths into
// ctx->d_nnz_row; else values into C at C->d_rp.
int rw_launch(b200_ctx *ctx, const b200_csr *A, const b200_csr *B, B200Ctrl *ctrl, int first_bin, int nbins, bool count, int mode,
              bool packed, bool bpat, u32 nw, u32 cap, b200_csr *C, cudaStream_t s) {
    const bool v64 = A->val_bits == 64;
    const RwKernel &k = count ? g_rw[0][0][packed ? 1 : 0][1][1] : g_rw[v64 ? 1 : 0][v64 ? mode : std::min(mode, 1)][packed ? 1 : 0][bpat ? 1 : 0][0];
    if (!k.fn) return set_err(B200_ERR_CUDA, "row-per-warp kernel variant is not registered");
    const size_t smem = rw_smem_per_warp(count, mode, nw, cap) * RW_WARPS;
    if (smem + k.static_smem > ctx->smem_optin) return set_err(B200_ERR_CUDA, "row-per-warp kernel needs %zu B of shared memory", smem);
    const int by_smem = (int)((size_t)(228 * 1024) / (smem + k.static_smem + 1024));
    const int regs_per_cta = ((k.regs * 32 + 511) / 512 * 512) * RW_WARPS;
    const int by_regs = regs_per_cta ? 65536 / regs_per_cta : 32;
    const int per_sm = std::max(1, std::min(std::min(by_smem, by_regs), std::min(2048 / RW_THREADS, 32)));
    const u64 want = (A->rows + RW_WARPS - 1) / RW_WARPS;
    int grid = (int)std::max<u64>(1, std::min<u64>(want, (u64)ctx->num_sms * per_sm));
    if (grid > ctx->num_sms) grid = grid / ctx->num_sms * ctx->num_sms;     // whole waves: the slice mapping keeps an SM's CTAs adjacent
    // (the count variant is instantiated for u32 values only and never reads one)
    const cudaError_t le = v64 && !count ? rw_go<u64>(ctx, A, B, ctrl, first_bin, nbins, nw, cap, C, k.fn, grid, smem, s)
                                         : rw_go<u32>(ctx, A, B, ctrl, first_bin, nbins, nw, cap, C, k.fn, grid, smem, s);
    ctx->launches++;
    if (ctx->trace) trace_mark(ctx, __LINE__);
    if (le != cudaSuccess) return set_err(B200_ERR_CUDA, "row-per-warp kernel launch failed: %s", cudaGetErrorString(le));
    return B200_OK;
}
