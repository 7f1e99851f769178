// leftmul.cu -- C = A x B when the LEFT operand's rows are short and the right operand's rows are long (pipeline 6).
//
// Row i of C is then the union of a handful of long, sorted rows of B: B[k,:] for the few k in A[i,:].  That is the
// shape of a power chain multiplied from the left (A^k = A . A^(k-1), A of degree ~3, rows of A^(k-1) in the hundreds),
// and of any short-row selector / stencil operator applied to a denser matrix.  The row-wise kernels of rowwarp.cu walk
// such a multiply one gathered 32-byte record per A entry; here a product is one element of a contiguous list:
//   * a warp owns an output row; lane l < len(A[i,:]) holds list l's extent (row_ptr_B[k], row_ptr_B[k+1]) and a_ik;
//   * the lists are streamed in batches of U x 32 consecutive entries: a batch's loads are issued together, and two register
//     sets alternate so that the NEXT batch's loads are in flight while the current one is consumed -- fully coalesced, no
//     dependent gather per product at all; the slots of a batch are predicated instructions, not branches;
//   * count   : every column sets its bit in the warp's window bitmap (red.shared.or); length = popcount of the touched words;
//   * numeric : mark again, exclusive prefix popcount kept beside each word, then every product goes to acc[rank(column)]
//               (accumulators dense in rank space, as rowwarp.cu), emitted with coalesced stores in column order -- no sort.
// The whole multiply is ONE cooperative launch with the phases of k_rw_fused: lengths -> grid sync -> row_ptr -> values;
// C is written once.  Rows are handed out from a device counter (one moving front of neighbouring rows over the whole grid).  The window is either one arc for every row ({org, words}; the whole column space when nothing
// better is known) or, for square operands whose entry offsets (c - row) are bounded on both sides, a window that
// travels with the row: bit d of row i is column (i + org + d) mod n.
//
// Replaces, for such operands, the symbolic and numeric passes of CsrMatrix::matmul_par
// (/root/reference/src/graph_csr.rs:362-403, :430-476); results are bit-identical.
#include <type_traits>
#include <cooperative_groups.h>
#include "engine.cuh"
#include "devutil.cuh"

#define LM_WARPS 4
#define LM_THREADS (LM_WARPS * 32)
// Two builds of every kernel: TUNE 0 keeps deep batches (8 / 4 blocks of 32 entries per load batch of the mark / accumulate
// phases; 128 registers, 4 CTAs per SM), TUNE 1 trades batch depth for residency (4 / 2 blocks; 80 registers, 6 CTAs per SM).
template <int TUNE> struct LmTune;
template <> struct LmTune<0> { static constexpr int U = 8, UA = 4, CTAS = 4; };
template <> struct LmTune<1> { static constexpr int U = 4, UA = 2, CTAS = 6; };
template <> struct LmTune<2> { static constexpr int U = 8, UA = 4, CTAS = 5; };
#define FULLMASK 0xFFFFFFFFu

template <typename VT>
struct LmArgs {
    const u64 *rpA; const u32 *colA; const VT *valA;
    const u64 *rpB; const u32 *colB; const VT *valB;
    u64 rows; u32 ncols;
    u32 row_off, nact;              // active row arc of A: tickets 0 .. nact-1 are rows (row_off + t) mod rows, the others are empty
    u32 org;                        // window origin (per_row: added to the row index)
    u32 per_row;                    // 1: the window travels with the row
    u32 nw, cap, nsm;               // bitmap words per warp (multiple of 32), accumulator slots per warp, SMs
    u32 *nnz_row; u64 *cta_tot;
    u64 *rpC; u32 *colC; VT *valC;
    B200Ctrl *ctrl; u64 *host_mirror; u32 epoch; ull *maxval_dst;
};

__device__ __forceinline__ void lm_red_or(u32 addr, u32 v) { asm volatile("red.shared.or.b32 [%0], %1;" :: "r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void lm_red_add(u32 addr, u32 v) { asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void lm_st_u16(u32 addr, u32 v) { asm volatile("st.shared.u16 [%0], %1;" :: "r"(addr), "h"((unsigned short)v) : "memory"); }
// the same, predicated in the instruction (no branch around a slot of a batch)
__device__ __forceinline__ void lm_red_or_if(bool on, u32 addr, u32 v) { asm volatile("{ .reg .pred q; setp.ne.u32 q, %2, 0; @q red.shared.or.b32 [%0], %1; }" :: "r"(addr), "r"(v), "r"((u32)on) : "memory"); }
__device__ __forceinline__ void lm_red_add_if(bool on, u32 addr, u32 v) { asm volatile("{ .reg .pred q; setp.ne.u32 q, %2, 0; @q red.shared.add.u32 [%0], %1; }" :: "r"(addr), "r"(v), "r"((u32)on) : "memory"); }
// offs[pos] = d and acc[pos] += v under one predicate (32-bit sums)
__device__ __forceinline__ void lm_put_if(bool on, u32 offs_addr, u32 d, u32 acc_addr, u32 v) {
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %4, 0; @q st.shared.u16 [%0], %1; @q red.shared.add.u32 [%2], %3; }"
                 :: "r"(offs_addr), "h"((unsigned short)d), "r"(acc_addr), "r"(v), "r"((u32)on) : "memory");
}
__device__ __forceinline__ void lm_st_u16_if(bool on, u32 addr, u32 v) { asm volatile("{ .reg .pred q; setp.ne.u32 q, %2, 0; @q st.shared.u16 [%0], %1; }" :: "r"(addr), "h"((unsigned short)v), "r"((u32)on) : "memory"); }
__device__ __forceinline__ uint4 lm_ld_v4(u32 addr) { uint4 r; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr) : "memory"); return r; }
__device__ __forceinline__ void lm_st_v4(u32 addr, uint4 v) { asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
__device__ __forceinline__ void lm_st_v2(u32 addr, u32 a, u32 b) { asm volatile("st.shared.v2.u32 [%0], {%1,%2};" :: "r"(addr), "r"(a), "r"(b) : "memory"); }
__device__ __forceinline__ u32 lm_ld_u32(u32 addr) { u32 r; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(addr) : "memory"); return r; }
__device__ __forceinline__ u32 lm_ld_u16(u32 addr) { unsigned short r; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(r) : "r"(addr) : "memory"); return r; }
__device__ __forceinline__ u32 lm_ldg_u32(const u32 *p) { return __ldg(p); }

template <int MODE, typename VT>
__device__ __forceinline__ u64 lm_product(VT a, VT b) {
    if (MODE == 0) return (u64)((u32)a * (u32)b);
    if (MODE == 2) return sat_mul((u64)a, (u64)b);
    u64 x = (u64)a * (u64)b;
    if (sizeof(VT) == 4) x = x > 0xFFFFFFFFull ? 0xFFFFFFFFull : x;
    return x;
}

// window bit of column c: (c - org) mod ncols
__device__ __forceinline__ u32 lm_dcol(u32 c, u32 org, u32 ncols) {
    const u32 d = c - org;
    return c >= org ? d : d + ncols;
}

// The batches of one block of <= 32 lists (lane l < nl: list l starts at entry b0 of B's arrays and has bl entries), in
// order.  `sel(list)` yields what a batch needs of its list besides the entries (a_ik), `load(payload, start, base, len)`
// fetches entries base + 32 u + lane (u < U) of a list, `proc(payload, sel's value, base, len)` consumes them.  Two payloads
// alternate: while one batch is consumed the next one's loads are in flight, and no register is copied between them (a copy
// would wait for the loads it copies) -- the loop body exists twice, once per payload.  Everything warp-wide (the
// shuffles that pick the next list) happens before a batch's loads are issued.
struct LmCur { u32 j, base, len; u64 st; };
template <int U, typename PAY, typename SEL, typename LOAD, typename PROC>
__device__ __forceinline__ void lm_batches(u32 nl, u64 b0, u32 bl, SEL sel, LOAD load, PROC proc) {
    auto next_list = [&](LmCur &c) {                                      // c.j: first list to try
        c.base = 0; c.len = 0;
        for (; c.j < nl; c.j++) { c.len = __shfl_sync(FULLMASK, bl, c.j); if (c.len) break; }
        if (c.j < nl) c.st = shfl_u64(b0, c.j);
    };
    auto advance = [&](LmCur c) -> LmCur { c.base += 32u * U; if (c.base >= c.len) { c.j++; next_list(c); } return c; };
    LmCur a; a.j = 0; a.st = 0;
    next_list(a);
    if (a.j >= nl) return;
    PAY A, B;
    auto xa = sel(a.j);
    load(A, a.st, a.base, a.len);
    while (true) {
        LmCur b = advance(a);
        const bool mb = b.j < nl;
        auto xb = xa;
        if (mb) { xb = sel(b.j); load(B, b.st, b.base, b.len); }
        proc(A, xa, a.base, a.len);
        if (!mb) break;
        a = advance(b);
        const bool ma = a.j < nl;
        if (ma) { xa = sel(a.j); load(A, a.st, a.base, a.len); }
        proc(B, xb, b.base, b.len);
        if (!ma) break;
    }
}

// one lane's list of the row's first 32 A entries: first entry and length in B's arrays, a_ik
template <typename XT> struct LmDesc { u64 b0; u32 bl; XT xa; };
template <int U> struct LmCols { u32 c[U]; };
template <int U, typename XV> struct LmEntries { u32 c[U]; XV v[U]; };

// One warp's share of shared memory: bits u32[nw] | prefix u16[nw] | acc[cap] | window offsets u16[cap]   (nw % 32 == 0).
// Every phase leaves the words it touched zeroed.  The sweeps over the bitmap come in two builds:
//   FULL = false  a window much wider than a row (whole column space): the marks track the lowest / highest word touched, a lane
//                 sweeps an ODD number of consecutive words of that span (the lanes' accesses fall on different banks);
//   FULL = true   a window cut to the rows (travelling with the row): no tracking, a lane owns nw / 32 = 4 x odd consecutive
//                 words and sweeps them with 128-bit accesses (a quarter warp covers all banks).
template <typename VT, int MODE, int TUNE, bool FULL>
struct LmWarp {
    static constexpr int LM_U = LmTune<TUNE>::U, LM_UA = LmTune<TUNE>::UA;
    typedef typename std::conditional<MODE == 0, u32, VT>::type XT;          // what the accumulate phase keeps of a value
    const LmArgs<VT> &p;
    u32 *bits; unsigned short *pref; Acc<MODE> acc; unsigned short *offs;
    u32 sm_bits, sm_pref, sm_acc, sm_offs, nw, cap, ncols; int lane;

    __device__ __forceinline__ LmWarp(const LmArgs<VT> &p_, unsigned char *base, int lane_) : p(p_), nw(p_.nw), cap(p_.cap), ncols(p_.ncols), lane(lane_) {
        bits = reinterpret_cast<u32 *>(base); pref = reinterpret_cast<unsigned short *>(base + (size_t)nw * 4);
        acc.bind(base + (size_t)nw * 6, cap);
        offs = reinterpret_cast<unsigned short *>(base + (size_t)nw * 6 + Acc<MODE>::bytes(cap));
        sm_bits = (u32)__cvta_generic_to_shared(base); sm_pref = sm_bits + nw * 4; sm_acc = sm_bits + nw * 6; sm_offs = sm_acc + (u32)Acc<MODE>::bytes(cap);
    }
    static __host__ __device__ size_t bytes(u32 nw, u32 cap) { return (size_t)nw * 6 + Acc<MODE>::bytes(cap) + (size_t)cap * 2; }
    __device__ __forceinline__ void zero() {
        for (u32 t = lane; t < nw; t += 32) bits[t] = 0;
        for (u32 t = lane; t < cap; t += 32) acc.clear(t);
        __syncwarp();
    }

    // every product sets its column's bit; wlo / whi: lowest and highest bitmap word touched; returns this lane's share of
    // the row's intermediate products
    __device__ __forceinline__ u32 mark(const LmDesc<XT> &dsc, u64 rs, u32 lenA, u32 org, u32 &wlo, u32 &whi) {
        u32 psum = 0, lo = 0xFFFFFFFFu, hi = 0;
        for (u32 ab = 0; ab < lenA; ab += 32) {
            const u32 t = ab + (u32)lane;
            u64 b0 = dsc.b0; u32 bl = dsc.bl;                               // (the first block's came with the row)
            if (ab) { b0 = 0; bl = 0; if (t < lenA) { const u32 k = p.colA[rs + t]; b0 = p.rpB[k]; bl = (u32)(p.rpB[k + 1] - b0); } }
            psum += bl;
            lm_batches<LM_U, LmCols<LM_U>>(min(32u, lenA - ab), b0, bl,
                [&](u32) { return 0; },
                [&](LmCols<LM_U> &q, u64 st, u32 base, u32 len) {
                    const u32 *src = p.colB + st + base + lane;
#pragma unroll
                    for (int u = 0; u < LM_U; u++) if (base + 32u * u + lane < len) q.c[u] = lm_ldg_u32(src + 32 * u);
                },
                [&](const LmCols<LM_U> &q, int, u32 base, u32 len) {
#pragma unroll
                    for (int u = 0; u < LM_U; u++) {
                        const bool on = base + 32u * u + lane < len;
                        const u32 d = lm_dcol(q.c[u], org, ncols);
                        lm_red_or_if(on, sm_bits + (d >> 5) * 4u, __funnelshift_l(0u, 1u, d));
                        if (!FULL && on) { lo = min(lo, d); hi = max(hi, d); }
                    }
                });
        }
        if (FULL) { wlo = 0; whi = nw - 1; }
        else { wlo = __reduce_min_sync(FULLMASK, lo) >> 5; whi = __reduce_max_sync(FULLMASK, hi) >> 5; }
        return psum;
    }

    template <typename MID>
    __device__ __forceinline__ u32 count_row(const LmDesc<XT> &dsc, u64 rs, u32 lenA, u32 org, u32 &P, MID mid) {
        u32 wlo, whi;
        const u32 psum = mark(dsc, rs, lenA, org, wlo, whi);
        mid();
        __syncwarp();
        u32 mine = 0;
        if (FULL) {
            const u32 wpl = nw >> 5, a0 = sm_bits + (u32)lane * wpl * 4u;
            for (u32 i = 0; i < wpl; i += 4) {
                const uint4 b = lm_ld_v4(a0 + i * 4u);
                mine += __popc(b.x) + __popc(b.y) + __popc(b.z) + __popc(b.w);
                lm_st_v4(a0 + i * 4u, make_uint4(0u, 0u, 0u, 0u));
            }
        } else if (wlo <= whi) {
            const u32 wpl = ((whi - wlo + 32u) >> 5) | 1u, w0 = wlo + (u32)lane * wpl;
            for (u32 i = 0; i < wpl; i++) if (w0 + i <= whi) { mine += __popc(bits[w0 + i]); bits[w0 + i] = 0; }
        }
        P = warp_sum_u32(psum);
        const u32 nnz = warp_sum_u32(mine);
        __syncwarp();
        return nnz;
    }

    template <typename MID>
    __device__ __forceinline__ u32 numeric_row(const LmDesc<XT> &dsc, u64 rs, u32 lenA, u32 org, u32 *colp, VT *valp, u64 &vmax, MID mid) {
        u32 wlo, whi;
        mark(dsc, rs, lenA, org, wlo, whi);
        mid();
        __syncwarp();
        if (wlo > whi) return 0u;
        // ---- rank: consecutive words per lane over the touched span, warp scan of the lanes' popcounts
        const u32 wpl = FULL ? nw >> 5 : ((whi - wlo + 32u) >> 5) | 1u, w0 = wlo + (u32)lane * wpl;
        const u32 a0 = sm_bits + w0 * 4u, q0 = sm_pref + w0 * 2u;
        u32 mine = 0;
        if (FULL) {
            for (u32 i = 0; i < wpl; i += 4) { const uint4 b = lm_ld_v4(a0 + i * 4u); mine += __popc(b.x) + __popc(b.y) + __popc(b.z) + __popc(b.w); }
        } else {
            for (u32 i = 0; i < wpl; i++) if (w0 + i <= whi) mine += __popc(bits[w0 + i]);
        }
        u32 incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const u32 t = __shfl_up_sync(FULLMASK, incl, d); if (lane >= d) incl += t; }
        const u32 nnz = __shfl_sync(FULLMASK, incl, 31);
        if (FULL && nnz == 0) return 0u;
        u32 run = incl - mine;
        if (FULL) {
            for (u32 i = 0; i < wpl; i += 4) {
                const uint4 b = lm_ld_v4(a0 + i * 4u);
                const u32 r1 = run + __popc(b.x), r2 = r1 + __popc(b.y), r3 = r2 + __popc(b.z);
                lm_st_v2(q0 + i * 2u, run | (r1 << 16), r2 | (r3 << 16));
                run = r3 + __popc(b.w);
            }
        } else {
            for (u32 i = 0; i < wpl; i++) if (w0 + i <= whi) { pref[w0 + i] = (unsigned short)run; run += __popc(bits[w0 + i]); }
        }
        __syncwarp();
        // ranks are in d order; with an origin > 0 the entries whose column lies below it (d >= ncols - org) belong in FRONT
        // of the others: the row is written rotated by r0 = entries with d < ncols - org
        u32 r0 = nnz;
        if (org) {
            const u32 split = ncols - org, sw = split >> 5;
            if (sw < wlo) r0 = 0;
            else if (sw <= whi) r0 = (u32)pref[sw] + __popc(bits[sw] & (__funnelshift_l(0u, 1u, split) - 1u));
        }
        const u32 shift_hi = nnz - r0;
        typedef typename std::conditional<MODE == 0, u32, VT>::type XV;       // a B value as loaded (low word when 32-bit sums are proven)
        auto accumulate = [&](auto multi, u32 pass) {
            constexpr bool MULTI = decltype(multi)::value;
            for (u32 ab = 0; ab < lenA; ab += 32) {
                const u32 t = ab + (u32)lane;
                u64 b0 = dsc.b0; u32 bl = dsc.bl; XT xa = dsc.xa;
                if (ab) { b0 = 0; bl = 0; xa = 0; if (t < lenA) { const u32 k = p.colA[rs + t]; b0 = p.rpB[k]; bl = (u32)(p.rpB[k + 1] - b0); xa = (XT)p.valA[rs + t]; } }
                lm_batches<LM_UA, LmEntries<LM_UA, XV>>(min(32u, lenA - ab), b0, bl,
                    [&](u32 j) { return shfl_any(xa, (int)j); },
                    [&](LmEntries<LM_UA, XV> &q, u64 st, u32 base, u32 len) {
                        const u32 *src = p.colB + st + base + lane;
                        const XV *vsrc = reinterpret_cast<const XV *>(p.valB + st + base + lane);
                        constexpr int VS = sizeof(VT) / sizeof(XV);          // (little endian: the low word comes first)
#pragma unroll
                        for (int u = 0; u < LM_UA; u++) if (base + 32u * u + lane < len) { q.c[u] = lm_ldg_u32(src + 32 * u); q.v[u] = __ldg(vsrc + 32 * u * VS); }
                    },
                    [&](const LmEntries<LM_UA, XV> &q, XT x, u32 base, u32 len) {
                        u32 d[LM_UA], sb[LM_UA], sp[LM_UA];
#pragma unroll
                        for (int u = 0; u < LM_UA; u++) {
                            d[u] = base + 32u * u + lane < len ? lm_dcol(q.c[u], org, ncols) : 0u;     // (an idle slot reads word 0 and stores nothing)
                            sb[u] = lm_ld_u32(sm_bits + (d[u] >> 5) * 4u); sp[u] = lm_ld_u16(sm_pref + (d[u] >> 5) * 2u);
                        }
#pragma unroll
                        for (int u = 0; u < LM_UA; u++) {
                            const u32 pos = sp[u] + __popc(sb[u] & (__funnelshift_l(0u, 1u, d[u]) - 1u)) - pass;
                            const bool on = base + 32u * u + lane < len && (!MULTI || pos < cap);
                            if (MODE == 0) {
                                lm_put_if(on, sm_offs + pos * 2u, d[u], sm_acc + pos * 4u, (u32)x * (u32)q.v[u]);
                            } else if (on) {
                                lm_st_u16(sm_offs + pos * 2u, d[u]);
                                acc.addv(pos, lm_product<MODE, VT>((VT)x, (VT)q.v[u]));
                            }
                        }
                    });
            }
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
            accumulate(std::false_type{}, 0u);
            __syncwarp();
            emit(0u);
            __syncwarp();
        } else {
            for (u32 pass = 0; pass < nnz; pass += cap) {
                accumulate(std::true_type{}, pass);
                __syncwarp();
                emit(pass);
                __syncwarp();
            }
        }
        if (FULL) { for (u32 i = 0; i < wpl; i += 4) lm_st_v4(a0 + i * 4u, make_uint4(0u, 0u, 0u, 0u)); }
        else { for (u32 i = 0; i < wpl; i++) if (w0 + i <= whi) bits[w0 + i] = 0; }
        __syncwarp();
        return nnz;
    }
};

template <typename VT, int MODE, int TUNE, bool FULL>
__global__ void __launch_bounds__(LM_THREADS, LmTune<TUNE>::CTAS) k_lm(LmArgs<VT> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ u64 s_base; __shared__ ull s_P, s_maxP; __shared__ u32 s_maxN, s_last, s_scan[LM_WARPS + 1];
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const u32 nslices = gridDim.x, slice = blockIdx.x;
    const u64 S = (p.rows + nslices - 1) / nslices;                        // rows per slice of the row_ptr phase
    const u64 r_lo = min(p.rows, (u64)slice * S), r_hi = min(p.rows, r_lo + S);
    typedef LmWarp<VT, MODE, TUNE, FULL> W;
    W w(p, smem_raw + (size_t)wid * W::bytes(p.nw, p.cap), lane);
    w.zero();
    if (tid == 0) { s_P = 0; s_maxP = 0; s_maxN = 0; }
    __syncthreads();
    auto origin = [&](u64 row) -> u32 {
        if (!p.per_row) return p.org;
        const u64 t = row + p.org;
        return (u32)(t >= p.ncols ? t - p.ncols : t);
    };
    // Rows are handed out one at a time from a device counter (two tickets ahead, so the atomic's latency hides behind a
    // row): the warps of the whole grid work on one moving front of neighbouring rows -- their lists are neighbours too and
    // come from L2 -- and nobody waits at the end of a phase for a slice that happened to hold the long rows.
    auto grab = [&](u32 *ticket) -> u32 { return lane == 0 ? atomicAdd(ticket, 1u) : 0u; };
    // (tickets run over A's active row arc only: a whole-size left operand that holds a row block plus a halo hands out no empty rows)
    auto torow = [&](u32 t) -> u64 { if (t >= p.nact) return p.rows; const u64 r = (u64)t + p.row_off; return r >= p.rows ? r - p.rows : r; };
    auto active = [&](u64 r) -> bool { return (r >= p.row_off ? r - p.row_off : r + p.rows - p.row_off) < (u64)p.nact; };
    typedef typename W::XT XT;
    // lists of the row's first 32 A entries (loaded once per row and phase; mark and accumulate share them)
    auto lists = [&](u64 rs, u32 lenA, bool needv) -> LmDesc<XT> {
        LmDesc<XT> d; d.b0 = 0; d.bl = 0; d.xa = 0;
        if ((u32)lane < lenA) { const u32 k = p.colA[rs + lane]; if (needv) d.xa = (XT)p.valA[rs + lane]; d.b0 = p.rpB[k]; d.bl = (u32)(p.rpB[k + 1] - d.b0); }
        return d;
    };
    // ---- phase A: lengths
    {
        u64 Psum = 0; u32 maxP = 0, maxN = 0;
        u32 *ticket = &p.ctrl->scan_ticket[0];
        u64 row = torow(__shfl_sync(FULLMASK, grab(ticket), 0)), nxt = torow(__shfl_sync(FULLMASK, grab(ticket), 0));
        u64 rs = 0; u32 lenA = 0;
        if (row < p.rows) { rs = p.rpA[row]; lenA = (u32)(p.rpA[row + 1] - rs); }
        while (row < p.rows) {
            const u32 nn_l0 = grab(ticket);
            u64 rs_n = 0; u32 lenA_n = 0;
            auto mid = [&]() { if (nxt < p.rows) { rs_n = p.rpA[nxt]; lenA_n = (u32)(p.rpA[nxt + 1] - rs_n); } };
            u32 P = 0, nnz = 0;
            if (lenA) nnz = w.count_row(lists(rs, lenA, false), rs, lenA, origin(row), P, mid); else mid();
            if (lane == 0) { p.nnz_row[row] = nnz; if (nnz) atomicAdd((ull *)&p.cta_tot[row / S], (ull)nnz); }
            Psum += P; maxP = max(maxP, P); maxN = max(maxN, nnz);
            row = nxt; nxt = torow(__shfl_sync(FULLMASK, nn_l0, 0)); rs = rs_n; lenA = lenA_n;
        }
        if (lane == 0) { atomicAdd(&s_P, (ull)Psum); atomicMax(&s_maxP, (ull)maxP); atomicMax(&s_maxN, maxN); }
    }
    __syncthreads();
    if (tid == 0) {
        if (s_P) atomicAdd(&p.ctrl->total_products, s_P);
        atomicMax(&p.ctrl->max_row_products, s_maxP);
        atomicMax(&p.ctrl->max_row_nnz, (ull)s_maxN);
        __threadfence();
    }
    grid.sync();
    // ---- phase B: row_ptr of this CTA's slice of rows; its first entry = the totals of the slices before it
    if (wid == 0) {
        u64 sum = 0;
        for (u32 i = lane; i < slice; i += 32) sum += __ldcg(&p.cta_tot[i]);
        sum = warp_sum_u64(sum);
        if (lane == 0) s_base = sum;
    }
    __syncthreads();
    {
        u64 run = s_base;
        for (u64 r0 = r_lo; r0 < r_hi; r0 += LM_THREADS) {
            const u64 r = r0 + tid;
            const u32 v = r < r_hi && active(r) ? __ldcg(&p.nnz_row[r]) : 0u;      // (rows outside the arc were never counted)
            u32 incl = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const u32 t = __shfl_up_sync(FULLMASK, incl, d); if (lane >= d) incl += t; }
            if (lane == 31) s_scan[wid] = incl;
            __syncthreads();
            u32 wbase = 0, total = 0;
#pragma unroll
            for (int i = 0; i < LM_WARPS; i++) { if (i < wid) wbase += s_scan[i]; total += s_scan[i]; }
            if (r < r_hi) p.rpC[r] = run + wbase + incl - v;
            run += total;
            __syncthreads();
        }
        if (slice == nslices - 1 && tid == 0) { p.rpC[p.rows] = run; p.ctrl->total_nnz = run; }
    }
    grid.sync();
    // ---- phase C: values
    {
        u64 vmax = 0;
        u32 *ticket = &p.ctrl->scan_ticket[1];
        u64 row = torow(__shfl_sync(FULLMASK, grab(ticket), 0)), nxt = torow(__shfl_sync(FULLMASK, grab(ticket), 0));
        u64 rs = 0, obase = 0; u32 lenA = 0;
        if (row < p.rows) { rs = p.rpA[row]; lenA = (u32)(p.rpA[row + 1] - rs); obase = __ldcg(&p.rpC[row]); }
        while (row < p.rows) {
            const u32 nn_l0 = grab(ticket);
            u64 rs_n = 0, obase_n = 0; u32 lenA_n = 0;
            auto mid = [&]() { if (nxt < p.rows) { rs_n = p.rpA[nxt]; lenA_n = (u32)(p.rpA[nxt + 1] - rs_n); obase_n = __ldcg(&p.rpC[nxt]); } };
            if (lenA) w.numeric_row(lists(rs, lenA, true), rs, lenA, origin(row), p.colC + obase, p.valC + obase, vmax, mid); else mid();
            row = nxt; nxt = torow(__shfl_sync(FULLMASK, nn_l0, 0)); rs = rs_n; lenA = lenA_n; obase = obase_n;
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
        for (u32 i = tid; i < nslices; i += blockDim.x) p.cta_tot[i] = 0;     // every CTA is past its last read of the totals
    }
}

// ---------------------------------------------------------------------------- host side
struct LmKernel { const void *fn; int regs; size_t static_smem; };
static LmKernel g_lm[2][3][3][2];                                         // [value width][accumulator mode][tune][full sweeps]
template <typename VT, int MODE, int TUNE, bool FULL>
static void lm_register(size_t optin) {
    LmKernel &k = g_lm[sizeof(VT) == 8][MODE][TUNE][FULL];
    k.fn = (const void *)k_lm<VT, MODE, TUNE, FULL>;
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, k.fn) == cudaSuccess) { k.regs = fa.numRegs; k.static_smem = fa.sharedSizeBytes; } else { cudaGetLastError(); k.regs = 64; k.static_smem = 0; }
    if (cudaFuncSetAttribute(k.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(optin - k.static_smem)) != cudaSuccess) cudaGetLastError();
}
template <typename VT, int MODE> static void lm_register2(size_t o) { lm_register<VT, MODE, 0, false>(o); lm_register<VT, MODE, 1, false>(o); lm_register<VT, MODE, 0, true>(o); lm_register<VT, MODE, 1, true>(o); lm_register<VT, MODE, 2, false>(o); lm_register<VT, MODE, 2, true>(o); }
void lm_setup(b200_ctx *ctx) {
    const size_t o = ctx->smem_optin;
    lm_register2<u32, 0>(o); lm_register2<u32, 1>(o);
    lm_register2<u64, 0>(o); lm_register2<u64, 1>(o); lm_register2<u64, 2>(o);
}

size_t lm_smem_per_warp(int mode, u32 nw, u32 cap) { return (size_t)nw * 6 + (size_t)cap * ((mode == 0 ? 4 : 8) + 2); }

template <typename VT>
static cudaError_t lm_go(b200_ctx *ctx, const b200_csr *A, const b200_csr *B, b200_csr *C, B200Ctrl *ctrl, u32 org, bool per_row, u32 nw, u32 cap,
                         u64 *mirror, u32 epoch, const void *fn, int grid, size_t smem, cudaStream_t s) {
    LmArgs<VT> p;
    p.rpA = A->d_rp; p.colA = A->d_col; p.valA = (const VT *)A->d_val;
    p.rpB = B->d_rp; p.colB = B->d_col; p.valB = (const VT *)B->d_val;
    p.row_off = A->ar_len && A->ar_len < A->rows ? (u32)A->ar_start : 0u; p.nact = (u32)(A->ar_len && A->ar_len < A->rows ? A->ar_len : A->rows);
    p.rows = A->rows; p.ncols = (u32)B->cols; p.org = org; p.per_row = per_row ? 1u : 0u; p.nw = nw; p.cap = cap; p.nsm = (u32)ctx->num_sms;
    p.nnz_row = ctx->d_nnz_row; p.cta_tot = ctx->d_lm_tot; p.rpC = C->d_rp; p.colC = C->d_col; p.valC = (VT *)C->d_val;
    p.ctrl = ctrl; p.host_mirror = mirror; p.epoch = epoch; p.maxval_dst = C->d_maxval;
    void *kargs[] = {(void *)&p};
    return cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(LM_THREADS), kargs, smem, s);
}

// The whole multiply in one cooperative launch; C's arrays are already allocated.  Window: bit d of row i is column
// (org + (per_row ? i : 0) + d) mod cols, nw bitmap words per warp.  tune: -1 auto, 0 deep batches, 1 more resident warps.
int lm_launch(b200_ctx *ctx, const b200_csr *A, const b200_csr *B, b200_csr *C, B200Ctrl *ctrl, int mode, u32 org, bool per_row, u32 nw, u32 cap,
              int tune, u64 *mirror, u32 epoch, cudaStream_t s) {
    const bool v64 = A->val_bits == 64;
    const size_t smem = lm_smem_per_warp(mode, nw, cap) * LM_WARPS;
    // measured on the 30^3 chain (A x A^4 .. A x A^6): six lighter CTAs per SM win while a CTA needs <= 24 KB of shared memory
    // (0.132 vs 0.145 ms); above that the deep batches of tune 0 do (A x A^6: 0.248 vs 0.258 with five CTAs, 0.303 with six)
    if (tune < 0) tune = smem <= 24 * 1024 ? 1 : 0;
    if (tune > 2) tune = 0;
    const bool full = per_row && nw % 128 == 0 && (nw / 128) % 2 == 1;     // a travelling window is cut to the rows: sweep all of it, 128 bits at a time
    const LmKernel &k = g_lm[v64 ? 1 : 0][v64 ? mode : std::min(mode, 1)][tune][full ? 1 : 0];
    if (!k.fn) return set_err(B200_ERR_CUDA, "left-multiply kernel variant is not registered");
    if (smem + k.static_smem > ctx->smem_optin) return set_err(B200_ERR_CUDA, "left-multiply kernel needs %zu B of shared memory", smem);
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k.fn, LM_THREADS, smem) != cudaSuccess || per_sm < 1) { cudaGetLastError(); return set_err(B200_ERR_CUDA, "left-multiply kernel does not fit an SM"); }
    const u64 want = ((A->ar_len && A->ar_len < A->rows ? A->ar_len : A->rows) + LM_WARPS - 1) / LM_WARPS;
    int grid = (int)std::max<u64>(1, std::min<u64>(want, (u64)ctx->num_sms * per_sm));
    if (grid > ctx->num_sms) grid = grid / ctx->num_sms * ctx->num_sms;
    if ((u64)grid > ctx->cap_cta_tot) return set_err(B200_ERR_CUDA, "internal: %d CTAs exceed the per-CTA scratch", grid);
    // (the slice totals are zero: the last CTA of the previous left multiply left them so; after a failed multiply they are cleared here)
    if (ctx->lm_tot_dirty) { if (cudaMemsetAsync(ctx->d_lm_tot, 0, ctx->cap_cta_tot * 8, s) != cudaSuccess) return set_err(B200_ERR_CUDA, "clearing the slice totals failed"); ctx->lm_tot_dirty = false; }
    const cudaError_t le = v64 ? lm_go<u64>(ctx, A, B, C, ctrl, org, per_row, nw, cap, mirror, epoch, k.fn, grid, smem, s)
                               : lm_go<u32>(ctx, A, B, C, ctrl, org, per_row, nw, cap, mirror, epoch, k.fn, grid, smem, s);
    ctx->launches++;
    if (ctx->trace) { fprintf(stderr, "[b200 trace] left multiply: full %d tune %d grid %d (%d/SM) regs %d smem %zu nw %u cap %u org %u per_row %d\n", (int)full, tune, grid, per_sm, k.regs, smem, nw, cap, org, (int)per_row); trace_mark(ctx, __LINE__); }
    if (le != cudaSuccess) { ctx->lm_tot_dirty = true; return set_err(B200_ERR_CUDA, "left-multiply kernel launch failed: %s", cudaGetErrorString(le)); }
    return B200_OK;
}
