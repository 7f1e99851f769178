// devutil.cuh -- device-side building blocks shared by the kernel translation units (kernels.cuh, fused.cu, heavy.cu):
// operand views, accumulator storage of the three overflow modes, packed B records, the tiny-row warp network.
// Only templates and __forceinline__ device functions live here, so any number of .cu files may include it.
#pragma once
#include "common.cuh"

// decoupled look-back scans (k_prepass, k_scan_rowptr): tile status = 2 flag bits + 62-bit value
#define SCAN_THREADS 256
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)
#define SCAN_FLAG_AGG (1ull << 62)
#define SCAN_FLAG_PRE (2ull << 62)
#define SCAN_VAL_MASK ((1ull << 62) - 1)

// =======================================================================================
// 0. operand views
// =======================================================================================
struct SymArgs {                 // what the symbolic pass touches (value-type independent)
    const u64 *rpA; const u32 *colA;
    const uint2 *bdesc;          // per B row: x = first entry (u32), y = length
    const u32 *colB;
};
template <typename VT>
struct NumArgs {
    const u64 *rpA; const u32 *colA; const VT *valA;
    const uint2 *bdesc; const u32 *colB; const VT *valB;
};

// Where a numeric kernel puts its rows.  Exact mode: base = row_ptr_C (from the count pass), nnz_out = null.
// Scratch mode: base = offsets from the scan of the per-row bounds min(P_i, cols), rows land in a scratch CSR and their
// exact lengths in nnz_out.  Either way the rows to process come from the pre-pass's per-bin lists.
template <typename VT>
struct OutArgs {
    const u64 *base; u32 *col; VT *val;
    u32 *nnz_out;                // may be null (exact mode)
    const u32 *bin_cnt;          // list sizes (ctrl->sym_bin_count)
    u32 bin_stride;              // bin b's row list starts at bin_rows + b * bin_stride
    u32 narrow;                  // scratch mode, u64 values proven < 2^32 (accumulator mode 0): the scratch holds them as
                                 // u32 (val is then a u32 array); the compaction kernel widens them on the way into C
};
template <typename VT>
__device__ __forceinline__ void put_val(const OutArgs<VT> &o, u64 idx, VT v) {
    if (sizeof(VT) == 8 && o.narrow) reinterpret_cast<u32 *>(o.val)[idx] = (u32)v;
    else o.val[idx] = v;
}

// Rows of the pre-pass's heavy list that the chunked kernels (heavy.cu) take; the other heavy-row kernels leave them alone.
// pmin = ~0: nothing is taken.
struct HvSkip { const u64 *prod; u64 pmin; u32 cap_li; };
__device__ __forceinline__ bool hv_skipped(const HvSkip &h, u32 li, u32 row) { return h.pmin != ~0ull && li < h.cap_li && h.prod[row] >= h.pmin; }

// Walk the intermediate products of one A row.  `ngrp` groups of G lanes each take A entries
// grp, grp+ngrp, ...; the G lanes of a group stride over that entry's B row.  Two A entries are
// in flight per iteration so their dependent loads (A.col -> desc -> B.col) overlap.
template <typename CT, typename P, typename F>
__device__ __forceinline__ void walk_products(const u32 *__restrict__ Ac, u32 lenA, const uint2 *__restrict__ bdesc, u32 grp,
                                              u32 ngrp, u32 sub, u32 G, P pre, F f) {
    for (u32 t = grp; t < lenA; t += 2 * ngrp) {
        const u32 t1 = t + ngrp;
        const bool has1 = t1 < lenA;
        const u32 k0 = Ac[t];
        const u32 k1 = has1 ? Ac[t1] : k0;
        const uint2 d0 = bdesc[k0];
        uint2 d1 = bdesc[k1];
        if (!has1) d1.y = 0;
        const CT c0 = pre(t);
        const CT c1 = pre(has1 ? t1 : t);
        for (u32 j = sub; j < d0.y; j += G) f(c0, d0.x + j);
        for (u32 j = sub; j < d1.y; j += G) f(c1, d1.x + j);
    }
}

// =======================================================================================
// 2. accumulator storage shared by the numeric kernels
// =======================================================================================
// MODE 0: 32-bit sums (host proved max_row_products*max(A)*max(B) < 2^32)
// MODE 1: 64-bit sums kept as two u32 words; the carry out of `lo` is recovered from the value
//         the atomic returns.  (Shared-memory u64 atomicAdd compiles to a CAS spin loop,
//         ATOMS.CAST.SPIN.64, measured 7x slower than ATOMS.ADD on B200.)  For u32 values the
//         products are clamped to 2^32-1 first; < 2^32 of them per row cannot wrap 64 bits.
// MODE 2: u64 saturating multiply + CAS-loop saturating add.
template <int MODE> struct Acc;
template <> struct Acc<0> {
    u32 *lo;
    static __host__ __device__ size_t bytes(u32 n) { return (size_t)n * 4; }
    __device__ void bind(unsigned char *base, u32) { lo = reinterpret_cast<u32 *>(base); }
    __device__ void clear(u32 i) { lo[i] = 0; }
    template <typename VT> __device__ void add(u32 i, VT a, VT b) { atomicAdd(&lo[i], (u32)a * (u32)b); }
    __device__ void addv(u32 i, u64 x) { atomicAdd(&lo[i], (u32)x); }                 // x: an already formed product
    __device__ u64 get(u32 i) const { return lo[i]; }
    __device__ void set(u32 i, u64 x) { lo[i] = (u32)x; }
};
template <> struct Acc<1> {
    u32 *lo, *hi;
    static __host__ __device__ size_t bytes(u32 n) { return (size_t)n * 8; }
    __device__ void bind(unsigned char *base, u32 n) { lo = reinterpret_cast<u32 *>(base); hi = lo + n; }
    __device__ void clear(u32 i) { lo[i] = 0; hi[i] = 0; }
    template <typename VT> __device__ void add(u32 i, VT a, VT b) {
        u64 x = (u64)a * (u64)b;
        if (sizeof(VT) == 4) x = x > 0xFFFFFFFFull ? 0xFFFFFFFFull : x;
        const u32 xlo = (u32)x, xhi = (u32)(x >> 32);
        const u32 old = atomicAdd(&lo[i], xlo);
        const u32 up = xhi + ((u32)(old + xlo) < xlo ? 1u : 0u);
        if (up) atomicAdd(&hi[i], up);
    }
    __device__ void addv(u32 i, u64 x) {
        const u32 xlo = (u32)x, xhi = (u32)(x >> 32);
        const u32 old = atomicAdd(&lo[i], xlo);
        const u32 up = xhi + ((u32)(old + xlo) < xlo ? 1u : 0u);
        if (up) atomicAdd(&hi[i], up);
    }
    __device__ u64 get(u32 i) const { return ((u64)hi[i] << 32) | lo[i]; }
    __device__ void set(u32 i, u64 x) { lo[i] = (u32)x; hi[i] = (u32)(x >> 32); }
};
template <> struct Acc<2> {
    ull *v;
    static __host__ __device__ size_t bytes(u32 n) { return (size_t)n * 8; }
    __device__ void bind(unsigned char *base, u32) { v = reinterpret_cast<ull *>(base); }
    __device__ void clear(u32 i) { v[i] = 0; }
    template <typename VT> __device__ void add(u32 i, VT a, VT b) {
        const ull x = sat_mul((u64)a, (u64)b);
        ull old = *reinterpret_cast<volatile ull *>(&v[i]), assumed;
        do {
            assumed = old;
            ull s = assumed + x; if (s < assumed) s = ~0ull;
            if (s == assumed) break;
            old = atomicCAS(&v[i], assumed, s);
        } while (old != assumed);
    }
    __device__ void addv(u32 i, u64 x) {
        ull old = *reinterpret_cast<volatile ull *>(&v[i]), assumed;
        do {
            assumed = old;
            ull s = assumed + x; if (s < assumed) s = ~0ull;
            if (s == assumed) break;
            old = atomicCAS(&v[i], assumed, s);
        } while (old != assumed);
    }
    __device__ u64 get(u32 i) const { return v[i]; }
    __device__ void set(u32 i, u64 x) { v[i] = x; }
};
template <typename VT> __device__ __forceinline__ VT emit_val(u64 v) {
    if (sizeof(VT) == 4 && v > 0xFFFFFFFFull) v = 0xFFFFFFFFull;   // u32 sums accumulated in 64 bits saturate here
    return (VT)v;
}

// Sector-packed right operand (low-degree B): one 32-byte record per B row,
//     rec[2k] = {start, len, col0, col1}, rec[2k+1] = {col2, col3, col4, col5}.
// One aligned 32-byte fetch (a single L2 sector, one LDG.E.256) returns a B row's length AND its first six columns,
// instead of a descriptor sector plus a column sector; the dependent-load chain of a product shrinks from
// A.col -> desc -> B.col to A.col -> record.  Rows longer than six columns continue in B.col.
#define B200_PACK_INLINE 6
struct PackRec { uint4 a, b; };

__device__ __forceinline__ PackRec load_pack(const uint4 *__restrict__ pack, u32 k) {
    // one 256-bit read-only load (LDG.E.256 on sm_100): a single L1 request per record instead of two
    PackRec r;
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.a.x), "=r"(r.a.y), "=r"(r.a.z), "=r"(r.a.w), "=r"(r.b.x), "=r"(r.b.y), "=r"(r.b.z), "=r"(r.b.w)
                 : "l"(pack + 2 * (u64)k));
    return r;
}

// =======================================================================================
// 3. tiny rows: one warp per row, <= 32 products held one per lane
// =======================================================================================
// One tiny row per warp.  The caller has already loaded the row's A entries (lane < dA holds column k and value a):
// the kernels below run a software pipeline over their rows -- row ids two rows ahead, row_ptr / output base and the
// A entries one row ahead -- so that an iteration only waits for its own desc -> B.col gathers.
template <typename VT, bool NUMERIC>
__device__ __forceinline__ u32 tiny_gather(u32 dA, u32 k, VT a, const uint2 *bdesc, const u32 *colB, const VT *valB, int lane,
                                           u32 &key, VT &val, bool bpat = false) {
    // (a sector-packed record per entry was tried here: the 32-byte gathers and six shuffles cost more than the
    //  descriptor + column gathers they replace -- 100^3 torus A^2 0.79 -> 0.87 ms -- so tiny rows keep bdesc/colB)
    u32 deg = 0, bstart = 0;
    if (lane < (int)dA) { const uint2 d = bdesc[k]; bstart = d.x; deg = d.y; }
    u32 incl = deg;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { u32 t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
    const u32 excl = incl - deg;
    const u32 P = __shfl_sync(0xFFFFFFFFu, incl, 31);
    // lane p: last entry e with excl_e <= p (empty B rows share their successor's offset and lose)
    int lo = 0;
#pragma unroll
    for (int step = 16; step > 0; step >>= 1) {
        const int cand = lo + step;
        const u32 t = __shfl_sync(0xFFFFFFFFu, excl, cand & 31);
        if (cand < (int)dA && t <= (u32)lane) lo = cand;
    }
    const u32 e_excl = __shfl_sync(0xFFFFFFFFu, excl, lo);
    const u32 e_bstart = __shfl_sync(0xFFFFFFFFu, bstart, lo);
    VT e_a = 0;
    if (NUMERIC) e_a = shfl_any(a, lo);
    key = B200_EMPTY_KEY; val = 0;
    if ((u32)lane < P) {
        const u32 j = e_bstart + ((u32)lane - e_excl);
        key = colB[j];
        if (NUMERIC) val = bpat ? e_a : sat_mul(e_a, valB[j]);
    }
    return P;
}

// bitonic sort of (key[,val]) across the 32 lanes of a warp, ascending by key
template <typename VT, bool NUMERIC>
__device__ __forceinline__ void warp_bitonic(u32 &key, VT &val, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const u32 ok = __shfl_xor_sync(0xFFFFFFFFu, key, j);
            VT ov = 0;
            if (NUMERIC) ov = shfl_xor_any(val, j);
            const bool up = ((lane & k) == 0);
            const bool lower = ((lane & j) == 0);
            const bool take_min = (up == lower);
            const bool swap = take_min ? (ok < key) : (ok > key);
            if (swap) { key = ok; if (NUMERIC) val = ov; }
        }
    }
}
