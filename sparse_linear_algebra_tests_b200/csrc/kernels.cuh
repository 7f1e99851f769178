// kernels.cuh -- device kernels of the SpGEMM engine (sm_100a).
//
// Pipeline of one C = A x B (host orchestration in api.cu):
//   k_build_desc / k_build_pack   per right operand, cached: row descriptors, column spans, sector-packed rows
//   k_prepass           product count P_i and column window per row, bin lists, scratch offsets (look-back scan)
//   k_sym_*             exact mode only: distinct output columns per row (count-only twins of the numeric kernels)
//   k_num_*             values: saturating accumulate, in-row column order, write col/val
//   k_scan_rowptr       decoupled look-back scan of the exact row lengths -> row_ptr_C (u64); its last CTA reports the
//                       control block to pinned host memory as {word, epoch} chunks (the host never synchronises)
//   k_compact_rows      scratch mode only: rows from their bound offsets to their exact places (widening the scratch's
//                       32-bit values when mode 0 let them travel narrow); leaves the scan area zeroed for the next multiply
// Row classes: tiny (warp register merge), window-bitmap rank (k_num_expand: CTA per row, no sort), hash + bitonic sort
// (warp or CTA per row, any column space), heavy (global-memory table).  All inner loops use 32-bit offsets; 64-bit only
// for row bases.  Column windows of the bitmap kernel: whole column space, one operand-level arc of the index circle
// (row blocks of a torus / banded matrix), or per-row plain / circular windows from the pre-pass (api.cu decides).
#pragma once
#include "devutil.cuh"


// desc = {first entry, length}; span = {length, first column, last column, -} (what the one-pass pre-pass gathers: the
// product count and the column window of a row of C follow from these alone because B's rows are sorted)
__global__ void __launch_bounds__(256) k_build_desc(u64 rows, const u64 *__restrict__ rp, const u32 *__restrict__ col,
                                                    uint2 *__restrict__ desc, uint4 *__restrict__ span) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += (u64)gridDim.x * blockDim.x) {
        const u64 s = rp[i], e = rp[i + 1];
        desc[i] = make_uint2((u32)s, (u32)(e - s));
        span[i] = e > s ? make_uint4((u32)(e - s), col[s], col[e - 1], 0u) : make_uint4(0u, 0xFFFFFFFFu, 0u, 0u);
    }
}

// Circular spans (square B only): for B row k the offsets of its columns from the diagonal, u = (c - k + n/2) mod n,
// as {length, min u, max u}.  A torus row that wraps around the end of the index space has a plain span of ~n columns but
// a circular span as narrow as any other row's.  Rows longer than 64 entries are not scanned: they get the full circle
// (any row of C that touches them is sent to the hash kernels, where long, spread-out rows belong anyway).
__global__ void __launch_bounds__(256) k_build_cspan(u64 n, const u64 *__restrict__ rp, const u32 *__restrict__ col, uint4 *__restrict__ cspan) {
    const u32 half = (u32)(n / 2);
    for (u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (u64)gridDim.x * blockDim.x) {
        const u64 s = rp[k], e = rp[k + 1];
        const u32 len = (u32)(e - s);
        u32 umin = 0xFFFFFFFFu, umax = 0;
        if (len > 64) { umin = 0; umax = (u32)(n - 1); }
        else for (u64 j = s; j < e; j++) {
            long long t = (long long)col[j] - (long long)k + (long long)half;
            if (t < 0) t += (long long)n; else if (t >= (long long)n) t -= (long long)n;
            umin = min(umin, (u32)t); umax = max(umax, (u32)t);
        }
        cspan[k] = make_uint4(len, umin, umax, 0u);
    }
}

// per hash bin: how many 128-column groups the bin's shared-memory bitmap holds (0: the bin has no bitmap kernel)
// `full`: groups that hold any row of this multiply (bins with cap >= full have no wide list)
struct WinCaps { u32 cap[B200_NUM_HASH_BINS]; u32 full; u32 heavy_from; };   // heavy_from: rows with at least this many products join the heavy list (chunked kernels; default 8193)


// =======================================================================================
// 1. product count per row + symbolic-bin histogram
// =======================================================================================
// bin of a row by its product count P (rows without products need no kernel)
__device__ __forceinline__ int sym_bin_of(u64 p, u64 dA) {
    if (p == 0) return B200_BIN_NONE;
    if (p <= 32 && dA <= 32) return B200_BIN_TINY;
    return b200_bin_by_size(p);
}

// Per-row intermediate-product counts only (b200_row_products / product-balanced sharding); G lanes per row.
template <int G>
__global__ void __launch_bounds__(256) k_row_products(u64 rows, const u64 *__restrict__ rpA, const u32 *__restrict__ colA,
                                                      const uint2 *__restrict__ bdesc, u64 *__restrict__ prod) {
    const u64 gtid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 row = gtid / G;
    const u32 sub = threadIdx.x % G;
    u64 p = 0;
    if (row < rows) {
        const u64 s = rpA[row];
        const u32 lenA = (u32)(rpA[row + 1] - s);
        const u32 *Ac = colA + s;
        for (u32 i = sub; i < lenA; i += G) p += bdesc[Ac[i]].y;
    }
#pragma unroll
    for (int m = G / 2; m > 0; m >>= 1) p += shfl_xor_u64(p, m);
    if (row < rows && sub == 0) prod[row] = p;
}

// One-pass pre-pass, one launch: product count P_i per row, totals, the bin lists (block-wise reservation in every
// bin's own list, so a CTA's rows stay consecutive) and -- through a decoupled look-back over the CTAs in ticket
// order -- the scratch offsets prefix(min(P_i, cols)) that the numeric kernels write their rows at.
// WMODE 0 (one window fits every row of this multiply -- the whole column space, or the arc of the index circle the host
// derived from the operands' column ranges): lengths come from the 8-byte descriptors and every row gets the window
// {0, all_groups, all_rot} -- half the gather bytes, no min/max reductions.  WMODE 1: plain window [cmin, cmax] from the sorted B
// rows' first/last columns.  WMODE 2 (square B, `bspan` holds circular spans): columns are measured from a per-row
// reference ref = the row's first A column, d(c) = (c - ref + n/2) mod n, so rows that wrap around the index space keep
// a narrow window; win = {window base in d (multiple of 128), 128-column groups, rot = (ref - n/2) mod n, 0}.
#define B200_PREPASS_MIN_ROWS 32
#define B200_PREPASS_ROWS(G) ((256 / (G)) > B200_PREPASS_MIN_ROWS ? (256 / (G)) : B200_PREPASS_MIN_ROWS)   // rows per CTA
template <int G, int WMODE>
__global__ void __launch_bounds__(256) k_prepass(u64 rows, const u64 *__restrict__ rpA, const u32 *__restrict__ colA,
                                                 const uint4 *__restrict__ bspan, const uint2 *__restrict__ bdesc, u64 ncols, u64 *__restrict__ prod,
                                                 u32 *__restrict__ nnz_row, u64 *__restrict__ tmp_ptr, u64 *tile_status,
                                                 B200Ctrl *ctrl, u32 *__restrict__ bin_rows, u32 bin_stride,
                                                 uint4 *__restrict__ win, WinCaps caps, u32 all_groups, u32 all_rot) {
    // G lanes per row, 256/G rows per step, enough (unrolled) steps for >= 32 rows per CTA (16, 64 and 128 measured slower): ncu showed half of this
    // kernel's stall samples at the barrier behind the look-back when every CTA was a tile of 8 rows (3375 tiles for
    // the 30^3 torus); fewer, larger tiles shorten that chain and the unrolled steps keep several rows' gathers in flight.
    constexpr int RPS = 256 / G;                                            // rows per step
    constexpr int TILE = B200_PREPASS_ROWS(G);
    constexpr int STEPS = TILE / RPS;
    constexpr bool WINDOWS = WMODE != 0;
    __shared__ u32 s_tile, s_cnt[B200_NBINS], s_base[B200_NBINS], s_binloc[TILE];
    __shared__ u64 s_bound[TILE], s_wsum[8], s_excl;
    __shared__ ull s_sum, s_max;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) { s_tile = atomicAdd(&ctrl->scan_ticket[1], 1u); s_sum = 0; s_max = 0; }
    if (tid < B200_NBINS) s_cnt[tid] = 0;
    __syncthreads();
    const u32 tile = s_tile;
    const u32 sub = tid % G;
    const long long n = (long long)ncols, half = n / 2;
    u64 wsum = 0, wmax = 0;
#pragma unroll
    for (int step = 0; step < STEPS; step++) {
        const u32 lrow = step * RPS + tid / G;                               // row inside the tile
        const u64 row = (u64)tile * TILE + lrow;
        u64 p = 0; u32 lenA = 0, cmin = 0xFFFFFFFFu, cmax = 0, rot = 0;
        if (row < rows) {
            const u64 s = rpA[row];
            lenA = (u32)(rpA[row + 1] - s);
            const u32 *Ac = colA + s;
            const long long ref = WMODE == 2 && lenA ? (long long)Ac[0] : 0;
            if (WMODE == 2) { long long t = ref - half; if (t < 0) t += n; rot = (u32)t; }
            // one entry's contribution to the window: plain [first, last], or its circular span shifted by (k - ref)
            auto widen = [&](u32 k, const uint4 &d) {
                if (d.x == 0) return;
                if (WMODE == 1) { cmin = min(cmin, d.y); cmax = max(cmax, d.z); return; }
                long long kap = (long long)k - ref + half;                   // (k - ref + n/2) mod n
                if (kap < 0) kap += n; else if (kap >= n) kap -= n;
                const long long lo = kap + (long long)d.y - half, hi = kap + (long long)d.z - half;
                if (lo < 0 || hi >= n) { cmin = 0; cmax = (u32)(n - 1); }     // offsets add up past half the circle: full window
                else { cmin = min(cmin, (u32)lo); cmax = max(cmax, (u32)hi); }
            };
            u32 i = sub;
            for (; i + 3 * G < lenA; i += 4 * G) {                           // four independent gathers in flight
                const u32 k0 = Ac[i], k1 = Ac[i + G], k2 = Ac[i + 2 * G], k3 = Ac[i + 3 * G];
                if (WINDOWS) {
                    const uint4 d0 = bspan[k0], d1 = bspan[k1], d2 = bspan[k2], d3 = bspan[k3];
                    p += (u64)d0.x + d1.x + d2.x + d3.x;
                    widen(k0, d0); widen(k1, d1); widen(k2, d2); widen(k3, d3);
                } else {
                    const u32 d0 = bdesc[k0].y, d1 = bdesc[k1].y, d2 = bdesc[k2].y, d3 = bdesc[k3].y;
                    p += (u64)d0 + d1 + d2 + d3;
                }
            }
            for (; i < lenA; i += G) {
                if (WINDOWS) { const u32 k = Ac[i]; const uint4 d = bspan[k]; p += d.x; widen(k, d); }
                else p += bdesc[Ac[i]].y;
            }
        }
#pragma unroll
        for (int m = G / 2; m > 0; m >>= 1) {
            p += shfl_xor_u64(p, m);
            if (WINDOWS) {
                cmin = min(cmin, __shfl_xor_sync(0xFFFFFFFFu, cmin, m));
                cmax = max(cmax, __shfl_xor_sync(0xFFFFFFFFu, cmax, m));
            }
        }
        if (!WINDOWS) { cmin = 0; cmax = all_groups * 128u - 1u; rot = all_rot; }   // one window for every row, chosen by the host
        if (sub == 0) {
            u64 bound = 0; u32 binloc = (u32)B200_BIN_NONE << 24;
            if (row < rows) {
                prod[row] = p;
                bound = p < ncols ? p : ncols;
                int b = sym_bin_of(p, lenA);
                if (p >= (u64)caps.heavy_from && b != B200_BIN_NONE) b = B200_BIN_HEAVY;
                if (b == B200_BIN_HASH0) b = B200_BIN_HASH0 + 1;           // the two smallest hash bins share a list
                if (b != B200_BIN_NONE) {
                    // column window of the row, in 128-column groups; rows too wide for their bin's bitmap go to the hash list
                    const u32 base = cmin & ~127u;
                    const u32 groups = (cmax - base) / 128u + 1u;
                    win[row] = make_uint4(base, groups, rot, 0u);
                    if (b >= B200_BIN_HASH0 && b < B200_BIN_HEAVY && groups > caps.cap[b - B200_BIN_HASH0]) b = B200_BIN_WIDE0 + (b - B200_BIN_HASH0);
                    binloc = ((u32)b << 24) | atomicAdd(&s_cnt[b], 1u);
                } else nnz_row[row] = 0;
                wsum += p; wmax = wmax > p ? wmax : p;
            }
            s_bound[lrow] = bound; s_binloc[lrow] = binloc;
        }
    }
    wsum = warp_sum_u64(wsum); wmax = warp_max_u64(wmax);
    if (lane == 0) { if (wsum) atomicAdd(&s_sum, (ull)wsum); atomicMax(&s_max, (ull)wmax); }
    __syncthreads();
    // ---- exclusive scan of the bounds inside the CTA (row order), aggregate, look-back
    const u64 mine = tid < TILE ? s_bound[tid] : 0ull;
    u64 incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const u64 t = shfl_up_u64(incl, d); if (lane >= d) incl += t; }
    if (lane == 31) s_wsum[w] = incl;
    if (tid < B200_NBINS) { const u32 c = s_cnt[tid]; s_base[tid] = c ? atomicAdd(&ctrl->sym_bin_count[tid], c) : 0u; }
    __syncthreads();
    u64 wbase = 0, agg = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { if (i < w) wbase += s_wsum[i]; agg += s_wsum[i]; }
    if (w == 0) {
        if (lane == 0) atomicExch((ull *)&tile_status[tile], (ull)((tile == 0 ? SCAN_FLAG_PRE : SCAN_FLAG_AGG) | agg));
        u64 excl = 0;
        if (tile > 0) {
            int look = (int)tile - 1;
            while (true) {
                const int idx = look - lane;
                u64 st;
                do { st = idx >= 0 ? ld_volatile_u64(&tile_status[idx]) : SCAN_FLAG_PRE; } while (__any_sync(0xFFFFFFFFu, (st >> 62) == 0));
                const u32 pre_mask = __ballot_sync(0xFFFFFFFFu, (st >> 62) == 2);
                const int first = pre_mask ? __ffs(pre_mask) - 1 : 32;
                excl += warp_sum_u64(lane <= first ? (st & SCAN_VAL_MASK) : 0ull);
                if (pre_mask) break;
                look -= 32;
            }
            if (lane == 0) atomicExch((ull *)&tile_status[tile], (ull)(SCAN_FLAG_PRE | (excl + agg)));
        }
        if (lane == 0) s_excl = excl;
    }
    __syncthreads();
    if (tid < TILE) {
        const u64 r = (u64)tile * TILE + tid;
        if (r < rows) {
            const u64 run = s_excl + wbase + incl;
            tmp_ptr[r + 1] = run;
            if (r + 1 == rows) ctrl->total_bound = run;
            const u32 bl = s_binloc[tid], b = bl >> 24;
            if (b != B200_BIN_NONE) bin_rows[(u64)b * bin_stride + s_base[b] + (bl & 0xFFFFFFu)] = (u32)r;
        }
        if (r == 0) tmp_ptr[0] = 0;
    }
    if (tid == 0) {
        if (s_sum) atomicAdd(&ctrl->total_products, s_sum);
        if (s_max) atomicMax(&ctrl->max_row_products, s_max);
    }
}

// r-th row of the run of `nbins` consecutive bins starting at first_bin (every bin has its own list)
__device__ __forceinline__ u32 bin_row_at(const u32 *__restrict__ bin_rows, const u32 *cnt, u32 stride, int first_bin, int nbins, u32 r) {
    int b = first_bin;
    while (nbins > 1 && r >= cnt[b]) { r -= cnt[b]; b++; nbins--; }
    return bin_rows[(u64)b * stride + r];
}


// insert column c into an open-addressing key table; returns the slot.  `fresh` = first insertion.
__device__ __forceinline__ u32 table_insert(u32 *keys, u32 mask, int shift, u32 c, bool &fresh) {
    u32 h = b200_hash(c, shift);
    fresh = false;
    while (true) {
        u32 cur = ld_volatile_u32(&keys[h]);
        if (cur == B200_EMPTY_KEY) {
            cur = atomicCAS(&keys[h], B200_EMPTY_KEY, c);
            if (cur == B200_EMPTY_KEY) { fresh = true; return h; }
        }
        if (cur == c) return h;
        h = (h + 1) & mask;
    }
}

// block-wide exclusive scan of one u32 per thread; returns exclusive prefix, total in `total`
__device__ __forceinline__ u32 block_excl_scan(u32 v, u32 *s_warp /*>=33*/, u32 &total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    u32 incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { u32 t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
    if (lane == 31) s_warp[w] = incl;
    __syncthreads();
    if (w == 0) {
        u32 x = lane < nw ? s_warp[lane] : 0, xi = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { u32 t = __shfl_up_sync(0xFFFFFFFFu, xi, d); if (lane >= d) xi += t; }
        s_warp[lane] = xi - x;
        if (lane == 31) s_warp[32] = xi;
    }
    __syncthreads();
    total = s_warp[32];
    const u32 r = s_warp[w] + incl - v;
    __syncthreads();
    return r;
}

// Balanced enumeration of one A row's intermediate products by a whole CTA.  A tile of blockDim.x A entries is
// loaded (one per thread), a block scan of their B-row lengths lays the tile's products out on a line, and thread q
// takes products q, q + blockDim.x, ...: a binary search in the scanned lengths finds the entry, the remainder the
// position inside its B row.  Every thread gets the same number of products whatever the row lengths are (a single
// long B row no longer pins one lane group while the rest of the CTA waits at the barrier), and consecutive threads
// read consecutive B entries.  f(p, jb, a_ik): p = running product index in the row, jb = index into B's arrays.
struct EnumSmem { u32 pre[1025]; u32 start[1024]; u64 av[1024]; };
// the same three arrays sized by the CTA's own thread count, carved out of dynamic shared memory: the hash kernels run with
// 32..1024 threads, and 16 KB of static arrays per 32-thread CTA held their occupancy at 15 % (R-MAT scale 20, ncu)
struct EnumPtrs {
    u64 *av; u32 *pre; u32 *start;
    __device__ __forceinline__ void bind(unsigned char *base, u32 nt) { av = reinterpret_cast<u64 *>(base); pre = reinterpret_cast<u32 *>(av + nt); start = pre + nt + 1; }
    static __host__ __device__ size_t bytes(u32 nt) { return (size_t)nt * 16 + 8; }
};
template <bool NEEDED> struct EnumStore { EnumSmem s; };                   // kernels that enumerate only in some variants
template <> struct EnumStore<false> { u32 s; };
template <typename VT, bool NUMERIC, typename ES, typename F>
__device__ __forceinline__ u32 enumerate_products(const u32 *__restrict__ Ac, const VT *__restrict__ Av, u32 lenA,
                                                  const uint2 *__restrict__ bdesc, ES &es, u32 *s_warp, F f) {
    const u32 nt = blockDim.x, tid = threadIdx.x;
    u32 done = 0;
    for (u32 base = 0; base < lenA; base += nt) {
        const u32 t = base + tid;
        u32 len = 0, start = 0; u64 av = 0;
        if (t < lenA) { const uint2 d = bdesc[Ac[t]]; start = d.x; len = d.y; if (NUMERIC) av = (u64)Av[t]; }
        u32 total;
        const u32 ex = block_excl_scan(len, s_warp, total);
        es.pre[tid] = ex; es.start[tid] = start;
        if (NUMERIC) es.av[tid] = av;
        __syncthreads();
        for (u32 q = tid; q < total; q += nt) {
            u32 lo = 0, hi = nt;                                            // last entry e with pre[e] <= q (its B row is not empty)
            while (hi - lo > 1) { const u32 mid = (lo + hi) >> 1; if (es.pre[mid] <= q) lo = mid; else hi = mid; }
            f(done + q, es.start[lo] + (q - es.pre[lo]), NUMERIC ? (VT)es.av[lo] : (VT)0);
        }
        done += total;
        __syncthreads();
    }
    return done;
}

__global__ void __launch_bounds__(256) k_build_pack(u64 rows, const u64 *__restrict__ rp, const u32 *__restrict__ col, uint4 *__restrict__ pack) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += (u64)gridDim.x * blockDim.x) {
        const u64 s = rp[i];
        const u32 len = (u32)(rp[i + 1] - s);
        u32 c[B200_PACK_INLINE];
#pragma unroll
        for (int j = 0; j < B200_PACK_INLINE; j++) c[j] = len ? col[s + (j < (int)len ? j : (int)len - 1)] : 0u;   // unused slots repeat the last column (rowwarp.cu runs them unpredicated)
        pack[2 * i] = make_uint4((u32)s, len, c[0], c[1]);
        pack[2 * i + 1] = make_uint4(c[2], c[3], c[4], c[5]);
    }
}

// pipeline state of one warp's row stream (see tiny_gather)
struct TinyRow { u32 row, dA, k; u64 base; };

__global__ void __launch_bounds__(256) k_sym_tiny(SymArgs a, const u32 *__restrict__ bin_rows, B200Ctrl *ctrl, u32 *__restrict__ nnz_row) {
    const u32 count = ctrl->sym_bin_count[B200_BIN_TINY];
    const int lane = threadIdx.x & 31;
    const u32 wpb = blockDim.x >> 5, stride = gridDim.x * wpb;
    u32 r = blockIdx.x * wpb + (threadIdx.x >> 5);
    if (r >= count) return;
    TinyRow cur, nxt;
    cur.row = bin_rows[r];
    { const u64 s = a.rpA[cur.row]; cur.dA = (u32)(a.rpA[cur.row + 1] - s); cur.k = lane < (int)cur.dA ? a.colA[s + lane] : 0u; }
    nxt.row = r + stride < count ? bin_rows[r + stride] : 0u;
    for (; r < count; r += stride) {
        const u32 row_nn = r + 2 * stride < count ? bin_rows[r + 2 * stride] : 0u;
        u32 key; u32 val;
        tiny_gather<u32, false>(cur.dA, cur.k, 0u, a.bdesc, a.colB, nullptr, lane, key, val);
        u64 s_n = 0; nxt.dA = 0;
        if (r + stride < count) { s_n = a.rpA[nxt.row]; nxt.dA = (u32)(a.rpA[nxt.row + 1] - s_n); }
        warp_bitonic<u32, false>(key, val, lane);
        const u32 prev = __shfl_up_sync(0xFFFFFFFFu, key, 1);
        const bool head = key != B200_EMPTY_KEY && (lane == 0 || prev != key);
        const u32 n = __popc(__ballot_sync(0xFFFFFFFFu, head));
        if (lane == 0) nnz_row[cur.row] = n;
        nxt.k = lane < (int)nxt.dA ? a.colA[s_n + lane] : 0u;
        cur = nxt; nxt.row = row_nn;
    }
}

template <typename VT>
__global__ void __launch_bounds__(256) k_num_tiny(NumArgs<VT> a, const u32 *__restrict__ bin_rows, B200Ctrl *ctrl, OutArgs<VT> o,
                                                  bool bpat, bool narrow, bool v32) {
    const u32 count = o.bin_cnt[B200_BIN_TINY];
    const int lane = threadIdx.x & 31;
    const u32 wpb = blockDim.x >> 5, stride = gridDim.x * wpb;
    u32 r = blockIdx.x * wpb + (threadIdx.x >> 5);
    if (r >= count) return;
    u64 vmax = 0;
    TinyRow cur, nxt;
    VT av = 0, av_n = 0;
    cur.row = bin_rows[r];
    {
        const u64 s = a.rpA[cur.row];
        cur.dA = (u32)(a.rpA[cur.row + 1] - s); cur.base = o.base[cur.row];
        cur.k = 0;
        if (lane < (int)cur.dA) { cur.k = a.colA[s + lane]; av = a.valA[s + lane]; }
    }
    nxt.row = r + stride < count ? bin_rows[r + stride] : 0u;
    for (; r < count; r += stride) {
        const u32 row_nn = r + 2 * stride < count ? bin_rows[r + 2 * stride] : 0u;
        u32 key; VT val;
        tiny_gather<VT, true>(cur.dA, cur.k, av, a.bdesc, a.colB, a.valB, lane, key, val, bpat);
        // next row's row_ptr and output base (its id arrived an iteration ago)
        u64 s_n = 0; nxt.dA = 0; nxt.base = 0;
        if (r + stride < count) { s_n = a.rpA[nxt.row]; nxt.dA = (u32)(a.rpA[nxt.row + 1] - s_n); nxt.base = o.base[nxt.row]; }
        // The kernel is shuffle-bound (sort + segmented scan), so shuffle as little as possible.  Column indices below 2^27
        // are sorted packed with their source lane in one 32-bit word (one SHFL per stage instead of one per key plus one
        // or two per value) and the value is fetched from its source lane afterwards; when the host proved that row sums
        // stay below 2^32 (accumulator mode 0) the segmented scan also runs on 32-bit values.
        if (narrow) {
            u32 packed = key == B200_EMPTY_KEY ? B200_EMPTY_KEY : (key << 5) | (u32)lane, none = 0;
            warp_bitonic<u32, false>(packed, none, lane);
            key = packed == B200_EMPTY_KEY ? B200_EMPTY_KEY : packed >> 5;
            val = shfl_any(val, (int)(packed & 31u));
        } else {
            warp_bitonic<VT, true>(key, val, lane);
        }
        // segmented inclusive scan (saturating) over runs of equal keys; the run's last lane has the total
        if (v32) {
            u32 v = (u32)val;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const u32 pk = __shfl_up_sync(0xFFFFFFFFu, key, d);
                const u32 pv = __shfl_up_sync(0xFFFFFFFFu, v, d);
                if (lane >= d && pk == key) v += pv;
            }
            val = (VT)v;
        } else {
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const u32 pk = __shfl_up_sync(0xFFFFFFFFu, key, d);
                const VT pv = shfl_up_any(val, d);
                if (lane >= d && pk == key) val = sat_add(val, pv);
            }
        }
        const u32 nk = __shfl_down_sync(0xFFFFFFFFu, key, 1);
        const bool tail = key != B200_EMPTY_KEY && (lane == 31 || nk != key);
        const u32 tails = __ballot_sync(0xFFFFFFFFu, tail);
        if (tail) {
            const u64 pos = cur.base + __popc(tails & ((1u << lane) - 1u));
            o.col[pos] = key; put_val(o, pos, val);
            vmax = vmax > (u64)val ? vmax : (u64)val;
        }
        if (lane == 0 && o.nnz_out) o.nnz_out[cur.row] = __popc(tails);
        // next row's A entries (its row_ptr was fetched before the sort)
        nxt.k = 0; av_n = 0;
        if (lane < (int)nxt.dA) { nxt.k = a.colA[s_n + lane]; av_n = a.valA[s_n + lane]; }
        cur = nxt; av = av_n; nxt.row = row_nn;
    }
    vmax = warp_max_u64(vmax);
    if (lane == 0 && vmax) atomicMax(&ctrl->max_val_out, (ull)vmax);
}

// =======================================================================================
// 4. symbolic: distinct output columns per row
// =======================================================================================
// 4a. warp per row (8 rows in flight per CTA), 256-slot key table per warp: rows with P <= 128
#define B200_WARP_SLOTS 256
#define B200_WARP_ORDER_BYTES 528   // (B200_WARP_SLOTS / 2 + 1) bucket counters of the ordering step, rounded to 16 bytes
__global__ void __launch_bounds__(256) k_sym_warp(SymArgs a, const u32 *__restrict__ bin_rows, B200Ctrl *ctrl, int first_bin,
                                                  int nbins, int lg, u32 *__restrict__ nnz_row, u32 bin_stride) {
    __shared__ u32 s_keys[8][B200_WARP_SLOTS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    u32 *keys = s_keys[wid];
    u32 count = 0;
    for (int b = 0; b < nbins; b++) count += ctrl->sym_bin_count[first_bin + b];
    const u32 G = 1u << lg, sub = lane & (G - 1), grp = lane >> lg, ngrp = 32u >> lg;
    const int shift = 32 - 8;
    for (u32 r = blockIdx.x * 8 + wid; r < count; r += gridDim.x * 8) {
        const u32 row = bin_row_at(bin_rows, ctrl->sym_bin_count, bin_stride, first_bin, nbins, r);
#pragma unroll
        for (int i = 0; i < B200_WARP_SLOTS / 32; i++) keys[i * 32 + lane] = B200_EMPTY_KEY;
        __syncwarp();
        const u64 s = a.rpA[row];
        const u32 lenA = (u32)(a.rpA[row + 1] - s);
        u32 local = 0;
        walk_products<int>(a.colA + s, lenA, a.bdesc, grp, ngrp, sub, G, [](u32) { return 0; },
                           [&](int, u32 jb) {
                               bool fresh;
                               table_insert(keys, B200_WARP_SLOTS - 1, shift, a.colB[jb], fresh);
                               local += fresh;
                           });
        local = warp_sum_u32(local);
        if (lane == 0) nnz_row[row] = local;
        __syncwarp();
    }
}

// 4b. CTA per row: key table, or a column bitmap when the whole column space fits shared memory
template <bool BITMAP>
__global__ void __launch_bounds__(1024) k_sym_cta(SymArgs a, const u32 *__restrict__ bin_rows, B200Ctrl *ctrl, int bin, u32 slots,
                                                  u32 nwords, int lg, u32 *__restrict__ nnz_row, u32 bin_stride, HvSkip skip) {
    extern __shared__ __align__(16) u32 smem[];
    __shared__ u32 s_count, s_warp[33];
    const u32 count = ctrl->sym_bin_count[bin];
    const u64 off = (u64)bin * bin_stride;
    const u32 nt = blockDim.x, tid = threadIdx.x;
    const u32 tabn = BITMAP ? nwords : slots;
    EnumPtrs s_enum; s_enum.bind(reinterpret_cast<unsigned char *>(smem + ((tabn + 3u) & ~3u)), nt);   // behind the table, 16-byte aligned
    const int shift = 32 - (31 - __clz(slots));
    for (u32 r = blockIdx.x; r < count; r += gridDim.x) {
        const u32 row = bin_rows[off + r];
        if (hv_skipped(skip, r, row)) continue;
        for (u32 t = tid; t < tabn; t += nt) smem[t] = BITMAP ? 0u : B200_EMPTY_KEY;
        if (tid == 0) s_count = 0;
        __syncthreads();
        const u64 s = a.rpA[row];
        const u32 lenA = (u32)(a.rpA[row + 1] - s);
        u32 local = 0;
        enumerate_products<u32, false>(a.colA + s, (const u32 *)nullptr, lenA, a.bdesc, s_enum, s_warp,
                           [&](u32, u32 jb, u32) {
                               const u32 c = a.colB[jb];
                               if (BITMAP) {
                                   const u32 bit = 1u << (c & 31);
                                   const u32 old = atomicOr(&smem[c >> 5], bit);
                                   local += !(old & bit);
                               } else {
                                   bool fresh;
                                   table_insert(smem, slots - 1, shift, c, fresh);
                                   local += fresh;
                               }
                           });
        local = warp_sum_u32(local);
        if ((tid & 31) == 0 && local) atomicAdd(&s_count, local);
        __syncthreads();
        if (tid == 0) nnz_row[row] = s_count;
        __syncthreads();
    }
}

// =======================================================================================
// 5. numeric
// =======================================================================================
// rows [begin,end) of a bin handled by this CTA: contiguous chunks keep consecutive rows (which share
// most of their B rows in lattice-like graphs) on one SM, so the records stay hot in its L1
__device__ __forceinline__ void cta_row_range(u32 count, u32 &begin, u32 &end) {
    const u32 rpc = (count + gridDim.x - 1) / gridDim.x;
    begin = blockIdx.x * rpc;
    end = begin + rpc < count ? begin + rpc : count;
}

// Bitonic sort of n (power of two) packed 64-bit words (column << 32 | table slot) by the whole CTA.  The payload travels
// inside the word, so a compare-exchange is two 64-bit loads and at most two stores (the accumulators stay where the hash
// put them).  Pair t of a stage always belongs to 64-element block t / 32 and the loop below gives warp w the pairs
// 32w .. 32w+31 (+ multiples of blockDim), so all stages with k <= 64 are warp-local: __syncwarp instead of a CTA barrier.
// ordering step of the hash kernels: buckets per row (a power of two, about the row's capacity) and the bucket population
// beyond which the row falls back to the bitonic network
#define B200_ORDER_MAXB 96u
__host__ __device__ __forceinline__ u32 b200_order_buckets(u32 slots) { const u32 nb = slots / 2; return nb < 2048u ? nb : 2048u; }
__device__ __forceinline__ void smem_bitonic_words(u64 *w, u32 n) {
    const u32 tid = threadIdx.x, nt = blockDim.x;
    for (u32 k = 2; k <= n; k <<= 1) {
        for (u32 j = k >> 1; j > 0; j >>= 1) {
            for (u32 t = tid; t < (n >> 1); t += nt) {
                const u32 i = 2 * t - (t & (j - 1));
                const u32 l = i + j;
                const bool up = ((i & k) == 0);
                const u64 x = w[i], y = w[l];
                if ((x > y) == up) { w[i] = y; w[l] = x; }
            }
            if (k <= 64) __syncwarp(); else __syncthreads();
        }
        if (k == 64) __syncthreads();                                       // the next stage crosses the warps' blocks
    }
    __syncthreads();
}

// 5a. warp per row (8 rows in flight per CTA): hash accumulate, compact, sort, stream out. nnz <= 128.
template <typename VT, int MODE>
__global__ void __launch_bounds__(256) k_num_warp(NumArgs<VT> a, const u32 *__restrict__ bin_rows, B200Ctrl *ctrl, int first_bin,
                                                  int nbins, int lg, OutArgs<VT> o) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const size_t per_warp = Acc<MODE>::bytes(B200_WARP_SLOTS) + (size_t)B200_WARP_SLOTS * 4 + B200_WARP_ORDER_BYTES;
    unsigned char *base = smem_raw + wid * per_warp;
    Acc<MODE> acc; acc.bind(base, B200_WARP_SLOTS);
    u32 *keys = reinterpret_cast<u32 *>(base + Acc<MODE>::bytes(B200_WARP_SLOTS));
    u32 *bcnt = keys + B200_WARP_SLOTS;                                    // bucket counters / offsets of the ordering step
    constexpr u32 NB = B200_WARP_SLOTS / 2;
    u32 count = 0;
    for (int b = 0; b < nbins; b++) count += o.bin_cnt[first_bin + b];
    const u32 G = 1u << lg, sub = lane & (G - 1), grp = lane >> lg, ngrp = 32u >> lg;
    const int shift = 32 - 8;
    constexpr int PER = B200_WARP_SLOTS / 32;
    u64 vmax = 0;
    for (u32 r = blockIdx.x * 8 + wid; r < count; r += gridDim.x * 8) {
        const u32 row = bin_row_at(bin_rows, o.bin_cnt, o.bin_stride, first_bin, nbins, r);
#pragma unroll
        for (int i = 0; i < PER; i++) { keys[i * 32 + lane] = B200_EMPTY_KEY; acc.clear(i * 32 + lane); }
        __syncwarp();
        const u64 s = a.rpA[row];
        const u32 lenA = (u32)(a.rpA[row + 1] - s);
        const VT *Av = a.valA + s;
        walk_products<VT>(a.colA + s, lenA, a.bdesc, grp, ngrp, sub, G, [&](u32 t) { return Av[t]; },
                          [&](VT av, u32 jb) {
                              bool fresh;
                              const u32 h = table_insert(keys, B200_WARP_SLOTS - 1, shift, a.colB[jb], fresh);
                              acc.add(h, av, a.valB[jb]);
                          });
        __syncwarp();
        // order the row (see k_num_cta): bucket ranking over NB = 128 buckets of the row's column range, by the warp alone;
        // the bitonic network only for rows whose columns crowd into one bucket
        u32 rk[PER]; u32 mine = 0, cmin = 0xFFFFFFFFu, cmax = 0;
#pragma unroll
        for (int i = 0; i < PER; i++) {
            rk[i] = keys[i * 32 + lane];
            if (rk[i] != B200_EMPTY_KEY) { mine++; cmin = min(cmin, rk[i]); cmax = max(cmax, rk[i]); }
        }
        for (u32 t = lane; t <= NB; t += 32) bcnt[t] = 0;
        u32 incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { u32 t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
        const u32 total = __shfl_sync(0xFFFFFFFFu, incl, 31);
        u32 pos = incl - mine;
        cmin = __reduce_min_sync(0xFFFFFFFFu, cmin); cmax = __reduce_max_sync(0xFFFFFFFFu, cmax);
        __syncwarp();
        u64 *words = reinterpret_cast<u64 *>(keys);
        const u64 obase = o.base[row];
        int bshift = 0;
        if (total) { const u32 span1 = cmax - cmin; const int bits = span1 ? 32 - __clz(span1) : 0; bshift = bits > 7 ? bits - 7 : 0; }
        static_assert(NB == 128, "bshift assumes 128 buckets");
        u32 bp[PER];
        bool crowded = false;
#pragma unroll
        for (int i = 0; i < PER; i++)
            if (rk[i] != B200_EMPTY_KEY) { bp[i] = atomicAdd(&bcnt[(rk[i] - cmin) >> bshift], 1u); crowded |= bp[i] >= 24u; }
        __syncwarp();
        if (__any_sync(0xFFFFFFFFu, crowded)) {
#pragma unroll
            for (int i = 0; i < PER; i++) if (rk[i] != B200_EMPTY_KEY) words[pos++] = ((u64)rk[i] << 32) | (u64)(i * 32 + lane);
            u32 n2 = 1; while (n2 < total) n2 <<= 1;
            for (u32 t = total + lane; t < n2; t += 32) words[t] = ~0ull;
            __syncwarp();
            for (u32 k = 2; k <= n2; k <<= 1) {
                for (u32 j = k >> 1; j > 0; j >>= 1) {
                    for (u32 t = lane; t < (n2 >> 1); t += 32) {
                        const u32 i = 2 * t - (t & (j - 1));
                        const u32 l = i + j;
                        const u64 x = words[i], y = words[l];
                        if ((x > y) == ((i & k) == 0)) { words[i] = y; words[l] = x; }
                    }
                    __syncwarp();
                }
            }
            for (u32 t = lane; t < total; t += 32) {
                const u64 wd = words[t];
                const VT v = emit_val<VT>(acc.get((u32)wd));
                o.col[obase + t] = (u32)(wd >> 32); put_val(o, obase + t, v);
                vmax = vmax > (u64)v ? vmax : (u64)v;
            }
        } else {
            {   // counters -> offsets: lane l owns counters 4l .. 4l + 3
                u32 c[4], sum = 0;
#pragma unroll
                for (int i = 0; i < 4; i++) { c[i] = bcnt[lane * 4 + i]; sum += c[i]; }
                u32 inc2 = sum;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { u32 t = __shfl_up_sync(0xFFFFFFFFu, inc2, d); if (lane >= d) inc2 += t; }
                u32 run = inc2 - sum;
#pragma unroll
                for (int i = 0; i < 4; i++) { bcnt[lane * 4 + i] = run; run += c[i]; }
                if (lane == 0) bcnt[NB] = total;
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < PER; i++)
                if (rk[i] != B200_EMPTY_KEY) words[bcnt[(rk[i] - cmin) >> bshift] + bp[i]] = ((u64)rk[i] << 32) | (u64)(i * 32 + lane);
            __syncwarp();
            for (u32 t = lane; t < total; t += 32) {
                const u64 wd = words[t];
                const u32 c = (u32)(wd >> 32), b = (c - cmin) >> bshift;
                const u32 lo = bcnt[b], hi = bcnt[b + 1];
                u32 rank = 0;
                for (u32 j = lo; j < hi; j++) rank += (u32)(words[j] >> 32) < c ? 1u : 0u;
                const VT v = emit_val<VT>(acc.get((u32)wd));
                o.col[obase + lo + rank] = c; put_val(o, obase + lo + rank, v);
                vmax = vmax > (u64)v ? vmax : (u64)v;
            }
        }
        if (lane == 0 && o.nnz_out) o.nnz_out[row] = total;
        __syncwarp();
    }
    vmax = warp_max_u64(vmax);
    if (lane == 0 && vmax) atomicMax(&ctrl->max_val_out, (ull)vmax);
}

// 5b. CTA per row, hash + sort emission (any column space).  slots = 2 * bin capacity.
template <typename VT, int MODE>
__global__ void __launch_bounds__(1024) k_num_cta(NumArgs<VT> a, const u32 *__restrict__ bin_rows, B200Ctrl *ctrl, int bin, u32 slots,
                                                  int lg, OutArgs<VT> o) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ u32 s_warp[33];
    __shared__ u32 s_mm[2];
    Acc<MODE> acc; acc.bind(smem_raw, slots);
    u32 *keys = reinterpret_cast<u32 *>(smem_raw + Acc<MODE>::bytes(slots));
    u32 *bcnt = keys + slots;                                               // bucket counters / offsets of the ordering step: NB + 1 words
    const u32 NB = b200_order_buckets(slots);
    EnumPtrs s_enum; s_enum.bind(reinterpret_cast<unsigned char *>(bcnt + ((NB + 1 + 3u) & ~3u)), blockDim.x);   // behind the counters, 16-byte aligned
    const u32 count = o.bin_cnt[bin];
    const u64 off = (u64)bin * o.bin_stride;
    const u32 nt = blockDim.x, tid = threadIdx.x;
    const int shift = 32 - (31 - __clz(slots));
    const u32 per = slots / nt;                                             // <= 16 by the bin table
    u64 vmax = 0;
    for (u32 r = blockIdx.x; r < count; r += gridDim.x) {
        const u32 row = bin_rows[off + r];
        for (u32 t = tid; t < slots; t += nt) { keys[t] = B200_EMPTY_KEY; acc.clear(t); }
        __syncthreads();
        const u64 s = a.rpA[row];
        const u32 lenA = (u32)(a.rpA[row + 1] - s);
        const VT *Av = a.valA + s;
        enumerate_products<VT, true>(a.colA + s, Av, lenA, a.bdesc, s_enum, s_warp,
                          [&](u32, u32 jb, VT av) {
                              bool fresh;
                              const u32 h = table_insert(keys, slots - 1, shift, a.colB[jb], fresh);
                              acc.add(h, av, a.valB[jb]);
                          });
        __syncthreads();
        // Order the row.  The stored columns go through registers; the accumulators stay in their hash slots and are read
        // through the low half of a packed (column << 32 | slot) word on the way out.  Bucket ranking instead of a sort: the
        // row's column range is cut into NB equal buckets (a shift, no division), a counter per bucket gives every column its
        // bucket and a place in it, a scan of the counters the buckets' offsets, and a column's final place is its bucket's
        // offset plus the number of smaller columns IN ITS BUCKET (a handful: NB is about the row's length).  ~40
        // instructions per entry where the bitonic network spends log^2(n) / 2 compare-exchange stages.  Rows whose columns
        // crowd into one bucket (> B200_ORDER_MAXB there) take the bitonic network instead.
        u32 rk[16]; u32 mine = 0, cmin = 0xFFFFFFFFu, cmax = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) {
            if (i < (int)per) {
                rk[i] = keys[i * nt + tid];
                if (rk[i] != B200_EMPTY_KEY) { mine++; cmin = min(cmin, rk[i]); cmax = max(cmax, rk[i]); }
            }
        }
        if (tid == 0) { s_mm[0] = 0xFFFFFFFFu; s_mm[1] = 0; }
        for (u32 t = tid; t <= NB; t += nt) bcnt[t] = 0;
        u32 total;
        u32 pos = block_excl_scan(mine, s_warp, total);                  // its barriers order the key reads above before the words below
        cmin = __reduce_min_sync(0xFFFFFFFFu, cmin); cmax = __reduce_max_sync(0xFFFFFFFFu, cmax);
        if ((tid & 31) == 0 && cmin <= cmax) { atomicMin(&s_mm[0], cmin); atomicMax(&s_mm[1], cmax); }
        __syncthreads();
        cmin = s_mm[0]; cmax = s_mm[1];
        u64 *words = reinterpret_cast<u64 *>(keys);                      // total <= cap = slots / 2 words fit the key array exactly
        const u64 obase = o.base[row];
        // bucket of a column: (c - cmin) >> bshift, at most NB of them
        int bshift = 0;
        if (total) { const u32 span1 = cmax - cmin; const int bits = span1 ? 32 - __clz(span1) : 0; const int lb = 31 - __clz(NB); bshift = bits > lb ? bits - lb : 0; }
        unsigned short bp[16];
        bool crowded = false;
#pragma unroll
        for (int i = 0; i < 16; i++)
            if (i < (int)per && rk[i] != B200_EMPTY_KEY) {
                const u32 p = atomicAdd(&bcnt[(rk[i] - cmin) >> bshift], 1u);
                bp[i] = (unsigned short)p;
                crowded |= p >= B200_ORDER_MAXB;
            }
        if (__syncthreads_or(crowded)) {
            // ---- fallback: gather the words at the front of the key array and sort them
#pragma unroll
            for (int i = 0; i < 16; i++)
                if (i < (int)per && rk[i] != B200_EMPTY_KEY) words[pos++] = ((u64)rk[i] << 32) | (u64)(i * nt + tid);
            u32 n2 = 1; while (n2 < total) n2 <<= 1;
            for (u32 t = total + tid; t < n2; t += nt) words[t] = ~0ull;
            __syncthreads();
            smem_bitonic_words(words, n2);
            for (u32 t = tid; t < total; t += nt) {
                const u64 wd = words[t];
                const VT v = emit_val<VT>(acc.get((u32)wd));
                o.col[obase + t] = (u32)(wd >> 32); put_val(o, obase + t, v);
                vmax = vmax > (u64)v ? vmax : (u64)v;
            }
        } else {
            // ---- counters -> offsets (thread t scans NB / nt consecutive counters, a block scan joins them); bcnt[NB] = total
            {
                const u32 cpt = (NB + nt - 1) / nt, c0 = tid * cpt;
                u32 sum = 0;
                for (u32 i = 0; i < cpt && c0 + i < NB; i++) sum += bcnt[c0 + i];
                u32 tt;
                u32 run = block_excl_scan(sum, s_warp, tt);
                for (u32 i = 0; i < cpt && c0 + i < NB; i++) { const u32 c = bcnt[c0 + i]; bcnt[c0 + i] = run; run += c; }
                if (tid == 0) bcnt[NB] = total;
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 16; i++)
                if (i < (int)per && rk[i] != B200_EMPTY_KEY) words[bcnt[(rk[i] - cmin) >> bshift] + bp[i]] = ((u64)rk[i] << 32) | (u64)(i * nt + tid);
            __syncthreads();
            for (u32 t = tid; t < total; t += nt) {
                const u64 wd = words[t];
                const u32 c = (u32)(wd >> 32), b = (c - cmin) >> bshift;
                const u32 lo = bcnt[b], hi = bcnt[b + 1];
                u32 rank = 0;
                for (u32 j = lo; j < hi; j++) rank += (u32)(words[j] >> 32) < c ? 1u : 0u;
                const VT v = emit_val<VT>(acc.get((u32)wd));
                o.col[obase + lo + rank] = c; put_val(o, obase + lo + rank, v);
                vmax = vmax > (u64)v ? vmax : (u64)v;
            }
        }
        if (tid == 0 && o.nnz_out) o.nnz_out[row] = total;
        __syncthreads();
    }
    vmax = warp_max_u64(vmax);
    if ((tid & 31) == 0 && vmax) atomicMax(&ctrl->max_val_out, (ull)vmax);
}

// 5c. CTA per row, rank emission: the whole column space fits a shared-memory bitmap.
// No hash table: walk 1 sets one bit per product column; a prefix popcount over the bitmap words
// gives every present column its rank (= its position in the sorted output row); walk 2 adds each
// product into acc[rank] and drops the column into cols[rank]; both arrays are then streamed out
// with coalesced stores.  Shared memory: nwords*6 + cap*(8|12) bytes.
// rank phase shared by both rank kernels: fills wpre[] (exclusive prefix popcount per word), returns nnz
__device__ __forceinline__ u32 rank_prefix(const u32 *bm, unsigned short *wpre, u32 nwords, u32 *s_warp) {
    const u32 nt = blockDim.x, tid = threadIdx.x;
    const u32 wpt = (nwords + nt - 1) / nt;                               // consecutive bitmap words per thread
    const u32 w0 = tid * wpt;
    u32 mine = 0;
    for (u32 i = 0; i < wpt; i++) if (w0 + i < nwords) mine += __popc(bm[w0 + i]);
    u32 total;
    u32 run = block_excl_scan(mine, s_warp, total);
    for (u32 i = 0; i < wpt; i++) {
        if (w0 + i < nwords) { wpre[w0 + i] = (unsigned short)run; run += __popc(bm[w0 + i]); }
    }
    return total;
}

template <typename VT, int MODE>
__global__ void __launch_bounds__(1024) k_num_rank(NumArgs<VT> a, const u32 *__restrict__ bin_rows, B200Ctrl *ctrl, int bin, u32 cap,
                                                   u32 nwords, int lg, OutArgs<VT> o) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ u32 s_warp[33];
    Acc<MODE> acc; acc.bind(smem_raw, cap);
    u32 *cols = reinterpret_cast<u32 *>(smem_raw + Acc<MODE>::bytes(cap));
    u32 *bm = cols + cap;
    unsigned short *wpre = reinterpret_cast<unsigned short *>(bm + nwords);
    const u32 count = o.bin_cnt[bin];
    const u64 off = (u64)bin * o.bin_stride;
    const u32 nt = blockDim.x, tid = threadIdx.x;
    const u32 G = 1u << lg, sub = tid & (G - 1), grp = tid >> lg, ngrp = nt >> lg;
    u64 vmax = 0;
    u32 r_begin, r_end;
    cta_row_range(count, r_begin, r_end);
    for (u32 r = r_begin; r < r_end; r++) {
        const u32 row = bin_rows[off + r];
        for (u32 t = tid; t < nwords; t += nt) bm[t] = 0;
        __syncthreads();
        const u64 s = a.rpA[row];
        const u32 lenA = (u32)(a.rpA[row + 1] - s);
        const u32 *Ac = a.colA + s;
        const VT *Av = a.valA + s;
        // ---- walk 1: column bitmap
        walk_products<int>(Ac, lenA, a.bdesc, grp, ngrp, sub, G, [](u32) { return 0; },
                           [&](int, u32 jb) { const u32 c = a.colB[jb]; atomicOr(&bm[c >> 5], 1u << (c & 31)); });
        __syncthreads();
        const u32 nnz = rank_prefix(bm, wpre, nwords, s_warp);
        for (u32 t = tid; t < nnz; t += nt) acc.clear(t);
        __syncthreads();
        // ---- walk 2: accumulate into acc[rank(col)]
        walk_products<VT>(Ac, lenA, a.bdesc, grp, ngrp, sub, G, [&](u32 t) { return Av[t]; },
                          [&](VT av, u32 jb) {
                              const u32 c = a.colB[jb];
                              const u32 pos = (u32)wpre[c >> 5] + __popc(bm[c >> 5] & ((1u << (c & 31)) - 1u));
                              cols[pos] = c;                               // same value from every product of this column
                              acc.add(pos, av, a.valB[jb]);
                          });
        __syncthreads();
        const u64 obase = o.base[row];
        for (u32 t = tid; t < nnz; t += nt) {
            const VT v = emit_val<VT>(acc.get(t));
            o.col[obase + t] = cols[t]; put_val(o, obase + t, v);
            vmax = vmax > (u64)v ? vmax : (u64)v;
        }
        if (tid == 0 && o.nnz_out) o.nnz_out[row] = nnz;
        __syncthreads();
    }
    vmax = warp_max_u64(vmax);
    if ((tid & 31) == 0 && vmax) atomicMax(&ctrl->max_val_out, (ull)vmax);
}

// =======================================================================================
// 5e. expand-then-rank (one-pass numeric, medium rows): CTA per row, every phase after the expansion
//     runs one PRODUCT (or one output entry) per thread, so no lane idles on a short B row.
//       expand : thread per A entry; warp scan of the B-row lengths hands every entry a slice of the
//                row's product buffer in shared memory; the entry's columns (from its sector-packed
//                record) and its product values a_ik*b_kj are written there, and each column is
//                marked in the row's column bitmap (atomicOr) while it is still in a register
//       rank   : prefix popcount over the bitmap, 128-bit shared loads, four words per thread-step
//       accum  : thread per product -> acc[rank(col)] += value, cols[rank(col)] = col
//       emit   : thread per output entry, coalesced stores; accumulators and bitmap are cleared on the
//                way out so the next row starts clean
//     Shared memory: pcap*(4+pv) + ncap*(4+acc) + 6*nwords bytes (pcap = bin's product capacity).
// =======================================================================================
template <int MODE, typename VT> struct PVal { typedef u64 type; };
template <typename VT> struct PVal<0, VT> { typedef u32 type; };
template <> struct PVal<1, u32> { typedef u32 type; };

template <int MODE, typename VT>
__device__ __forceinline__ typename PVal<MODE, VT>::type product_value(VT a, VT b) {
    if (MODE == 0) return (typename PVal<MODE, VT>::type)((u32)a * (u32)b);
    if (MODE == 2) return (typename PVal<MODE, VT>::type)sat_mul((u64)a, (u64)b);
    u64 x = (u64)a * (u64)b;
    if (sizeof(VT) == 4) x = x > 0xFFFFFFFFull ? 0xFFFFFFFFull : x;
    return (typename PVal<MODE, VT>::type)x;
}
// exclusive prefix popcount of the bitmap, four words per step (nw4 = number of uint4 groups); returns nnz.
// Prefixes are stored as four u16 per group (a row handled here has < 65536 entries).
__device__ __forceinline__ u32 rank_prefix_v4(const uint4 *bm4, uint2 *wpre4, u32 nw4, u32 *s_warp) {
    const u32 nt = blockDim.x, tid = threadIdx.x;
    if (nw4 <= nt) {                                                        // one group per thread: no loops
        uint4 w = make_uint4(0, 0, 0, 0);
        if (tid < nw4) w = bm4[tid];
        const u32 c0 = __popc(w.x), c1 = __popc(w.y), c2 = __popc(w.z), c3 = __popc(w.w);
        u32 total;
        const u32 p0 = block_excl_scan(c0 + c1 + c2 + c3, s_warp, total);
        const u32 p1 = p0 + c0, p2 = p1 + c1, p3 = p2 + c2;
        if (tid < nw4) wpre4[tid] = make_uint2(p0 | (p1 << 16), p2 | (p3 << 16));
        return total;
    }
    const u32 per = (nw4 + nt - 1) / nt;
    const u32 g0 = tid * per;
    u32 mine = 0;
#pragma unroll 1
    for (u32 i = 0; i < per; i++) {
        if (g0 + i < nw4) { const uint4 w = bm4[g0 + i]; mine += __popc(w.x) + __popc(w.y) + __popc(w.z) + __popc(w.w); }
    }
    u32 total;
    u32 run = block_excl_scan(mine, s_warp, total);
#pragma unroll 1
    for (u32 i = 0; i < per; i++) {
        if (g0 + i < nw4) {
            const uint4 w = bm4[g0 + i];
            const u32 p0 = run, p1 = p0 + __popc(w.x), p2 = p1 + __popc(w.y), p3 = p2 + __popc(w.z);
            wpre4[g0 + i] = make_uint2(p0 | (p1 << 16), p2 | (p3 << 16));
            run = p3 + __popc(w.w);
        }
    }
    return total;
}

// B row of one A entry as the expansion sees it: where it starts, how long it is, and (packed B) its first columns
struct BRowRef { u32 start, len; uint4 ca, cb; };
template <bool PACK>
__device__ __forceinline__ BRowRef load_brow(const uint4 *__restrict__ pack, const uint2 *__restrict__ bdesc, u32 k) {
    BRowRef b;
    if (PACK) { const PackRec r = load_pack(pack, k); b.start = r.a.x; b.len = r.a.y; b.ca = r.a; b.cb = r.b; }
    else { const uint2 d = bdesc[k]; b.start = d.x; b.len = d.y; b.ca = make_uint4(0, 0, 0, 0); b.cb = b.ca; }
    return b;
}

#define B200_EXPAND_PRE 2   // A entries per thread whose loads are issued one row ahead
#define B200_EXPAND_REC_EARLY 1   // 1: next row's B records are fetched right after the rank phase (live across accumulate + emit)
#define B200_EXPAND_ILP 2   // products each thread keeps in flight in the mark / accumulate loops

template <typename VT, int MODE, bool PACK, bool BPAT, bool SPAN>
__global__ void __launch_bounds__(512) k_num_expand(NumArgs<VT> a, const uint4 *__restrict__ pack, const u32 *__restrict__ bin_rows,
                                                    B200Ctrl *ctrl, int bin, int nbins, u32 pcap, u32 ncap, u32 nw4,
                                                    const uint4 *__restrict__ win, u32 ncols, OutArgs<VT> o) {
    typedef typename PVal<MODE, VT>::type PV;
    constexpr bool PAIR = sizeof(PV) == 4;       // products kept as {col, value} pairs: one 64-bit shared access each
    constexpr int E = B200_EXPAND_PRE;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ u32 s_warp[33];
    __shared__ u32 s_P, s_wlo, s_whi;
    __shared__ EnumStore<!PACK> s_enum;                                     // only the non-packed expansion needs it
    u32 count = 0;
    for (int b = 0; b < nbins; b++) count += o.bin_cnt[bin + b];            // consecutive bins share one launch
    u32 r_begin, r_end;
    cta_row_range(count, r_begin, r_end);
    if (r_begin >= r_end) return;
    // layout: bitmap | word prefixes | products ({col,val} pairs, or values then cols) | accumulators | output cols
    uint4 *bm4 = reinterpret_cast<uint4 *>(smem_raw);
    u32 *bm = reinterpret_cast<u32 *>(smem_raw);
    uint2 *wpre4 = reinterpret_cast<uint2 *>(smem_raw + (size_t)nw4 * 16);
    const unsigned short *wpre = reinterpret_cast<const unsigned short *>(wpre4);
    unsigned char *p8 = smem_raw + (size_t)nw4 * 24;
    uint2 *pp = reinterpret_cast<uint2 *>(p8);                               // PAIR
    u64 *pv = reinterpret_cast<u64 *>(p8);                                   // !PAIR
    unsigned char *q8 = p8 + (size_t)pcap * 8;
    Acc<MODE> acc; acc.bind(q8, ncap);
    u32 *pc = reinterpret_cast<u32 *>(q8 + Acc<MODE>::bytes(ncap));          // !PAIR only
    u32 *cols = PAIR ? pc : pc + pcap;
    const u32 nt = blockDim.x, tid = threadIdx.x, lane = tid & 31;
    for (u32 t = tid; t < nw4; t += nt) bm4[t] = make_uint4(0, 0, 0, 0);
    for (u32 t = tid; t < ncap; t += nt) acc.clear(t);
    if (tid == 0) { s_P = 0; s_wlo = 0xFFFFFFFFu; s_whi = 0; }
    // SPAN (bitmaps with more groups than the CTA has threads): the products also track the bitmap words they touch, and
    // the rank, emit and clear phases walk only the groups between the lowest and the highest of them.
    u32 wlo = 0xFFFFFFFFu, whi = 0;
    // The row's column window: bit d of the bitmap is column (org + d) mod ncols, `groups` 128-column groups long.  The
    // pre-pass hands out {first bit in the rotated column space, groups, rotation}; org folds the two offsets into one.
    u32 groups = nw4, org = 0;
    auto dcol = [&](u32 c) -> u32 { return c >= org ? c - org : c - org + ncols; };
    auto origin = [&](const uint4 &wn) -> u32 { const u64 t = (u64)wn.x + wn.z; return (u32)(t >= ncols ? t - ncols : t); };

    auto put = [&](u32 dst, VT av, u32 c, u32 jb) {
        const PV x = BPAT ? (PV)av : product_value<MODE, VT>(av, a.valB[jb]);
        if (PAIR) pp[dst] = make_uint2(c, (u32)x);
        else { pc[dst] = c; pv[dst] = (u64)x; }
        const u32 d = dcol(c);
        const u32 w = d >> 5;
        atomicOr(&bm[w], __funnelshift_l(0u, 1u, d));                        // mark the column while it is in a register
        if (SPAN) { wlo = min(wlo, w); whi = max(whi, w); }
    };
    // products of one warp's 32 A entries -> a slice of the product buffer
    auto expand32 = [&](bool valid, VT av, const BRowRef &b) {
        const u32 len = valid ? b.len : 0u;
        u32 incl = len;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const u32 x = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += x; }
        u32 pbase = 0;
        if (lane == 31) pbase = atomicAdd(&s_P, incl);
        pbase = __shfl_sync(0xFFFFFFFFu, pbase, 31);
        const u32 dst = pbase + incl - len;
        if (PACK) {
            if (len > 0) put(dst, av, b.ca.z, b.start);
            if (len > 1) put(dst + 1, av, b.ca.w, b.start + 1);
            if (len > 2) put(dst + 2, av, b.cb.x, b.start + 2);
            if (len > 3) put(dst + 3, av, b.cb.y, b.start + 3);
            if (len > 4) put(dst + 4, av, b.cb.z, b.start + 4);
            if (len > 5) put(dst + 5, av, b.cb.w, b.start + 5);
            for (u32 j = B200_PACK_INLINE; j < len; j++) put(dst + j, av, a.colB[b.start + j], b.start + j);
        } else {
            for (u32 j = 0; j < len; j++) put(dst + j, av, a.colB[b.start + j], b.start + j);
        }
    };

    // ---- software pipeline over the CTA's rows: row ids two rows ahead, A entries and B row records one row ahead
    u32 row = bin_row_at(bin_rows, o.bin_cnt, o.bin_stride, bin, nbins, r_begin);
    u64 rs = a.rpA[row];
    u32 lenA = (u32)(a.rpA[row + 1] - rs);
    uint4 wn = win[row];
    groups = wn.y; org = origin(wn);
    u32 row_n = 0, lenA_n = 0; u64 rs_n = 0;
    wn = make_uint4(0, 0, 0, 0);
    if (r_begin + 1 < r_end) { row_n = bin_row_at(bin_rows, o.bin_cnt, o.bin_stride, bin, nbins, r_begin + 1); rs_n = a.rpA[row_n]; lenA_n = (u32)(a.rpA[row_n + 1] - rs_n); wn = win[row_n]; }
    VT av[E]; BRowRef br[E];
#pragma unroll
    for (int e = 0; e < E; e++) {
        const u32 t = tid + e * nt;
        av[e] = 0; br[e].start = 0; br[e].len = 0; br[e].ca = make_uint4(0, 0, 0, 0); br[e].cb = br[e].ca;
        if (PACK && t < lenA) { av[e] = a.valA[rs + t]; br[e] = load_brow<PACK>(pack, a.bdesc, a.colA[rs + t]); }
    }
    __syncthreads();
    u64 vmax = 0;
    for (u32 r = r_begin; r < r_end; r++) {
        const bool has_next = r + 1 < r_end;
        const u64 obase = o.base[row];
        if constexpr (PACK) {
            // ---- expand: one A entry per thread, whole warps iterate together
            const u32 lenA_w = (lenA + 31u) & ~31u;
#pragma unroll
            for (int e = 0; e < E; e++) {
                const u32 t = tid + e * nt;
                if (t - lane < lenA_w) expand32(t < lenA, av[e], br[e]);    // warp-uniform condition
            }
            for (u32 t = tid + E * nt; t < lenA_w; t += nt) {
                const bool valid = t < lenA;
                VT x = 0; BRowRef b; b.start = 0; b.len = 0; b.ca = make_uint4(0, 0, 0, 0); b.cb = b.ca;
                if (valid) { x = a.valA[rs + t]; b = load_brow<PACK>(pack, a.bdesc, a.colA[rs + t]); }
                expand32(valid, x, b);
            }
        } else {
            // ---- expand (B rows of any length): balanced enumeration, product p lands in slot p
            const u32 P = enumerate_products<VT, true>(a.colA + rs, a.valA + rs, lenA, a.bdesc, s_enum.s, s_warp,
                                                       [&](u32 p, u32 jb, VT x) { put(p, x, a.colB[jb], jb); });
            if (tid == 0) s_P = P;
        }
        // next row: A entries now, their B rows after the mark phase (the column indices have arrived by then)
        u32 kn[E]; VT avn[E];
        u32 row_nn = 0;
#pragma unroll
        for (int e = 0; e < E; e++) {
            const u32 t = tid + e * nt;
            kn[e] = 0; avn[e] = 0;
            if (PACK && has_next && t < lenA_n) { kn[e] = a.colA[rs_n + t]; avn[e] = a.valA[rs_n + t]; }
        }
        if (r + 2 < r_end) row_nn = bin_row_at(bin_rows, o.bin_cnt, o.bin_stride, bin, nbins, r + 2);
        if (SPAN) {
            wlo = __reduce_min_sync(0xFFFFFFFFu, wlo); whi = __reduce_max_sync(0xFFFFFFFFu, whi);
            if (lane == 0 && wlo <= whi) { atomicMin(&s_wlo, wlo); atomicMax(&s_whi, whi); }
            wlo = 0xFFFFFFFFu; whi = 0;
        }
        __syncthreads();
        const u32 P = s_P;
        u32 g0 = 0, gn = groups;                                             // groups that hold a product
        if (SPAN) { const u32 lo = s_wlo, hi = s_whi; g0 = lo <= hi ? lo >> 2 : 0u; gn = lo <= hi ? (hi >> 2) - g0 + 1 : 0u; }
        const u32 nnz = rank_prefix_v4(bm4 + g0, wpre4 + g0, gn, s_warp);
#if B200_EXPAND_REC_EARLY
#pragma unroll
        for (int e = 0; e < E; e++) {
            const u32 t = tid + e * nt;
            br[e].len = 0;
            if (PACK && has_next && t < lenA_n) br[e] = load_brow<PACK>(pack, a.bdesc, kn[e]);
            av[e] = avn[e];
        }
#endif
        u64 rs_nn = 0; u32 lenA_nn = 0; uint4 wnn = make_uint4(0, 0, 0, 0);
        if (r + 2 < r_end) { rs_nn = a.rpA[row_nn]; lenA_nn = (u32)(a.rpA[row_nn + 1] - rs_nn); wnn = win[row_nn]; }
        __syncthreads();
        // ---- accumulate at the column's rank
        for (u32 p0 = tid; p0 < P; p0 += B200_EXPAND_ILP * nt) {
            u32 c[B200_EXPAND_ILP], pos[B200_EXPAND_ILP]; u64 x[B200_EXPAND_ILP];
#pragma unroll
            for (int j = 0; j < B200_EXPAND_ILP; j++) {
                const u32 p = p0 + j * nt;
                c[j] = B200_EMPTY_KEY; x[j] = 0;
                if (p < P) { if (PAIR) { const uint2 q = pp[p]; c[j] = q.x; x[j] = q.y; } else { c[j] = pc[p]; x[j] = pv[p]; } }
            }
#pragma unroll
            for (int j = 0; j < B200_EXPAND_ILP; j++) {
                const u32 dd = c[j] != B200_EMPTY_KEY ? dcol(c[j]) : 0u;          // padding lanes read word 0
                const u32 w = dd >> 5;
                pos[j] = (u32)wpre[w] + __popc(bm[w] & (__funnelshift_l(0u, 1u, dd) - 1u));
            }
#pragma unroll
            for (int j = 0; j < B200_EXPAND_ILP; j++)
                if (c[j] != B200_EMPTY_KEY) { cols[pos[j]] = c[j]; acc.addv(pos[j], x[j]); }
        }
        __syncthreads();
        // ---- emit (coalesced), leaving accumulators, bitmap and the product counter clean.  Ranks are in d order; when the
        // window starts at a column org > 0 the entries whose column is below org (d >= ncols - org) belong in FRONT of
        // the others: the row is written rotated by r0 = number of entries with d < ncols - org.
        u32 r0 = nnz;
        if (org) {
            const u32 split = ncols - org;                                   // first d that maps to a column below org
            const u32 slo = g0 << 7, shi = slo + (gn << 7);                  // d range whose prefixes are valid (holds every entry)
            if (split <= slo) r0 = 0;
            else if (split < shi) { const u32 sw = split >> 5; r0 = (u32)wpre[sw] + __popc(bm[sw] & (__funnelshift_l(0u, 1u, split) - 1u)); }
        }
        const u32 shift_hi = nnz - r0;                                       // d-rank t -> t - r0 (t >= r0) or t + shift_hi
        for (u32 t0 = tid; t0 < nnz; t0 += 2 * nt) {
            const u32 t1 = t0 + nt;
            const bool h1 = t1 < nnz;
            const u32 c0 = cols[t0], c1 = h1 ? cols[t1] : 0u;
            const VT v0 = emit_val<VT>(acc.get(t0)), v1 = h1 ? emit_val<VT>(acc.get(t1)) : (VT)0;
            const u32 q0 = t0 >= r0 ? t0 - r0 : t0 + shift_hi, q1 = t1 >= r0 ? t1 - r0 : t1 + shift_hi;
            acc.clear(t0);
            o.col[obase + q0] = c0; put_val(o, obase + q0, v0);
            if (h1) { acc.clear(t1); o.col[obase + q1] = c1; put_val(o, obase + q1, v1); }
            const u64 m = (u64)(v0 > v1 ? v0 : v1);
            vmax = vmax > m ? vmax : m;
        }
#if !B200_EXPAND_REC_EARLY
        // next row's B records: issued only now, so that their 16 registers are not live across the accumulate and emit
        // loops (63 -> fewer registers: one more CTA per SM); the clear below and the barrier cover part of their latency
#pragma unroll
        for (int e = 0; e < E; e++) {
            const u32 t = tid + e * nt;
            br[e].len = 0;
            if (PACK && has_next && t < lenA_n) br[e] = load_brow<PACK>(pack, a.bdesc, kn[e]);
            av[e] = avn[e];
        }
#endif
        for (u32 t = tid; t < gn; t += nt) bm4[g0 + t] = make_uint4(0, 0, 0, 0);
        if (tid == 0) { s_P = 0; if (SPAN) { s_wlo = 0xFFFFFFFFu; s_whi = 0; } if (o.nnz_out) o.nnz_out[row] = nnz; }
        __syncthreads();
        row = row_n; rs = rs_n; lenA = lenA_n; groups = wn.y; org = origin(wn);
        row_n = row_nn; rs_n = rs_nn; lenA_n = lenA_nn; wn = wnn;
    }
    vmax = warp_max_u64(vmax);
    if (lane == 0 && vmax) atomicMax(&ctrl->max_val_out, (ull)vmax);
}

// 5f. count-only twin of k_num_expand (exact mode: C is written once, at its final offsets): the same expansion, but a
//     product only marks its column; the row's nnz is the popcount of its window.  Shared memory: the bitmap alone.
template <bool PACK>
__global__ void __launch_bounds__(512) k_sym_expand(SymArgs a, const uint4 *__restrict__ pack, const u32 *__restrict__ bin_rows,
                                                    B200Ctrl *ctrl, int bin, int nbins, u32 nw4, const uint4 *__restrict__ win,
                                                    u32 ncols, u32 *__restrict__ nnz_row, u32 bin_stride) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ u32 s_warp[33];
    __shared__ EnumStore<!PACK> s_enum;
    u32 count = 0;
    for (int b = 0; b < nbins; b++) count += ctrl->sym_bin_count[bin + b];
    u32 r_begin, r_end;
    cta_row_range(count, r_begin, r_end);
    if (r_begin >= r_end) return;
    uint4 *bm4 = reinterpret_cast<uint4 *>(smem_raw);
    u32 *bm = reinterpret_cast<u32 *>(smem_raw);
    const u32 nt = blockDim.x, tid = threadIdx.x;
    for (u32 t = tid; t < nw4; t += nt) bm4[t] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    for (u32 r = r_begin; r < r_end; r++) {
        const u32 row = bin_row_at(bin_rows, ctrl->sym_bin_count, bin_stride, bin, nbins, r);
        const u64 s = a.rpA[row];
        const u32 lenA = (u32)(a.rpA[row + 1] - s);
        const uint4 wn = win[row];
        const u32 wbase = wn.x >> 5, groups = wn.y, rot = wn.z;
        auto mark = [&](u32 c) { const u32 d = c >= rot ? c - rot : c - rot + ncols; atomicOr(&bm[(d >> 5) - wbase], __funnelshift_l(0u, 1u, d)); };
        if constexpr (!PACK) {
            enumerate_products<u32, false>(a.colA + s, (const u32 *)nullptr, lenA, a.bdesc, s_enum.s, s_warp,
                                           [&](u32, u32 jb, u32) { mark(a.colB[jb]); });
        }
        // packed B: two A entries per thread and iteration so that their dependent loads overlap
        for (u32 t = tid; PACK && t < lenA; t += 2 * nt) {
            const u32 t1 = t + nt;
            const bool h1 = t1 < lenA;
            const u32 k0 = a.colA[s + t], k1 = h1 ? a.colA[s + t1] : k0;
            const BRowRef b0 = load_brow<PACK>(pack, a.bdesc, k0);
            BRowRef b1 = load_brow<PACK>(pack, a.bdesc, k1);
            if (!h1) b1.len = 0;
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const BRowRef &b = e ? b1 : b0;
                if (PACK) {
                    if (b.len > 0) mark(b.ca.z);
                    if (b.len > 1) mark(b.ca.w);
                    if (b.len > 2) mark(b.cb.x);
                    if (b.len > 3) mark(b.cb.y);
                    if (b.len > 4) mark(b.cb.z);
                    if (b.len > 5) mark(b.cb.w);
                    for (u32 j = B200_PACK_INLINE; j < b.len; j++) mark(a.colB[b.start + j]);
                } else {
                    for (u32 j = 0; j < b.len; j++) mark(a.colB[b.start + j]);
                }
            }
        }
        __syncthreads();
        u32 mine = 0;
        for (u32 g = tid; g < groups; g += nt) {
            const uint4 w = bm4[g];
            mine += __popc(w.x) + __popc(w.y) + __popc(w.z) + __popc(w.w);
            bm4[g] = make_uint4(0, 0, 0, 0);
        }
        u32 total;
        block_excl_scan(mine, s_warp, total);                              // ends on a barrier: the bitmap is clean for the next row
        if (tid == 0) nnz_row[row] = total;
    }
}

// =======================================================================================
// 6. heavy rows: table, bitmap and rank array in global scratch (one CTA per row at a time)
// =======================================================================================
__global__ void __launch_bounds__(1024) k_sym_heavy(SymArgs a, const u32 *__restrict__ bin_rows, B200Ctrl *ctrl, u32 nwords,
                                                    u32 *__restrict__ scratch_bm, u32 *__restrict__ nnz_row, u32 bin_stride, HvSkip skip) {
    __shared__ u32 s_count, s_warp[33];
    __shared__ EnumSmem s_enum;
    const u32 count = ctrl->sym_bin_count[B200_BIN_HEAVY];
    const u64 off = (u64)B200_BIN_HEAVY * bin_stride;
    u32 *bm = scratch_bm + (u64)blockIdx.x * nwords;
    const u32 nt = blockDim.x, tid = threadIdx.x, lane = tid & 31;
    for (u32 r = blockIdx.x; r < count; r += gridDim.x) {
        const u32 row = bin_rows[off + r];
        if (hv_skipped(skip, r, row)) continue;
        for (u32 t = tid; t < nwords; t += nt) bm[t] = 0;
        if (tid == 0) s_count = 0;
        __syncthreads();
        const u64 s = a.rpA[row];
        const u32 lenA = (u32)(a.rpA[row + 1] - s);
        u32 local = 0;
        enumerate_products<u32, false>(a.colA + s, (const u32 *)nullptr, lenA, a.bdesc, s_enum, s_warp,
                           [&](u32, u32 jb, u32) {
                               const u32 c = a.colB[jb];
                               const u32 bit = 1u << (c & 31);
                               const u32 old = atomicOr(&bm[c >> 5], bit);
                               local += !(old & bit);
                           });
        local = warp_sum_u32(local);
        if (lane == 0 && local) atomicAdd(&s_count, local);
        __syncthreads();
        if (tid == 0) nnz_row[row] = s_count;
        __syncthreads();
    }
}

#define B200_HEAVY_NB_MAX 16384u   // bucket counters of the ordering step, at most
template <typename VT, int MODE>
__global__ void __launch_bounds__(1024) k_num_heavy(NumArgs<VT> a, const u32 *__restrict__ bin_rows, B200Ctrl *ctrl,
                                                    const u32 *__restrict__ nnz_row, u64 max_slots,
                                                    u32 *__restrict__ scratch_place, u32 *__restrict__ scratch_order, u32 *__restrict__ scratch_cnt,
                                                    u32 *__restrict__ scratch_keys, u64 *__restrict__ scratch_vals, OutArgs<VT> o, HvSkip skip) {
    // Heavy rows that the chunked kernels leave alone (wide column space, too few products per chunk): one CTA per row, an
    // open-addressing table in global memory sized from the row's exact length (count pass), then the bucket ranking of the
    // hash kernels in global memory -- the row's column range cut into ~nnz / 8 buckets, a counter per bucket, a scan of
    // the counters, and each column's place = its bucket's offset + the smaller columns in its bucket.  (Round 1 ranked
    // through a bitmap of the whole column space: clearing and prefix-scanning ncols / 32 words per row cost more than the
    // row's products on a 4 M-column R-MAT.)
    __shared__ u32 s_warp[33], s_mm[2];
    __shared__ EnumSmem s_enum;
    const u32 count = o.bin_cnt[B200_BIN_HEAVY];
    const u64 off = (u64)B200_BIN_HEAVY * o.bin_stride;
    u32 *place = scratch_place + (u64)blockIdx.x * max_slots;
    u32 *order = scratch_order + (u64)blockIdx.x * (max_slots / 2 + 1);
    u32 *bcnt = scratch_cnt + (u64)blockIdx.x * (B200_HEAVY_NB_MAX + 1);
    u32 *keys = scratch_keys + (u64)blockIdx.x * max_slots;
    ull *vals = (ull *)(scratch_vals + (u64)blockIdx.x * max_slots);
    const u32 nt = blockDim.x, tid = threadIdx.x, lane = tid & 31;
    u64 vmax = 0;
    for (u32 r = blockIdx.x; r < count; r += gridDim.x) {
        const u32 row = bin_rows[off + r];
        if (hv_skipped(skip, r, row)) continue;
        const u32 nnz = nnz_row[row];
        u64 slots = 1; while (slots < 2ull * nnz) slots <<= 1;            // <= max_slots by host sizing
        const int shift = 64 - (63 - __clzll(slots));
        u32 NB = 256; while (NB < B200_HEAVY_NB_MAX && (u64)NB * 8 < nnz) NB <<= 1;
        for (u64 t = tid; t < slots; t += nt) { keys[t] = B200_EMPTY_KEY; vals[t] = 0; }
        for (u32 t = tid; t <= NB; t += nt) bcnt[t] = 0;
        if (tid == 0) { s_mm[0] = 0xFFFFFFFFu; s_mm[1] = 0; }
        __syncthreads();
        const u64 s = a.rpA[row];
        const u32 lenA = (u32)(a.rpA[row + 1] - s);
        const VT *Av = a.valA + s;
        u32 cmin = 0xFFFFFFFFu, cmax = 0;
        enumerate_products<VT, true>(a.colA + s, Av, lenA, a.bdesc, s_enum, s_warp,
                          [&](u32, u32 jb, VT av) {
                              const u32 c = a.colB[jb];
                              cmin = min(cmin, c); cmax = max(cmax, c);
                              ull x;
                              if (MODE == 2) x = sat_mul((u64)av, (u64)a.valB[jb]);
                              else if (sizeof(VT) == 4) { u64 p = (u64)av * (u64)a.valB[jb]; x = p > 0xFFFFFFFFull ? 0xFFFFFFFFull : p; }
                              else x = (u64)av * (u64)a.valB[jb];
                              u64 h = ((u64)c * 0x9E3779B97F4A7C15ull) >> shift;
                              while (true) {
                                  u32 cur = ld_volatile_u32(&keys[h]);
                                  if (cur == B200_EMPTY_KEY) {
                                      cur = atomicCAS(&keys[h], B200_EMPTY_KEY, c);
                                      if (cur == B200_EMPTY_KEY) cur = c;
                                  }
                                  if (cur == c) {
                                      if (MODE == 2) {
                                          ull old = ld_volatile_u64((const u64 *)&vals[h]), assumed;
                                          do {
                                              assumed = old;
                                              ull sum = assumed + x; if (sum < assumed) sum = ~0ull;
                                              if (sum == assumed) break;
                                              old = atomicCAS(&vals[h], assumed, sum);
                                          } while (old != assumed);
                                      } else atomicAdd(&vals[h], x);
                                      break;
                                  }
                                  h = (h + 1) & (slots - 1);
                              }
                          });
        cmin = __reduce_min_sync(0xFFFFFFFFu, cmin); cmax = __reduce_max_sync(0xFFFFFFFFu, cmax);
        if (lane == 0 && cmin <= cmax) { atomicMin(&s_mm[0], cmin); atomicMax(&s_mm[1], cmax); }
        __threadfence_block();
        __syncthreads();
        cmin = s_mm[0]; cmax = s_mm[1];
        int bshift = 0;
        if (nnz) { const u32 span1 = cmax - cmin; const int bits = span1 ? 32 - __clz(span1) : 0; const int lb = 31 - __clz(NB); bshift = bits > lb ? bits - lb : 0; }
        // ---- a place in its bucket for every stored column
        for (u64 t = tid; t < slots; t += nt) {
            const u32 c = __ldcg(&keys[t]);
            if (c != B200_EMPTY_KEY) place[t] = atomicAdd(&bcnt[(c - cmin) >> bshift], 1u);
        }
        __syncthreads();
        // ---- counters -> offsets
        {
            const u32 cpt = (NB + nt - 1) / nt, c0 = tid * cpt;
            u32 sum = 0;
            for (u32 i = 0; i < cpt && c0 + i < NB; i++) sum += __ldcg(&bcnt[c0 + i]);
            u32 tt;
            u32 run = block_excl_scan(sum, s_warp, tt);
            for (u32 i = 0; i < cpt && c0 + i < NB; i++) { const u32 c = __ldcg(&bcnt[c0 + i]); bcnt[c0 + i] = run; run += c; }
            if (tid == 0) bcnt[NB] = nnz;
        }
        __threadfence_block();
        __syncthreads();
        // ---- slots in bucket order
        for (u64 t = tid; t < slots; t += nt) {
            const u32 c = __ldcg(&keys[t]);
            if (c != B200_EMPTY_KEY) order[__ldcg(&bcnt[(c - cmin) >> bshift]) + place[t]] = (u32)t;
        }
        __threadfence_block();
        __syncthreads();
        // ---- rank inside the bucket, emit
        const u64 obase = o.base[row];
        for (u32 t = tid; t < nnz; t += nt) {
            const u32 sl = __ldcg(&order[t]);
            const u32 c = __ldcg(&keys[sl]), b = (c - cmin) >> bshift;
            const u32 lo = __ldcg(&bcnt[b]), hi = __ldcg(&bcnt[b + 1]);
            u32 rank = 0;
            for (u32 j = lo; j < hi; j++) rank += __ldcg(&keys[__ldcg(&order[j])]) < c ? 1u : 0u;
            ull v = __ldcg(&vals[sl]);
            if (sizeof(VT) == 4 && v > 0xFFFFFFFFull) v = 0xFFFFFFFFull;
            o.col[obase + lo + rank] = c; put_val(o, obase + lo + rank, (VT)v);
            vmax = vmax > v ? vmax : v;
        }
        __syncthreads();
    }
    vmax = warp_max_u64(vmax);
    if (lane == 0 && vmax) atomicMax(&ctrl->max_val_out, (ull)vmax);
}

// =======================================================================================
// 7. row_ptr: single-pass decoupled look-back exclusive scan of nnz_row (u32 -> u64); its last CTA reports to the host.
//    Tiles of THREADS*ITEMS rows.  Small matrices use 1024 threads x 2 rows: a thread's two row_ptr words are one 16-byte
//    piece of a contiguous 512-byte warp store (eight rows per thread made every store a sector of its own).
// =======================================================================================

template <int THREADS, int ITEMS>
__global__ void __launch_bounds__(THREADS) k_scan_rowptr(u64 rows, const u32 *__restrict__ nnz_row, u64 *__restrict__ rpC,
                                                              u64 *tile_status, B200Ctrl *ctrl,
                                                              u64 *host_mirror = nullptr, u32 epoch = 0) {
    __shared__ u32 s_tile, s_last;
    __shared__ u64 s_wsum[THREADS / 32];
    __shared__ u64 s_excl;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(&ctrl->scan_ticket[0], 1u);            // tiles start in ticket order
    __syncthreads();
    const u32 tile = s_tile;
    const u64 base = (u64)tile * (THREADS * ITEMS) + (u64)tid * ITEMS;
    u32 item[ITEMS]; u64 tsum = 0; u32 tmaxv = 0;
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        item[i] = base + i < rows ? nnz_row[base + i] : 0u;
        tsum += item[i]; tmaxv = item[i] > tmaxv ? item[i] : tmaxv;
    }
    u64 incl = tsum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { u64 t = shfl_up_u64(incl, d); if (lane >= d) incl += t; }
    if (lane == 31) s_wsum[w] = incl;
    __syncthreads();
    u64 wbase = 0, agg = 0;
#pragma unroll
    for (int i = 0; i < THREADS / 32; i++) { if (i < w) wbase += s_wsum[i]; agg += s_wsum[i]; }
    const u64 texcl = wbase + incl - tsum;                                   // exclusive within the tile
    // publish the tile aggregate, then look back over the predecessors
    if (w == 0) {
        if (lane == 0) {
            const u64 st = (tile == 0 ? SCAN_FLAG_PRE : SCAN_FLAG_AGG) | agg;
            atomicExch((ull *)&tile_status[tile], (ull)st);
        }
        u64 excl = 0;
        if (tile > 0) {
            int look = (int)tile - 1;
            while (true) {
                const int idx = look - lane;
                u64 st;
                do {
                    st = idx >= 0 ? ld_volatile_u64(&tile_status[idx]) : SCAN_FLAG_PRE;
                } while (__any_sync(0xFFFFFFFFu, (st >> 62) == 0));
                const u32 pre_mask = __ballot_sync(0xFFFFFFFFu, (st >> 62) == 2);
                const int first = pre_mask ? __ffs(pre_mask) - 1 : 32;      // nearest tile that knows its full prefix
                const u64 contrib = lane <= first ? (st & SCAN_VAL_MASK) : 0ull;
                excl += warp_sum_u64(contrib);
                if (pre_mask) break;
                look -= 32;
            }
            if (lane == 0) atomicExch((ull *)&tile_status[tile], (ull)(SCAN_FLAG_PRE | (excl + agg)));
        }
        if (lane == 0) s_excl = excl;
    }
    __syncthreads();
    u64 run = s_excl + texcl;
    if (tile == 0 && tid == 0) rpC[0] = 0;
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        run += item[i];
        if (base + i < rows) rpC[base + i + 1] = run;
    }
    if (base < rows && base + ITEMS >= rows) ctrl->total_nnz = run;      // thread holding the last row
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) tmaxv = max(tmaxv, __shfl_xor_sync(0xFFFFFFFFu, tmaxv, m));
    if (lane == 0 && tmaxv) atomicMax(&ctrl->max_row_nnz, (ull)tmaxv);
    // Report to the host without a stream synchronise: the last CTA to finish copies the control block into pinned host
    // memory as self-validating 8-byte chunks {word, epoch}.  An aligned 8-byte store reaches the host whole, so the
    // host simply waits until every chunk carries this multiply's epoch: no system-scope fence, no second "ready" word
    // (two fences over PCIe cost ~10 us of the ~17 us this kernel took with them).
    if (host_mirror) {
        __syncthreads();
        if (tid == 0) { __threadfence(); s_last = atomicAdd(&ctrl->scan_done, 1u) == gridDim.x - 1 ? 1u : 0u; }
        __syncthreads();
        if (s_last) {
            __threadfence();
            const volatile u32 *src = reinterpret_cast<const volatile u32 *>(ctrl);
            for (u32 i = tid; i < sizeof(B200Ctrl) / 4; i += blockDim.x) st_volatile_u64(host_mirror + i, ((u64)epoch << 32) | (u64)src[i]);
        }
    }
}

// one-pass mode: move every row from the scratch CSR (bound offsets) to its exact place
template <typename VT, typename SV>
__global__ void __launch_bounds__(256) k_compact_rows(u64 rows, const u64 *__restrict__ src_ptr, const u64 *__restrict__ rpC,
                                                      const u32 *__restrict__ src_col, const SV *__restrict__ src_val,
                                                      u32 *__restrict__ colC, VT *__restrict__ valC, int lanes_lg,
                                                      const ull *max_val_src, ull *max_val_dst,
                                                      u64 *scan_area, u32 ctrl_words, u64 scan_words) {
    const u64 gthread = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    // The last kernel of a multiply: the product's max-value scalar rides along, and the control block and scan status
    // words (scan_area[0..scan_words), the first ctrl_words of them the control block that holds the scalar) are left
    // zeroed, so the next multiply starts without a memset.
    if (blockIdx.x == 0) {
        if (threadIdx.x == 0) *max_val_dst = *max_val_src;
        __syncthreads();
        for (u32 i = threadIdx.x; i < ctrl_words; i += blockDim.x) scan_area[i] = 0;
    }
    for (u64 i = ctrl_words + gthread; i < scan_words; i += (u64)gridDim.x * blockDim.x) scan_area[i] = 0;
    const u32 L = 1u << lanes_lg;                                           // lanes per row
    const u64 nsub = ((u64)gridDim.x * blockDim.x) >> lanes_lg;
    const u32 sub = threadIdx.x & (L - 1);
    for (u64 row = gthread >> lanes_lg; row < rows; row += nsub) {
        const u64 d = rpC[row];
        const u32 n = (u32)(rpC[row + 1] - d);
        const u64 sbase = src_ptr[row];
        u32 t = sub;
        for (; t + 3 * L < n; t += 4 * L) {                                  // eight loads in flight per lane
            const u32 c0 = src_col[sbase + t], c1 = src_col[sbase + t + L], c2 = src_col[sbase + t + 2 * L], c3 = src_col[sbase + t + 3 * L];
            const VT v0 = src_val[sbase + t], v1 = src_val[sbase + t + L], v2 = src_val[sbase + t + 2 * L], v3 = src_val[sbase + t + 3 * L];
            colC[d + t] = c0; colC[d + t + L] = c1; colC[d + t + 2 * L] = c2; colC[d + t + 3 * L] = c3;
            valC[d + t] = v0; valC[d + t + L] = v1; valC[d + t + 2 * L] = v2; valC[d + t + 3 * L] = v3;
        }
        for (; t < n; t += L) { colC[d + t] = src_col[sbase + t]; valC[d + t] = src_val[sbase + t]; }
    }
}

// =======================================================================================
// 8. small utilities: value max / zero check on upload, index narrowing, add, pattern compare
// =======================================================================================
// Largest value, format check, and the circular column range of the matrix: offsets o = (c - col[0] + cols/2) mod cols of
// all its columns, reduced to range[0] = max(~o) (i.e. the minimum, kept as a maximum so that a zeroed word is its
// identity), range[1] = max(o), in four frames (range[2f], range[2f+1]: the circle cut at ref + n/2 + f*(n/4), so that any arc
// of at most three quarters of the circle is seen unbroken by one of them), range[8] = col[0].  A row block of a banded / torus matrix covers a short arc of the
// index circle even when it wraps around its end; the host turns the arc into one bitmap window for the whole multiply.
template <typename VT>
__global__ void __launch_bounds__(256) k_value_stats(u64 nnz, const VT *__restrict__ val, const u32 *__restrict__ col, u64 cols,
                                                     ull *maxval, u32 *bad, u32 *range) {
    u64 m = 0; u32 z = 0;
    u32 ninv[4] = {0, 0, 0, 0}, omax[4] = {0, 0, 0, 0};
    const u32 ref = col[0], half = (u32)(cols / 2), quarter = (u32)(cols / 4);
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (u64)gridDim.x * blockDim.x) {
        const u64 v = val[i];
        const u32 c = col[i];
        m = v > m ? v : m; z |= (v == 0) | ((u64)c >= cols);
        long long t = (long long)c - (long long)ref + (long long)half;
        if (t < 0) t += (long long)cols; else if (t >= (long long)cols) t -= (long long)cols;
#pragma unroll
        for (int f = 0; f < 4; f++) {                                       // frame f: the circle cut f quarters further on
            long long tf = t - (long long)f * quarter;
            if (tf < 0) tf += (long long)cols;
            ninv[f] = max(ninv[f], ~(u32)tf); omax[f] = max(omax[f], (u32)tf);
        }
    }
    m = warp_max_u64(m);
    z = __any_sync(0xFFFFFFFFu, z);
#pragma unroll
    for (int f = 0; f < 4; f++) {
#pragma unroll
        for (int k = 16; k > 0; k >>= 1) { ninv[f] = max(ninv[f], __shfl_xor_sync(0xFFFFFFFFu, ninv[f], k)); omax[f] = max(omax[f], __shfl_xor_sync(0xFFFFFFFFu, omax[f], k)); }
    }
    if ((threadIdx.x & 31) == 0) {
        if (m) atomicMax(maxval, (ull)m);
        if (z) atomicOr(bad, 1u);
#pragma unroll
        for (int f = 0; f < 4; f++) { atomicMax(&range[2 * f], ninv[f]); atomicMax(&range[2 * f + 1], omax[f]); }
        if (blockIdx.x == 0 && threadIdx.x == 0) range[8] = ref;
    }
}

// Operand-wide circular column offsets of a square right operand: u = (c - k + n/2) mod n over the entries of every row k
// with at most B200_CSPAN_MAXLEN entries, reduced to out[0] = max(~u), out[1] = max(u); out[2] counts the longer rows
// (not scanned: with any of them present the operand has no useful bound).
#define B200_CSPAN_MAXLEN 64
__global__ void __launch_bounds__(256) k_cspan_bounds(u64 n, const u64 *__restrict__ rp, const u32 *__restrict__ col, u32 *out) {
    const u32 half = (u32)(n / 2);
    u32 ninv = 0, umax = 0, skipped = 0;
    for (u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (u64)gridDim.x * blockDim.x) {
        const u64 s = rp[k], e = rp[k + 1];
        if (e - s > B200_CSPAN_MAXLEN) { skipped++; continue; }
        for (u64 j = s; j < e; j++) {
            long long t = (long long)col[j] - (long long)k + (long long)half;
            if (t < 0) t += (long long)n; else if (t >= (long long)n) t -= (long long)n;
            ninv = max(ninv, ~(u32)t); umax = max(umax, (u32)t);
        }
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        ninv = max(ninv, __shfl_xor_sync(0xFFFFFFFFu, ninv, m)); umax = max(umax, __shfl_xor_sync(0xFFFFFFFFu, umax, m));
        skipped += __shfl_xor_sync(0xFFFFFFFFu, skipped, m);
    }
    if ((threadIdx.x & 31) == 0) { atomicMax(&out[0], ninv); atomicMax(&out[1], umax); if (skipped) atomicAdd(&out[2], skipped); }
}

// longest row + row_ptr sanity for handles adopted from device arrays (the host never sees their row_ptr)
__global__ void __launch_bounds__(256) k_rowptr_stats(u64 rows, u64 nnz, const u64 *__restrict__ rp, u32 *max_len, u32 *bad) {
    u32 m = 0, z = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += (u64)gridDim.x * blockDim.x) {
        const u64 s = rp[i], e = rp[i + 1];
        z |= (e < s) | (e > nnz) | (i == 0 && s != 0) | (i + 1 == rows && e != nnz);
        const u64 l = e - s;
        m = l > m && e >= s ? (u32)(l > 0xFFFFFFFFull ? 0xFFFFFFFFull : l) : m;
    }
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) m = max(m, __shfl_xor_sync(0xFFFFFFFFu, m, k));
    z = __any_sync(0xFFFFFFFFu, z);
    if ((threadIdx.x & 31) == 0) { if (m) atomicMax(max_len, m); if (z) atomicOr(bad, 2u); }
}

// columns strictly ascending inside every row (format check of every new handle): thread per row, bad |= 4
__global__ void __launch_bounds__(256) k_rows_sorted(u64 rows, const u64 *__restrict__ rp, const u32 *__restrict__ col, u64 nnz, u32 *bad) {
    u32 z = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += (u64)gridDim.x * blockDim.x) {
        const u64 s = rp[i], e = rp[i + 1];
        if (e < s || e > nnz) continue;                                     // reported by the row_ptr checks
        for (u64 j = s + 1; j < e; j++) z |= col[j - 1] >= col[j];
    }
    if (__any_sync(0xFFFFFFFFu, z) && (threadIdx.x & 31) == 0) atomicOr(bad, 4u);
}

__global__ void __launch_bounds__(256) k_narrow_idx(u64 n, const u64 *__restrict__ in, u32 *__restrict__ out, u32 *bad) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const u64 v = in[i];
        if (v >= 0xFFFFFFFFull) atomicOr(bad, 1u);
        out[i] = (u32)v;
    }
}
__global__ void __launch_bounds__(256) k_widen_idx(u64 n, const u32 *__restrict__ in, u64 *__restrict__ out) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) out[i] = in[i];
}
__global__ void __launch_bounds__(256) k_rebase_rowptr(u64 n, const u64 *__restrict__ in, u64 base, u64 *__restrict__ out) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) out[i] = in[i] - base;
}

// element-wise add (src/graph_csr.rs:487-542): count pass then fill pass, one thread per row
template <typename VT, bool FILL>
__global__ void __launch_bounds__(256) k_add_rows(CsrView<VT> A, CsrView<VT> B, u32 *__restrict__ nnz_row, const u64 *__restrict__ rpC,
                                                  u32 *__restrict__ colC, VT *__restrict__ valC, B200Ctrl *ctrl) {
    const u64 row = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 vmax = 0;
    if (row < A.rows) {
        u64 ai = A.rp[row], ae = A.rp[row + 1], bi = B.rp[row], be = B.rp[row + 1];
        u64 pos = FILL ? rpC[row] : 0; u32 n = 0;
        while (ai < ae || bi < be) {
            const u32 ac = ai < ae ? A.col[ai] : 0xFFFFFFFFu, bc = bi < be ? B.col[bi] : 0xFFFFFFFFu;
            u32 c; VT v = 0;
            if (ac < bc) { c = ac; if (FILL) v = A.val[ai]; ai++; }
            else if (bc < ac) { c = bc; if (FILL) v = B.val[bi]; bi++; }
            else { c = ac; if (FILL) v = sat_add(A.val[ai], B.val[bi]); ai++; bi++; }
            if (FILL) { colC[pos] = c; valC[pos] = v; pos++; vmax = vmax > (u64)v ? vmax : (u64)v; }
            n++;
        }
        if (!FILL) nnz_row[row] = n;
    }
    if (FILL) { vmax = warp_max_u64(vmax); if ((threadIdx.x & 31) == 0 && vmax) atomicMax(&ctrl->max_val_out, (ull)vmax); }
}

__global__ void __launch_bounds__(256) k_compare_u32(u64 n, const u32 *__restrict__ a, const u32 *__restrict__ b, u32 *diff) {
    u32 d = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) d |= a[i] != b[i];
    if (__any_sync(0xFFFFFFFFu, d) && (threadIdx.x & 31) == 0) atomicOr(diff, 1u);
}
__global__ void __launch_bounds__(256) k_compare_u64(u64 n, const u64 *__restrict__ a, const u64 *__restrict__ b, u32 *diff) {
    u32 d = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) d |= a[i] != b[i];
    if (__any_sync(0xFFFFFFFFu, d) && (threadIdx.x & 31) == 0) atomicOr(diff, 1u);
}
