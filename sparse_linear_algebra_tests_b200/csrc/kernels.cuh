// kernels.cuh -- device kernels of the SpGEMM engine (sm_100a).
//
// Pipeline of one C = A x B (host orchestration in api.cu):
//   k_row_products      per-row intermediate-product count P_i, symbolic-bin histogram
//   k_bin_scatter       row ids grouped by bin (device-side counts, no host round trip)
//   k_sym_tiny/_hash    exact nnz_i per row (distinct output columns)
//   k_scan_rowptr       hand-written decoupled look-back scan: row_ptr_C (u64), total nnz
//   k_num_classify      numeric-bin histogram on exact nnz_i, then k_bin_scatter again
//   k_num_tiny/_hash    values: saturating accumulate, in-row column order, write col/val
// Heavy rows (beyond a CTA's shared memory) use the *_heavy kernels with a global table.
#pragma once
#include "common.cuh"

// =======================================================================================
// 1. product count per row + symbolic-bin histogram
// =======================================================================================
__device__ __forceinline__ int sym_bin_of(u64 p, u64 dA) {
    if (p == 0 || dA == 1) return B200_BIN_NONE;           // nnz known: 0, or P (one B row: distinct cols)
    if (p <= 32 && dA <= 32) return B200_BIN_TINY;
    return b200_bin_by_size(p);
}

template <int G>  // lanes per row (power of two <= 32)
__global__ void __launch_bounds__(256) k_row_products(u64 rows, const u64 *__restrict__ rpA, const u32 *__restrict__ colA,
                                                      const u64 *__restrict__ rpB, u64 *__restrict__ prod,
                                                      u32 *__restrict__ nnz_row, B200Ctrl *ctrl) {
    __shared__ u32 s_hist[B200_NBINS];
    __shared__ ull s_sum, s_max;
    if (threadIdx.x < B200_NBINS) s_hist[threadIdx.x] = 0;
    if (threadIdx.x == 0) { s_sum = 0; s_max = 0; }
    __syncthreads();
    const u64 gtid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 row = gtid / G;
    const int sub = threadIdx.x % G;
    u64 p = 0, s = 0, e = 0;
    if (row < rows) {
        s = rpA[row]; e = rpA[row + 1];
        for (u64 i = s + sub; i < e; i += G) {
            u32 c = colA[i];
            p += rpB[c + 1] - rpB[c];
        }
    }
#pragma unroll
    for (int m = G / 2; m > 0; m >>= 1) p += shfl_xor_u64(p, m);
    u64 wsum = 0, wmax = 0;
    if (row < rows && sub == 0) {
        prod[row] = p;
        int b = sym_bin_of(p, e - s);
        if (b == B200_BIN_NONE) nnz_row[row] = (u32)p;       // 0, or the single B row's length
        else atomicAdd(&s_hist[b], 1u);
        wsum = p; wmax = p;
    }
    wsum = warp_sum_u64(wsum); wmax = warp_max_u64(wmax);
    if ((threadIdx.x & 31) == 0) { atomicAdd(&s_sum, (ull)wsum); atomicMax(&s_max, (ull)wmax); }
    __syncthreads();
    if (threadIdx.x < B200_NBINS && s_hist[threadIdx.x]) atomicAdd(&ctrl->sym_bin_count[threadIdx.x], s_hist[threadIdx.x]);
    if (threadIdx.x == 0) {
        if (s_sum) atomicAdd(&ctrl->total_products, s_sum);
        atomicMax(&ctrl->max_row_products, s_max);
    }
}

// numeric classification on exact nnz (rows with P<=32 and deg_A<=32 stay in the warp-merge bin)
__device__ __forceinline__ int num_bin_of(u32 nnz, u64 p, u64 dA) {
    if (nnz == 0) return B200_BIN_NONE;
    if (p <= 32 && dA <= 32) return B200_BIN_TINY;
    return b200_bin_by_size(nnz);
}

__global__ void __launch_bounds__(256) k_num_classify(u64 rows, const u64 *__restrict__ rpA, const u64 *__restrict__ prod,
                                                      const u32 *__restrict__ nnz_row, B200Ctrl *ctrl) {
    __shared__ u32 s_hist[B200_NBINS];
    if (threadIdx.x < B200_NBINS) s_hist[threadIdx.x] = 0;
    __syncthreads();
    const u64 row = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (row < rows) {
        int b = num_bin_of(nnz_row[row], prod[row], rpA[row + 1] - rpA[row]);
        if (b != B200_BIN_NONE) atomicAdd(&s_hist[b], 1u);
    }
    __syncthreads();
    if (threadIdx.x < B200_NBINS && s_hist[threadIdx.x]) atomicAdd(&ctrl->num_bin_count[threadIdx.x], s_hist[threadIdx.x]);
}

// scatter row ids into their bin's segment of bin_rows; PHASE 0 = symbolic bins, 1 = numeric bins
template <int PHASE>
__global__ void __launch_bounds__(256) k_bin_scatter(u64 rows, const u64 *__restrict__ rpA, const u64 *__restrict__ prod,
                                                     const u32 *__restrict__ nnz_row, B200Ctrl *ctrl, u32 *__restrict__ bin_rows) {
    __shared__ u32 s_cnt[B200_NBINS], s_base[B200_NBINS];
    if (threadIdx.x < B200_NBINS) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const u64 row = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    int b = B200_BIN_NONE; u32 local = 0;
    if (row < rows) {
        u64 dA = rpA[row + 1] - rpA[row];
        b = PHASE == 0 ? sym_bin_of(prod[row], dA) : num_bin_of(nnz_row[row], prod[row], dA);
        if (b != B200_BIN_NONE) local = atomicAdd(&s_cnt[b], 1u);
    }
    __syncthreads();
    if (threadIdx.x < B200_NBINS) {
        const u32 *cnt = PHASE == 0 ? ctrl->sym_bin_count : ctrl->num_bin_count;
        u32 *fill = PHASE == 0 ? ctrl->sym_bin_fill : ctrl->num_bin_fill;
        u32 off = 0;
        for (int i = 0; i < (int)threadIdx.x; i++) off += cnt[i];
        u32 c = s_cnt[threadIdx.x];
        s_base[threadIdx.x] = off + (c ? atomicAdd(&fill[threadIdx.x], c) : 0u);
    }
    __syncthreads();
    if (b != B200_BIN_NONE) bin_rows[s_base[b] + local] = (u32)row;
}

__device__ __forceinline__ u32 bin_offset(const u32 *cnt, int bin) {
    u32 off = 0;
    for (int i = 0; i < bin; i++) off += cnt[i];
    return off;
}

// =======================================================================================
// 2. tiny rows: one warp per row, <= 32 products held one per lane
// =======================================================================================
// Gathers the row's products into registers: lane p gets product p (column, and the pair of
// operand values when NUMERIC).  Returns P (same in all lanes).
template <typename VT, bool NUMERIC>
__device__ __forceinline__ u32 tiny_gather(const CsrView<VT> &A, const CsrView<VT> &B, u32 row, int lane, u32 &key, VT &val) {
    const u64 s = A.rp[row];
    const u32 dA = (u32)(A.rp[row + 1] - s);               // <= 32 by bin construction
    u32 k = 0, deg = 0; u64 bstart = 0; VT a = 0;
    if (lane < (int)dA) {
        k = A.col[s + lane];
        bstart = B.rp[k];
        deg = (u32)(B.rp[k + 1] - bstart);
        if (NUMERIC) a = A.val[s + lane];
    }
    // exclusive scan of deg across lanes
    u32 incl = deg;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { u32 t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
    const u32 excl = incl - deg;
    const u32 P = __shfl_sync(0xFFFFFFFFu, incl, 31);
    // lane p: last entry e with excl_e <= p (empty B rows share their successor's offset and lose)
    int lo = 0;
#pragma unroll
    for (int step = 16; step > 0; step >>= 1) {
        int cand = lo + step;
        u32 t = __shfl_sync(0xFFFFFFFFu, excl, cand & 31);
        if (cand < (int)dA && t <= (u32)lane) lo = cand;
    }
    const u32 e_excl = __shfl_sync(0xFFFFFFFFu, excl, lo);
    const u64 e_bstart = shfl_u64(bstart, lo);
    VT e_a = 0;
    if (NUMERIC) e_a = shfl_any(a, lo);
    key = B200_EMPTY_KEY; val = 0;
    if ((u32)lane < P) {
        const u64 j = e_bstart + ((u32)lane - e_excl);
        key = B.col[j];
        if (NUMERIC) val = sat_mul(e_a, B.val[j]);
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
            // lower lane of an ascending pair keeps the min; of a descending pair keeps the max
            const bool take_min = (up == lower);
            const bool swap = take_min ? (ok < key) : (ok > key);
            if (swap) { key = ok; if (NUMERIC) val = ov; }
        }
    }
}

template <typename VT>
__global__ void __launch_bounds__(256) k_sym_tiny(CsrView<VT> A, CsrView<VT> B, const u32 *__restrict__ bin_rows,
                                                  B200Ctrl *ctrl, u32 *__restrict__ nnz_row) {
    const u32 count = ctrl->sym_bin_count[B200_BIN_TINY];
    const u32 off = 0;
    const int lane = threadIdx.x & 31;
    const u32 wpb = blockDim.x >> 5;
    for (u32 r = blockIdx.x * wpb + (threadIdx.x >> 5); r < count; r += gridDim.x * wpb) {
        const u32 row = bin_rows[off + r];
        u32 key; VT val;
        tiny_gather<VT, false>(A, B, row, lane, key, val);
        warp_bitonic<VT, false>(key, val, lane);
        const u32 prev = __shfl_up_sync(0xFFFFFFFFu, key, 1);
        const bool head = key != B200_EMPTY_KEY && (lane == 0 || prev != key);
        const u32 n = __popc(__ballot_sync(0xFFFFFFFFu, head));
        if (lane == 0) nnz_row[row] = n;
    }
}

template <typename VT>
__global__ void __launch_bounds__(256) k_num_tiny(CsrView<VT> A, CsrView<VT> B, const u32 *__restrict__ bin_rows,
                                                  B200Ctrl *ctrl, const u64 *__restrict__ rpC, u32 *__restrict__ colC,
                                                  VT *__restrict__ valC) {
    const u32 count = ctrl->num_bin_count[B200_BIN_TINY];
    const int lane = threadIdx.x & 31;
    const u32 wpb = blockDim.x >> 5;
    u64 vmax = 0;
    for (u32 r = blockIdx.x * wpb + (threadIdx.x >> 5); r < count; r += gridDim.x * wpb) {
        const u32 row = bin_rows[r];
        u32 key; VT val;
        tiny_gather<VT, true>(A, B, row, lane, key, val);
        warp_bitonic<VT, true>(key, val, lane);
        // segmented inclusive scan (saturating) over runs of equal keys; the run's last lane has the total
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u32 pk = __shfl_up_sync(0xFFFFFFFFu, key, d);
            const VT pv = shfl_up_any(val, d);
            if (lane >= d && pk == key) val = sat_add(val, pv);
        }
        const u32 nk = __shfl_down_sync(0xFFFFFFFFu, key, 1);
        const bool tail = key != B200_EMPTY_KEY && (lane == 31 || nk != key);
        const u32 tails = __ballot_sync(0xFFFFFFFFu, tail);
        if (tail) {
            const u64 pos = rpC[row] + __popc(tails & ((1u << lane) - 1u));
            colC[pos] = key; valC[pos] = val;
            vmax = vmax > (u64)val ? vmax : (u64)val;
        }
    }
    vmax = warp_max_u64(vmax);
    if (lane == 0 && vmax) atomicMax(&ctrl->max_val_out, (ull)vmax);
}

// =======================================================================================
// 3. hash rows -- symbolic: distinct columns via a shared-memory key table, or via a
//    shared-memory column bitmap when the whole column space fits (BITMAP)
// =======================================================================================
// One "group" of `blockDim.x` threads owns one row at a time (grid-stride over the bin).
// Within the group, `1<<lg` lanes cooperate on one A entry and stride over its B row.
template <typename VT, bool BITMAP>
__global__ void __launch_bounds__(1024) k_sym_hash(CsrView<VT> A, CsrView<VT> B, const u32 *__restrict__ bin_rows, B200Ctrl *ctrl, int bin,
                           u32 slots, u32 nwords, int lg, u32 *__restrict__ nnz_row) {
    extern __shared__ u32 smem[];
    __shared__ u32 s_count;
    const u32 count = ctrl->sym_bin_count[bin];
    const u32 off = bin_offset(ctrl->sym_bin_count, bin);
    const int nt = blockDim.x, tid = threadIdx.x;
    const int G = 1 << lg, sub = tid & (G - 1);
    const u32 tabn = BITMAP ? nwords : slots;
    const int shift = 32 - (31 - __clz(slots));
    for (u32 r = blockIdx.x; r < count; r += gridDim.x) {
        const u32 row = bin_rows[off + r];
        for (u32 t = tid; t < tabn; t += nt) smem[t] = BITMAP ? 0u : B200_EMPTY_KEY;
        if (tid == 0) s_count = 0;
        __syncthreads();
        const u64 s = A.rp[row], e = A.rp[row + 1];
        u32 local = 0;
        for (u64 ia = s + (tid >> lg); ia < e; ia += (nt >> lg)) {
            const u32 k = A.col[ia];
            const u64 bs = B.rp[k], be = B.rp[k + 1];
            for (u64 jb = bs + sub; jb < be; jb += G) {
                const u32 c = B.col[jb];
                if (BITMAP) {
                    const u32 bit = 1u << (c & 31);
                    const u32 old = atomicOr(&smem[c >> 5], bit);
                    local += !(old & bit);
                } else {
                    u32 h = b200_hash(c, shift);
                    while (true) {
                        u32 cur = ld_volatile_u32(&smem[h]);
                        if (cur == c) break;
                        if (cur == B200_EMPTY_KEY) {
                            cur = atomicCAS(&smem[h], B200_EMPTY_KEY, c);
                            if (cur == B200_EMPTY_KEY) { local++; break; }
                            if (cur == c) break;
                        }
                        h = (h + 1) & (slots - 1);
                    }
                }
            }
        }
        local = warp_sum_u32(local);
        if ((tid & 31) == 0 && local) atomicAdd(&s_count, local);
        __syncthreads();
        if (tid == 0) nnz_row[row] = s_count;
        __syncthreads();
    }
}

// =======================================================================================
// 4. hash rows -- numeric
// =======================================================================================
// MODE 0: 32-bit accumulators (host proved max_row_products*max(A)*max(B) < 2^32)
// MODE 1: 64-bit accumulators, plain adds (u64: proved < 2^64; u32: products clamped to 2^32-1,
//         sums of < 2^32 such terms cannot wrap 64 bits, clamp on emit)
// MODE 2: u64 saturating multiply + CAS-loop saturating add
template <int MODE> struct AccOf { typedef u64 type; };
template <> struct AccOf<0> { typedef u32 type; };

template <typename VT, int MODE>
__device__ __forceinline__ typename AccOf<MODE>::type make_product(VT a, VT b) {
    if (MODE == 0) return (u32)a * (u32)b;
    if (MODE == 1) {
        if (sizeof(VT) == 4) { u64 p = (u64)a * (u64)b; return p > 0xFFFFFFFFull ? 0xFFFFFFFFull : p; }
        return (u64)a * (u64)b;
    }
    return sat_mul((u64)a, (u64)b);
}
template <int MODE>
__device__ __forceinline__ void acc_add(typename AccOf<MODE>::type *p, typename AccOf<MODE>::type x) {
    if (MODE == 0) atomicAdd((u32 *)p, (u32)x);
    else if (MODE == 1) atomicAdd((ull *)p, (ull)x);
    else {
        ull *q = (ull *)p;
        ull old = *reinterpret_cast<volatile ull *>(q), assumed;
        do {
            assumed = old;
            ull s = assumed + (ull)x; if (s < assumed) s = ~0ull;
            if (s == assumed) break;
            old = atomicCAS(q, assumed, s);
        } while (old != assumed);
    }
}
template <typename VT, int MODE>
__device__ __forceinline__ VT emit_val(typename AccOf<MODE>::type v) {
    if (sizeof(VT) == 4 && MODE == 1) return (VT)(v > 0xFFFFFFFFull ? 0xFFFFFFFFull : v);
    return (VT)v;
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

// in-shared-memory bitonic sort of n (power of two) key/value pairs
template <typename AccT>
__device__ __forceinline__ void smem_bitonic(u32 *keys, AccT *vals, u32 n) {
    for (u32 k = 2; k <= n; k <<= 1) {
        for (u32 j = k >> 1; j > 0; j >>= 1) {
            for (u32 t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
                const u32 i = 2 * t - (t & (j - 1));
                const u32 l = i + j;
                const bool up = ((i & k) == 0);
                const u32 a = keys[i], b = keys[l];
                if ((a > b) == up) {
                    keys[i] = b; keys[l] = a;
                    const AccT va = vals[i]; vals[i] = vals[l]; vals[l] = va;
                }
            }
            __syncthreads();
        }
    }
}

template <typename VT, int MODE, bool BITMAP>
__global__ void __launch_bounds__(1024) k_num_hash(CsrView<VT> A, CsrView<VT> B, const u32 *__restrict__ bin_rows, B200Ctrl *ctrl, int bin,
                           u32 slots, u32 nwords, int lg, const u64 *__restrict__ rpC, u32 *__restrict__ colC,
                           VT *__restrict__ valC) {
    typedef typename AccOf<MODE>::type AccT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ u32 s_warp[33];
    AccT *vals = reinterpret_cast<AccT *>(smem_raw);                       // slots
    u32 *keys = reinterpret_cast<u32 *>(vals + slots);                     // slots
    u32 *bm = keys + slots;                                                // nwords   (BITMAP)
    u32 *wpre = bm + nwords;                                               // nwords   (BITMAP)
    const u32 count = ctrl->num_bin_count[bin];
    const u32 off = bin_offset(ctrl->num_bin_count, bin);
    const int nt = blockDim.x, tid = threadIdx.x;
    const int G = 1 << lg, sub = tid & (G - 1);
    const int shift = 32 - (31 - __clz(slots));
    u64 vmax = 0;
    for (u32 r = blockIdx.x; r < count; r += gridDim.x) {
        const u32 row = bin_rows[off + r];
        for (u32 t = tid; t < slots; t += nt) { keys[t] = B200_EMPTY_KEY; vals[t] = 0; }
        if (BITMAP) for (u32 t = tid; t < nwords; t += nt) bm[t] = 0;
        __syncthreads();
        const u64 s = A.rp[row], e = A.rp[row + 1];
        for (u64 ia = s + (tid >> lg); ia < e; ia += (nt >> lg)) {
            const u32 k = A.col[ia];
            const VT a = A.val[ia];
            const u64 bs = B.rp[k], be = B.rp[k + 1];
            for (u64 jb = bs + sub; jb < be; jb += G) {
                const u32 c = B.col[jb];
                const AccT x = make_product<VT, MODE>(a, B.val[jb]);
                u32 h = b200_hash(c, shift);
                while (true) {
                    u32 cur = ld_volatile_u32(&keys[h]);
                    if (cur == B200_EMPTY_KEY) {
                        cur = atomicCAS(&keys[h], B200_EMPTY_KEY, c);
                        if (cur == B200_EMPTY_KEY) {
                            if (BITMAP) atomicOr(&bm[c >> 5], 1u << (c & 31));
                            cur = c;
                        }
                    }
                    if (cur == c) { acc_add<MODE>(&vals[h], x); break; }
                    h = (h + 1) & (slots - 1);
                }
            }
        }
        __syncthreads();
        const u64 obase = rpC[row];
        if (BITMAP) {
            // rank of a column = number of set bits below it: prefix popcount over the bitmap words
            u32 carry = 0;
            for (u32 base = 0; base < nwords; base += nt) {
                const u32 w = base + tid < nwords ? bm[base + tid] : 0u;
                u32 total;
                const u32 ex = block_excl_scan(__popc(w), s_warp, total);
                if (base + tid < nwords) wpre[base + tid] = carry + ex;
                carry += total;
            }
            __syncthreads();
            for (u32 t = tid; t < slots; t += nt) {
                const u32 c = keys[t];
                if (c != B200_EMPTY_KEY) {
                    const u32 w = bm[c >> 5];
                    const u64 pos = obase + wpre[c >> 5] + __popc(w & ((1u << (c & 31)) - 1u));
                    const VT v = emit_val<VT, MODE>(vals[t]);
                    colC[pos] = c; valC[pos] = v;
                    vmax = vmax > (u64)v ? vmax : (u64)v;
                }
            }
        } else {
            // compact the occupied slots to the front (through registers), sort by column, stream out
            const u32 per = slots / nt;                                    // 4 .. 16
            u32 rk[16]; AccT rv[16]; u32 mine = 0;
#pragma unroll
            for (int i = 0; i < 16; i++) {
                if (i < (int)per) {
                    rk[i] = keys[tid * per + i]; rv[i] = vals[tid * per + i];
                    mine += rk[i] != B200_EMPTY_KEY;
                }
            }
            u32 total;
            u32 pos = block_excl_scan(mine, s_warp, total);             // contains the barriers that protect the reads above
            u32 n2 = 1; while (n2 < total) n2 <<= 1;
#pragma unroll
            for (int i = 0; i < 16; i++) {
                if (i < (int)per && rk[i] != B200_EMPTY_KEY) { keys[pos] = rk[i]; vals[pos] = rv[i]; pos++; }
            }
            __syncthreads();
            for (u32 t = total + tid; t < n2; t += nt) keys[t] = B200_EMPTY_KEY;
            __syncthreads();
            smem_bitonic<AccT>(keys, vals, n2);
            for (u32 t = tid; t < total; t += nt) {
                const VT v = emit_val<VT, MODE>(vals[t]);
                colC[obase + t] = keys[t]; valC[obase + t] = v;
                vmax = vmax > (u64)v ? vmax : (u64)v;
            }
        }
        __syncthreads();
    }
    vmax = warp_max_u64(vmax);
    if ((tid & 31) == 0 && vmax) atomicMax(&ctrl->max_val_out, (ull)vmax);
}

// ---------------------------------------------------------------------------------------
// 4b. numeric for rows whose whole column space fits a shared-memory bitmap ("rank" kernel).
// No hash table: walk 1 sets one bit per product column; a prefix popcount over the bitmap
// words gives every present column its rank (= its position in the sorted output row); walk 2
// adds each product into acc[rank]; columns are emitted by enumerating the set bits, values by
// streaming acc[0..nnz).  Shared memory: nwords*6 + cap*4 (MODE 0) bytes -> high occupancy.
// MODE 1 keeps a 64-bit sum as two u32 words (lo/hi) and propagates the carry with the value
// returned by the atomic on lo: shared-memory u64 atomicAdd is a CAS spin loop on sm_100
// (ATOMS.CAST.SPIN.64, measured 7x slower than ATOMS.ADD).
// ---------------------------------------------------------------------------------------
template <typename VT, int MODE>
__global__ void __launch_bounds__(1024) k_num_rank(CsrView<VT> A, CsrView<VT> B, const u32 *__restrict__ bin_rows, B200Ctrl *ctrl,
                                                   int bin, u32 cap, u32 nwords, int lg, const u64 *__restrict__ rpC,
                                                   u32 *__restrict__ colC, VT *__restrict__ valC) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ u32 s_warp[33];
    // layout: [acc64 cap (MODE 2)] | lo cap | [hi cap (MODE 1)] | bm nwords | wpre nwords (u16)
    ull *acc64 = reinterpret_cast<ull *>(smem_raw);
    u32 *lo = reinterpret_cast<u32 *>(smem_raw + (MODE == 2 ? (size_t)cap * 8 : 0));
    u32 *hi = lo + (MODE == 2 ? 0 : cap);
    u32 *bm = hi + (MODE == 1 ? cap : 0);
    unsigned short *wpre = reinterpret_cast<unsigned short *>(bm + nwords);
    const u32 count = ctrl->num_bin_count[bin];
    const u32 off = bin_offset(ctrl->num_bin_count, bin);
    const int nt = blockDim.x, tid = threadIdx.x;
    const int G = 1 << lg, sub = tid & (G - 1), grp = tid >> lg, ngrp = nt >> lg;
    const u32 wpt = (nwords + nt - 1) / nt;                               // bitmap words per thread
    u64 vmax = 0;
    for (u32 r = blockIdx.x; r < count; r += gridDim.x) {
        const u32 row = bin_rows[off + r];
        const u64 obase = rpC[row];
        const u32 nnz = (u32)(rpC[row + 1] - obase);
        for (u32 t = tid; t < nwords; t += nt) bm[t] = 0;
        for (u32 t = tid; t < nnz; t += nt) {
            if (MODE == 2) acc64[t] = 0; else lo[t] = 0;
            if (MODE == 1) hi[t] = 0;
        }
        __syncthreads();
        const u64 s = A.rp[row], e = A.rp[row + 1];
        // ---- walk 1: column bitmap
        for (u64 ia = s + grp; ia < e; ia += ngrp) {
            const u32 k = A.col[ia];
            const u64 bs = B.rp[k], be = B.rp[k + 1];
            for (u64 jb = bs + sub; jb < be; jb += G) {
                const u32 c = B.col[jb];
                atomicOr(&bm[c >> 5], 1u << (c & 31));
            }
        }
        __syncthreads();
        // ---- ranks: exclusive prefix popcount over the words (each thread owns wpt consecutive words)
        {
            const u32 w0 = tid * wpt;
            u32 mine = 0;
            for (u32 i = 0; i < wpt; i++) if (w0 + i < nwords) mine += __popc(bm[w0 + i]);
            u32 total;
            u32 run = block_excl_scan(mine, s_warp, total);
            for (u32 i = 0; i < wpt; i++) {
                if (w0 + i < nwords) {
                    const u32 w = bm[w0 + i];
                    wpre[w0 + i] = (unsigned short)run;
                    // columns come out in ascending order: emit them here
                    u32 bits = w;
                    u64 p = obase + run;
                    while (bits) { const int b = __ffs(bits) - 1; bits &= bits - 1; colC[p++] = ((w0 + i) << 5) + b; }
                    run += __popc(w);
                }
            }
        }
        __syncthreads();
        // ---- walk 2: accumulate into acc[rank(col)]
        for (u64 ia = s + grp; ia < e; ia += ngrp) {
            const u32 k = A.col[ia];
            const VT a = A.val[ia];
            const u64 bs = B.rp[k], be = B.rp[k + 1];
            for (u64 jb = bs + sub; jb < be; jb += G) {
                const u32 c = B.col[jb];
                const u32 w = bm[c >> 5];
                const u32 pos = (u32)wpre[c >> 5] + __popc(w & ((1u << (c & 31)) - 1u));
                if (MODE == 0) {
                    atomicAdd(&lo[pos], (u32)a * (u32)B.val[jb]);
                } else if (MODE == 1) {
                    u64 x;
                    if (sizeof(VT) == 4) { x = (u64)a * (u64)B.val[jb]; x = x > 0xFFFFFFFFull ? 0xFFFFFFFFull : x; }
                    else x = (u64)a * (u64)B.val[jb];
                    const u32 xlo = (u32)x, xhi = (u32)(x >> 32);
                    const u32 old = atomicAdd(&lo[pos], xlo);
                    const u32 up = xhi + ((u32)(old + xlo) < xlo ? 1u : 0u);
                    if (up) atomicAdd(&hi[pos], up);
                } else {
                    acc_add<2>((u64 *)&acc64[pos], sat_mul((u64)a, (u64)B.val[jb]));
                }
            }
        }
        __syncthreads();
        for (u32 t = tid; t < nnz; t += nt) {
            u64 v;
            if (MODE == 0) v = lo[t];
            else if (MODE == 1) v = ((u64)hi[t] << 32) | lo[t];
            else v = acc64[t];
            if (sizeof(VT) == 4 && v > 0xFFFFFFFFull) v = 0xFFFFFFFFull;
            valC[obase + t] = (VT)v;
            vmax = vmax > v ? vmax : v;
        }
        __syncthreads();
    }
    vmax = warp_max_u64(vmax);
    if ((tid & 31) == 0 && vmax) atomicMax(&ctrl->max_val_out, (ull)vmax);
}

// =======================================================================================
// 5. heavy rows: table, bitmap and rank array in global scratch (one CTA per row at a time)
// =======================================================================================
template <typename VT>
__global__ void __launch_bounds__(1024) k_sym_heavy(CsrView<VT> A, CsrView<VT> B, const u32 *__restrict__ bin_rows,
                                                    B200Ctrl *ctrl, u32 nwords, u32 *__restrict__ scratch_bm,
                                                    u32 *__restrict__ nnz_row) {
    __shared__ u32 s_count;
    const u32 count = ctrl->sym_bin_count[B200_BIN_HEAVY];
    const u32 off = bin_offset(ctrl->sym_bin_count, B200_BIN_HEAVY);
    u32 *bm = scratch_bm + (u64)blockIdx.x * nwords;
    const int nt = blockDim.x, tid = threadIdx.x, lane = tid & 31, nwarp = nt >> 5, w = tid >> 5;
    for (u32 r = blockIdx.x; r < count; r += gridDim.x) {
        const u32 row = bin_rows[off + r];
        for (u32 t = tid; t < nwords; t += nt) bm[t] = 0;
        if (tid == 0) s_count = 0;
        __syncthreads();
        const u64 s = A.rp[row], e = A.rp[row + 1];
        u32 local = 0;
        for (u64 ia = s + w; ia < e; ia += nwarp) {                        // a warp per A entry
            const u32 k = A.col[ia];
            const u64 bs = B.rp[k], be = B.rp[k + 1];
            for (u64 jb = bs + lane; jb < be; jb += 32) {
                const u32 c = B.col[jb];
                const u32 bit = 1u << (c & 31);
                const u32 old = atomicOr(&bm[c >> 5], bit);
                local += !(old & bit);
            }
        }
        local = warp_sum_u32(local);
        if (lane == 0 && local) atomicAdd(&s_count, local);
        __syncthreads();
        if (tid == 0) nnz_row[row] = s_count;
        __syncthreads();
    }
}

template <typename VT, int MODE>
__global__ void __launch_bounds__(1024) k_num_heavy(CsrView<VT> A, CsrView<VT> B, const u32 *__restrict__ bin_rows,
                                                    B200Ctrl *ctrl, const u32 *__restrict__ nnz_row, u32 nwords, u64 max_slots,
                                                    u32 *__restrict__ scratch_bm, u32 *__restrict__ scratch_pre,
                                                    u32 *__restrict__ scratch_keys, u64 *__restrict__ scratch_vals,
                                                    const u64 *__restrict__ rpC, u32 *__restrict__ colC, VT *__restrict__ valC) {
    __shared__ u32 s_warp[33];
    const u32 count = ctrl->num_bin_count[B200_BIN_HEAVY];
    const u32 off = bin_offset(ctrl->num_bin_count, B200_BIN_HEAVY);
    u32 *bm = scratch_bm + (u64)blockIdx.x * nwords;
    u32 *wpre = scratch_pre + (u64)blockIdx.x * nwords;
    u32 *keys = scratch_keys + (u64)blockIdx.x * max_slots;
    ull *vals = (ull *)(scratch_vals + (u64)blockIdx.x * max_slots);
    const int nt = blockDim.x, tid = threadIdx.x, lane = tid & 31, nwarp = nt >> 5, w = tid >> 5;
    u64 vmax = 0;
    for (u32 r = blockIdx.x; r < count; r += gridDim.x) {
        const u32 row = bin_rows[off + r];
        const u32 nnz = nnz_row[row];
        u64 slots = 1; while (slots < 2ull * nnz) slots <<= 1;            // <= max_slots by host sizing
        const int shift = 64 - (63 - __clzll(slots));
        for (u64 t = tid; t < slots; t += nt) { keys[t] = B200_EMPTY_KEY; vals[t] = 0; }
        for (u32 t = tid; t < nwords; t += nt) bm[t] = 0;
        __syncthreads();
        const u64 s = A.rp[row], e = A.rp[row + 1];
        for (u64 ia = s + w; ia < e; ia += nwarp) {
            const u32 k = A.col[ia];
            const VT a = A.val[ia];
            const u64 bs = B.rp[k], be = B.rp[k + 1];
            for (u64 jb = bs + lane; jb < be; jb += 32) {
                const u32 c = B.col[jb];
                ull x;
                if (MODE == 2) x = sat_mul((u64)a, (u64)B.val[jb]);
                else if (sizeof(VT) == 4) { u64 p = (u64)a * (u64)B.val[jb]; x = p > 0xFFFFFFFFull ? 0xFFFFFFFFull : p; }
                else x = (u64)a * (u64)B.val[jb];
                u64 h = ((u64)c * 0x9E3779B97F4A7C15ull) >> shift;
                while (true) {
                    u32 cur = ld_volatile_u32(&keys[h]);
                    if (cur == B200_EMPTY_KEY) {
                        cur = atomicCAS(&keys[h], B200_EMPTY_KEY, c);
                        if (cur == B200_EMPTY_KEY) { atomicOr(&bm[c >> 5], 1u << (c & 31)); cur = c; }
                    }
                    if (cur == c) {
                        if (MODE == 2) acc_add<2>((u64 *)&vals[h], (u64)x);
                        else atomicAdd(&vals[h], x);
                        break;
                    }
                    h = (h + 1) & (slots - 1);
                }
            }
        }
        __threadfence_block();
        __syncthreads();
        u32 carry = 0;
        for (u32 base = 0; base < nwords; base += nt) {
            const u32 wd = base + tid < nwords ? bm[base + tid] : 0u;
            u32 total;
            const u32 ex = block_excl_scan(__popc(wd), s_warp, total);
            if (base + tid < nwords) wpre[base + tid] = carry + ex;
            carry += total;
        }
        __syncthreads();
        const u64 obase = rpC[row];
        for (u64 t = tid; t < slots; t += nt) {
            const u32 c = keys[t];
            if (c != B200_EMPTY_KEY) {
                const u32 wd = bm[c >> 5];
                const u64 pos = obase + wpre[c >> 5] + __popc(wd & ((1u << (c & 31)) - 1u));
                ull v = vals[t];
                if (sizeof(VT) == 4 && v > 0xFFFFFFFFull) v = 0xFFFFFFFFull;
                colC[pos] = c; valC[pos] = (VT)v;
                vmax = vmax > v ? vmax : v;
            }
        }
        __syncthreads();
    }
    vmax = warp_max_u64(vmax);
    if (lane == 0 && vmax) atomicMax(&ctrl->max_val_out, (ull)vmax);
}

// =======================================================================================
// 6. row_ptr: single-pass decoupled look-back exclusive scan of nnz_row (u32 -> u64)
// =======================================================================================
#define SCAN_THREADS 256
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)
#define SCAN_FLAG_AGG (1ull << 62)
#define SCAN_FLAG_PRE (2ull << 62)
#define SCAN_VAL_MASK ((1ull << 62) - 1)

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_rowptr(u64 rows, const u32 *__restrict__ nnz_row, u64 *__restrict__ rpC,
                                                              u64 *tile_status, B200Ctrl *ctrl) {
    __shared__ u32 s_tile;
    __shared__ u64 s_wsum[SCAN_THREADS / 32];
    __shared__ u64 s_excl;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(&ctrl->scan_ticket, 1u);               // tiles start in ticket order
    __syncthreads();
    const u32 tile = s_tile;
    const u64 base = (u64)tile * SCAN_TILE + (u64)tid * SCAN_ITEMS;
    u32 item[SCAN_ITEMS]; u64 tsum = 0; u32 tmaxv = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
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
    for (int i = 0; i < SCAN_THREADS / 32; i++) { if (i < w) wbase += s_wsum[i]; agg += s_wsum[i]; }
    const u64 texcl = wbase + incl - tsum;                                   // exclusive within the tile
    // publish, then look back
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
                u64 contrib = lane <= first ? (st & SCAN_VAL_MASK) : 0ull;
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
    for (int i = 0; i < SCAN_ITEMS; i++) {
        run += item[i];
        if (base + i < rows) rpC[base + i + 1] = run;
    }
    if (base < rows && base + SCAN_ITEMS >= rows) ctrl->total_nnz = run;   // the thread holding the last row
    tmaxv = max(tmaxv, __shfl_xor_sync(0xFFFFFFFFu, tmaxv, 16)); tmaxv = max(tmaxv, __shfl_xor_sync(0xFFFFFFFFu, tmaxv, 8));
    tmaxv = max(tmaxv, __shfl_xor_sync(0xFFFFFFFFu, tmaxv, 4)); tmaxv = max(tmaxv, __shfl_xor_sync(0xFFFFFFFFu, tmaxv, 2));
    tmaxv = max(tmaxv, __shfl_xor_sync(0xFFFFFFFFu, tmaxv, 1));
    if (lane == 0 && tmaxv) atomicMax(&ctrl->max_row_nnz, (ull)tmaxv);
}

// =======================================================================================
// 7. small utilities: value max / zero check on upload, index narrowing, add, pattern compare
// =======================================================================================
template <typename VT>
__global__ void __launch_bounds__(256) k_value_stats(u64 nnz, const VT *__restrict__ val, const u32 *__restrict__ col, u64 cols,
                                                     ull *maxval, u32 *bad) {
    u64 m = 0; u32 z = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (u64)gridDim.x * blockDim.x) {
        const u64 v = val[i];
        m = v > m ? v : m; z |= (v == 0) | ((u64)col[i] >= cols);
    }
    m = warp_max_u64(m);
    z = __any_sync(0xFFFFFFFFu, z);
    if ((threadIdx.x & 31) == 0) { if (m) atomicMax(maxval, (ull)m); if (z) atomicOr(bad, 1u); }
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
