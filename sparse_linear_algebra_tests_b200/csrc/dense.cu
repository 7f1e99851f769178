// dense.cu -- the multiply as ONE pass over the intermediate products (pipeline 5).
//
// Every other pipeline touches a product two or three times (count its column, mark it again, accumulate it at its rank).
// Here a product is touched once: a CTA owns a row, the row's columns lie on an arc of the index circle that the CTA finds
// from the row's A columns (largest circular gap) and the right operand's offset range [cs_lo, cs_hi] (c - k over all of B's
// entries, a per-operand constant known on the host), and that arc is a DENSE window in shared memory -- one 32-bit
// accumulator per window column plus a bitmap.  A product is then one red.add at acc[c - origin] and one red.or in the
// bitmap; the row's length is the bitmap's popcount, its entries come out of the accumulators in bitmap order (ascending
// columns, no rank lookup, no sort).  Where the row goes in C is not known beforehand, so the rows are placed by a
// decoupled look-back over ROWS: CTA b takes rows b, b + G, b + 2G, ... (all G CTAs co-resident: cooperative launch, no
// grid sync), publishes the row's length as soon as it is known and sums the predecessors' published values -- the
// row_ptr scan, the count pass and the numeric pass of CsrMatrix::matmul_par (/root/reference/src/graph_csr.rs:362-476) in
// one kernel.  Row header, A entries and B records are fetched two / one rows ahead, so an iteration waits for none of them.
//
// Applies when: 32-bit sums are proven (mode 0), the right operand is square, low-degree (sector-packed records) with a
// known offset range, and C can be allocated from the host-known bound.  Rows whose arc is wider than the window are
// produced in column pieces (their products enumerated once per piece, twice for the count).
#include <cooperative_groups.h>
#include "engine.cuh"
#include "devutil.cuh"

#define DN_THREADS 256
#define DN_WARPS (DN_THREADS / 32)

template <typename VT>
struct DnArgs {
    NumArgs<VT> a; const uint4 *pack;
    u64 rows; u32 ncols;
    long long cs_lo, cs_hi;          // offsets c - k of B's entries lie in [cs_lo, cs_hi]
    u32 nolook;                      // developer probe: skip the look-back (rows land at offset 0: timing only)
    u32 wmax;                        // window columns a CTA holds (multiple of 32 * DN_THREADS / 4 ... see host)
    u64 *status;                     // per row: flag << 62 | value (decoupled look-back)
    u64 *rpC; u32 *colC; VT *valC;
    B200Ctrl *ctrl; u64 *host_mirror; u32 epoch; ull *maxval_dst;
};

__device__ __forceinline__ PackRec dn_load_pack(const uint4 *__restrict__ pack, u32 k) {
    PackRec r;
    asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(r.a.x), "=r"(r.a.y), "=r"(r.a.z), "=r"(r.a.w), "=r"(r.b.x), "=r"(r.b.y), "=r"(r.b.z), "=r"(r.b.w)
        : "l"(pack + 2 * (u64)k));
    return r;
}
__device__ __forceinline__ void dn_red_add(u32 addr, u32 v) { asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void dn_red_or(u32 addr, u32 v) { asm volatile("red.shared.or.b32 [%0], %1;" :: "r"(addr), "r"(v) : "memory"); }

static __device__ __forceinline__ u32 dn_block_scan(u32 v, u32 *s_warp /* DN_WARPS + 1 */, u32 &total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    u32 incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const u32 t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
    if (lane == 31) s_warp[w] = incl;
    __syncthreads();
    u32 base = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < DN_WARPS; i++) { const u32 x = s_warp[i]; if (i < w) base += x; tot += x; }
    total = tot;
    __syncthreads();
    return base + incl - v;
}

template <typename VT, bool BPAT>
__global__ void __launch_bounds__(DN_THREADS, 3) k_dense(DnArgs<VT> p) {
    extern __shared__ __align__(16) unsigned char dn_smem[];
    __shared__ u32 s_warp[DN_WARPS + 1];
    __shared__ ull s_gap[2 * DN_WARPS];
    __shared__ u32 s_r0, s_last, s_P;
    __shared__ u64 s_excl;
    u32 *acc = reinterpret_cast<u32 *>(dn_smem);
    u32 *bm = acc + p.wmax;
    const u32 sm_acc = (u32)__cvta_generic_to_shared(acc), sm_bm = (u32)__cvta_generic_to_shared(bm);
    const u32 tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const u32 n = p.ncols;
    const u32 G = gridDim.x;
    const u32 bwords = p.wmax >> 5;
    for (u32 t = tid; t < p.wmax; t += DN_THREADS) acc[t] = 0;
    for (u32 t = tid; t < bwords; t += DN_THREADS) bm[t] = 0;
    __syncthreads();

    u64 vmax = 0, psum = 0; u32 pmax = 0, nmax = 0, row_par = 0;
    if (tid == 0) s_P = 0;
    // ---- software pipeline over this CTA's rows: header two rows ahead, first 256 A entries and their records one row ahead
    u64 row = blockIdx.x;
    u64 rs1 = 0; u32 len1 = 0;                                  // header of `row`
    u64 rs2 = 0; u32 len2 = 0;                                  // header of row + G
    u32 k1 = 0, kn1 = 0; VT a1 = 0; PackRec rec1; rec1.a.y = 0;  // entry tid of `row`
    if (row < p.rows) { rs1 = p.a.rpA[row]; len1 = (u32)(p.a.rpA[row + 1] - rs1); }
    if (row + G < p.rows) { rs2 = p.a.rpA[row + G]; len2 = (u32)(p.a.rpA[row + G + 1] - rs2); }
    if (tid < len1) {
        k1 = p.a.colA[rs1 + tid]; kn1 = tid + 1 < len1 ? p.a.colA[rs1 + tid + 1] : p.a.colA[rs1];
        a1 = p.a.valA[rs1 + tid];
        rec1 = dn_load_pack(p.pack, k1);
    }
    for (; row < p.rows; row += G) {
        const u64 rs = rs1; const u32 lenA = len1;
        const u32 k0 = k1, kn0 = kn1; const VT a0 = a1; const PackRec rec0 = rec1;
        // next row's entries (their header arrived an iteration ago), the header after that
        rs1 = rs2; len1 = len2;
        k1 = 0; kn1 = 0; a1 = 0; rec1.a.y = 0;
        if (row + G < p.rows && tid < len1) {
            k1 = p.a.colA[rs1 + tid]; kn1 = tid + 1 < len1 ? p.a.colA[rs1 + tid + 1] : p.a.colA[rs1];
            a1 = p.a.valA[rs1 + tid];
        }
        rs2 = 0; len2 = 0;
        if (row + 2 * (u64)G < p.rows) { rs2 = p.a.rpA[row + 2 * (u64)G]; len2 = (u32)(p.a.rpA[row + 2 * (u64)G + 1] - rs2); }

        // ---- the row's arc: largest circular gap between consecutive A columns (entries beyond the first 256 are read here)
        ull best = 0;                                                        // gap << 32 | column after the gap
        if (tid < lenA) {
            const u32 gap = tid + 1 < lenA ? kn0 - k0 : kn0 + n - k0;
            best = ((ull)gap << 32) | kn0;
        }
        for (u32 t = DN_THREADS + tid; t < lenA; t += DN_THREADS) {
            const u32 kc = p.a.colA[rs + t], kx = t + 1 < lenA ? p.a.colA[rs + t + 1] : p.a.colA[rs];
            const u32 gap = t + 1 < lenA ? kx - kc : kx + n - kc;
            const ull cand = ((ull)gap << 32) | kx;
            best = cand > best ? cand : best;
        }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) { const ull o = shfl_xor_u64(best, m); best = o > best ? o : best; }
        if (lane == 0) s_gap[(row_par << 3) + wid] = best;                  // (two sets of slots: the next row's writes do not wait for this row's readers)
        if (tid == 0) s_r0 = 0xFFFFFFFFu;
        __syncthreads();
        u32 w0 = 0, width = n;
        {
            ull b = 0;
#pragma unroll
            for (int i = 0; i < DN_WARPS; i++) { const ull g = s_gap[(row_par << 3) + i]; b = g > b ? g : b; }
            if (lenA) {
                const u32 gap = (u32)(b >> 32), kstart = (u32)b;
                const long long arc = (long long)n - gap + 1 + (p.cs_hi - p.cs_lo);      // columns the row's products can fall on
                if (arc < (long long)n) {
                    long long st = ((long long)kstart + p.cs_lo) % (long long)n; if (st < 0) st += n;
                    w0 = (u32)st; width = (u32)arc;
                }
            }
        }
        row_par ^= 1u;
        const u32 npieces = (width + p.wmax - 1) / p.wmax;

        // every product of the row whose window offset d lies in [lo, lo + wmax): f(d - lo, value)
        u32 P = 0;
        auto sweep = [&](u32 lo, bool numeric, bool countP) {
            auto one = [&](u32 c, u32 v) {
                u32 d = c - w0; if (c < w0) d += n;
                d -= lo;
                if (d < p.wmax) {
                    dn_red_or(sm_bm + (d >> 5) * 4u, 1u << (d & 31));
                    if (numeric) dn_red_add(sm_acc + d * 4u, v);
                }
            };
            auto entry = [&](const PackRec &rec, VT av) {
                const u32 len = rec.a.y, st = rec.a.x;
                if (!len) return;
                if (countP) P += len;
                const u32 c[B200_PACK_INLINE] = {rec.a.z, rec.a.w, rec.b.x, rec.b.y, rec.b.z, rec.b.w};
#pragma unroll
                for (int j = 0; j < B200_PACK_INLINE; j++) {
                    // unused slots repeat the record's last column (k_build_pack): marking twice is harmless, they add zero
                    u32 v = 0;
                    if ((u32)j < len) v = BPAT ? (u32)av : (u32)av * (u32)p.a.valB[st + j];
                    one(c[j], v);
                }
                for (u32 j = B200_PACK_INLINE; j < len; j++) one(p.a.colB[st + j], BPAT ? (u32)av : (u32)av * (u32)p.a.valB[st + j]);
            };
            if (tid < lenA) entry(rec0, a0);
            for (u32 t = DN_THREADS + tid; t < lenA; t += DN_THREADS) entry(dn_load_pack(p.pack, p.a.colA[rs + t]), p.a.valA[rs + t]);
        };
        // popcount of the window's bitmap: this thread's words [w_lo, w_hi), exclusive prefix over the CTA, total
        const u32 wpt = (bwords + DN_THREADS - 1) / DN_THREADS, w_lo = tid * wpt, w_hi = min(bwords, w_lo + wpt);
        auto popcount = [&](u32 &total) {
            u32 mine = 0;
            for (u32 w = w_lo; w < w_hi; w++) mine += __popc(bm[w]);
            return dn_block_scan(mine, s_warp, total);
        };

        u32 nnz = 0, pre = 0;
        const u32 split = w0 ? n - w0 : 0xFFFFFFFFu;                           // entries with d >= split wrapped below the origin: they come first in the row
        if (npieces == 1) {
            sweep(0, true, true);
            __syncthreads();
            pre = popcount(nnz);
            if (split < width) {                                               // r0 = entries with d < split (the owner of the split's word knows)
                const u32 sw = split >> 5;
                if (sw >= w_lo && sw < w_hi) {
                    u32 r = pre;
                    for (u32 w = w_lo; w < sw; w++) r += __popc(bm[w]);
                    s_r0 = r + __popc(bm[sw] & ((1u << (split & 31)) - 1u));
                }
            }
        } else {
            // wide row: lengths of all pieces first (bitmap only), then produce them one by one below
            for (u32 pc = 0; pc < npieces; pc++) {
                sweep(pc * p.wmax, false, pc == 0);
                __syncthreads();
                u32 t2; popcount(t2);
                nnz += t2;
                for (u32 w = w_lo; w < w_hi; w++) bm[w] = 0;
                __syncthreads();
            }
        }
        { const u32 Pw = warp_sum_u32(P); if (lane == 0 && Pw) atomicAdd(&s_P, Pw); }
        // the next row's records can go out now: their columns were requested at the top of the iteration and the emit phase
        // below covers the records' own latency
        if (row + G < p.rows && tid < len1) rec1 = dn_load_pack(p.pack, k1);
        // ---- place the row: publish its length, sum the lengths of all rows before it (decoupled look-back over rows)
        if (wid == 0) {
            if (lane == 0) atomicExch((ull *)&p.status[row], (ull)((row == 0 ? SCAN_FLAG_PRE : SCAN_FLAG_AGG) | (u64)nnz));
            u64 excl = 0;
            if (row > 0 && !p.nolook) {
                long long look = (long long)row - 1;
                while (true) {
                    const long long idx = look - lane;
                    u64 st;
                    do { st = idx >= 0 ? ld_volatile_u64(&p.status[idx]) : SCAN_FLAG_PRE; } while (__any_sync(0xFFFFFFFFu, (st >> 62) == 0));
                    const u32 pre_mask = __ballot_sync(0xFFFFFFFFu, (st >> 62) == 2);
                    const int first = pre_mask ? __ffs(pre_mask) - 1 : 32;
                    excl += warp_sum_u64((int)lane <= first ? (st & SCAN_VAL_MASK) : 0ull);
                    if (pre_mask) break;
                    look -= 32;
                }
                if (lane == 0) atomicExch((ull *)&p.status[row], (ull)(SCAN_FLAG_PRE | (excl + nnz)));
            }
            if (lane == 0) {
                s_excl = excl;
                p.rpC[row] = excl;
                if (row == p.rows - 1) { p.rpC[p.rows] = excl + nnz; p.ctrl->total_nnz = excl + nnz; }
            }
        }
        __syncthreads();
        const u64 obase = s_excl;
        if (tid == 0) { const u32 Prow = s_P; s_P = 0; psum += Prow; pmax = max(pmax, Prow); nmax = max(nmax, nnz); }
        if (npieces == 1) {
            const u32 r0 = s_r0 == 0xFFFFFFFFu ? nnz : s_r0, hi_cnt = nnz - r0;
            u32 g = pre;
            for (u32 w = w_lo; w < w_hi; w++) {
                u32 wd = bm[w];
                if (!wd) continue;
                bm[w] = 0;
                while (wd) {
                    const u32 b = __ffs(wd) - 1; wd &= wd - 1;
                    const u32 d = (w << 5) + b;
                    const u32 v = acc[d];
                    acc[d] = 0;
                    u32 c = w0 + d; if (c >= n) c -= n;
                    const u64 q = obase + (g >= r0 ? g - r0 : g + hi_cnt);
                    p.colC[q] = c; p.valC[q] = (VT)v;
                    vmax = vmax > (u64)v ? vmax : (u64)v;
                    g++;
                }
            }
            __syncthreads();
        } else {
            // pieces in d order; r0 needs the count below `split` first: one more bitmap-only pass over the pieces before it
            u32 r0 = nnz;
            if (split < width) {
                r0 = 0;
                for (u32 pc = 0; pc * p.wmax < split; pc++) {
                    sweep(pc * p.wmax, false, false);
                    __syncthreads();
                    const u32 lim = min(p.wmax, split - pc * p.wmax);        // window offsets below the split inside this piece
                    u32 mine = 0;
                    for (u32 w = w_lo; w < w_hi; w++) {
                        const u32 bit0 = w << 5;
                        u32 wd = bm[w];
                        if (bit0 >= lim) wd = 0; else if (bit0 + 32 > lim) wd &= (1u << (lim - bit0)) - 1u;
                        mine += __popc(wd);
                        bm[w] = 0;
                    }
                    u32 t2; dn_block_scan(mine, s_warp, t2);
                    r0 += t2;
                    __syncthreads();
                }
            }
            const u32 hi_cnt = nnz - r0;
            u32 done = 0;
            for (u32 pc = 0; pc < npieces; pc++) {
                sweep(pc * p.wmax, true, false);
                __syncthreads();
                u32 t2;
                u32 g = done + popcount(t2);
                for (u32 w = w_lo; w < w_hi; w++) {
                    u32 wd = bm[w];
                    if (!wd) continue;
                    bm[w] = 0;
                    while (wd) {
                        const u32 b = __ffs(wd) - 1; wd &= wd - 1;
                        const u32 dl = (w << 5) + b;
                        const u32 v = acc[dl];
                        acc[dl] = 0;
                        u32 c = w0 + pc * p.wmax + dl; if (c >= n) c -= n;
                        const u64 q = obase + (g >= r0 ? g - r0 : g + hi_cnt);
                        p.colC[q] = c; p.valC[q] = (VT)v;
                        vmax = vmax > (u64)v ? vmax : (u64)v;
                        g++;
                    }
                }
                done += t2;
                __syncthreads();
            }
        }
    }
    // ---- totals, then the last CTA reports the control block to the pinned ring and leaves it zeroed
    {
        if (tid == 0) {
            if (psum) atomicAdd(&p.ctrl->total_products, (ull)psum);
            atomicMax(&p.ctrl->max_row_products, (ull)pmax);
            atomicMax(&p.ctrl->max_row_nnz, (ull)nmax);
        }
        vmax = warp_max_u64(vmax);
        if (lane == 0 && vmax) atomicMax(&p.ctrl->max_val_out, (ull)vmax);
    }
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
struct DnKernel { const void *fn; size_t static_smem; };
static DnKernel g_dn[2][2];     // [value width][pattern-only B]
template <typename VT, bool BPAT>
static void dn_register(DnKernel &k, size_t optin) {
    k.fn = (const void *)k_dense<VT, BPAT>;
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, k.fn) == cudaSuccess) k.static_smem = fa.sharedSizeBytes; else { cudaGetLastError(); k.static_smem = 256; }
    if (cudaFuncSetAttribute(k.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(optin - k.static_smem)) != cudaSuccess) cudaGetLastError();
}
void dn_setup(b200_ctx *ctx) {
    const size_t o = ctx->smem_optin;
    dn_register<u32, false>(g_dn[0][0], o); dn_register<u32, true>(g_dn[0][1], o);
    dn_register<u64, false>(g_dn[1][0], o); dn_register<u64, true>(g_dn[1][1], o);
}

// window columns per CTA for `ctas` CTAs per SM
u32 dn_window_cols(const b200_ctx *ctx, int ctas) {
    const size_t per = (size_t)(228 * 1024) / ctas - 1024 - 512;
    size_t w = per * 8 / 33;                                               // 4 bytes + 1 bit per column
    w = w / 1024 * 1024;
    return (u32)std::min<size_t>(w, 48 * 1024);
}

template <typename VT>
static cudaError_t dn_go(b200_ctx *ctx, const b200_csr *A, const b200_csr *B, b200_csr *C, B200Ctrl *ctrl, u32 wmax, u64 *mirror, u32 epoch,
                         const void *fn, int grid, size_t smem, cudaStream_t s) {
    DnArgs<VT> p;
    p.a = NumArgs<VT>{A->d_rp, A->d_col, (const VT *)A->d_val, B->d_desc, B->d_col, (const VT *)B->d_val};
    p.nolook = getenv("B200_DN_NOLOOK") ? 1u : 0u;
    p.pack = B->d_pack; p.rows = A->rows; p.ncols = (u32)B->cols; p.cs_lo = B->cs_lo; p.cs_hi = B->cs_hi; p.wmax = wmax;
    p.status = ctx->d_rowstat; p.rpC = C->d_rp; p.colC = C->d_col; p.valC = (VT *)C->d_val;
    p.ctrl = ctrl; p.host_mirror = mirror; p.epoch = epoch; p.maxval_dst = C->d_maxval;
    void *kargs[] = {(void *)&p};
    return cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(DN_THREADS), kargs, smem, s);
}

// The whole multiply in one launch; C's arrays are already allocated (from the host-known bound), ctrl is zeroed.
int dn_launch(b200_ctx *ctx, const b200_csr *A, const b200_csr *B, b200_csr *C, B200Ctrl *ctrl, bool bpat, int ctas_per_sm, u64 *mirror, u32 epoch,
              cudaStream_t s) {
    const bool v64 = A->val_bits == 64;
    const DnKernel &k = g_dn[v64 ? 1 : 0][bpat ? 1 : 0];
    if (!k.fn) return set_err(B200_ERR_CUDA, "dense one-pass kernel variant is not registered");
    if (A->rows > ctx->cap_rowstat) {
        if (ctx->d_rowstat) cudaFree(ctx->d_rowstat);
        ctx->d_rowstat = nullptr; ctx->cap_rowstat = 0;
        const u64 cap = A->rows + A->rows / 8 + 1024;
        CUDA_TRY(cudaMalloc((void **)&ctx->d_rowstat, cap * 8));
        ctx->cap_rowstat = cap;
    }
    CUDA_TRY(cudaMemsetAsync(ctx->d_rowstat, 0, A->rows * 8, s));
    const u32 wmax = dn_window_cols(ctx, ctas_per_sm);
    const size_t smem = (size_t)wmax * 4 + wmax / 8;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k.fn, DN_THREADS, smem) != cudaSuccess || per_sm < 1) { cudaGetLastError(); return set_err(B200_ERR_CUDA, "dense one-pass kernel does not fit an SM"); }
    const int grid = (int)std::max<u64>(1, std::min<u64>(A->rows, (u64)ctx->num_sms * per_sm));
    const cudaError_t le = v64 ? dn_go<u64>(ctx, A, B, C, ctrl, wmax, mirror, epoch, k.fn, grid, smem, s)
                               : dn_go<u32>(ctx, A, B, C, ctrl, wmax, mirror, epoch, k.fn, grid, smem, s);
    ctx->launches++;
    if (ctx->trace) { fprintf(stderr, "[b200 trace] dense one-pass: grid %d (%d/SM) smem %zu window %u columns\n", grid, per_sm, smem, wmax); trace_mark(ctx, __LINE__); }
    if (le != cudaSuccess) return set_err(B200_ERR_CUDA, "dense one-pass kernel launch failed: %s", cudaGetErrorString(le));
    return B200_OK;
}
