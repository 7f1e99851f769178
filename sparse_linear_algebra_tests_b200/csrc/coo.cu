// coo.cu -- COO -> CSR on the device, and the R-MAT generator that feeds it.
//
//   b200_csr_from_coo[_device]   CsrMatrix::from_coo (/root/reference/src/graph_csr.rs:83-129; saturating flavour:
//                                linalg/src/csr.rs:158-195; MagnusMatrix::from_coo, src/graph_magnus.rs:34-76): order the
//                                triplets by (row, column), sum duplicates, drop zeros, build row_ptr.
//   b200_rmat                    recursive-quadrant generator with a counter-based splitmix64 (App. C): edge e, level l
//                                uses draw number e * scale + l, so every edge is independent and the device builds the
//                                16 M-node input of BASELINE configs[3] without a host pass.
//
// The ordering is a hand-written least-significant-digit radix sort on the packed key row << bits(cols) | column
// (8 bits a pass, only as many passes as the key has bits): per pass a tile histogram, an exclusive scan of the
// digit-major table, and a stable scatter (warp-private digit counters fed by match.any, so equal digits keep their
// order).  Duplicate sums do not depend on the order inside a run (wrapping and saturating addition of unsigned values
// are both commutative and associative), so the unstable sort of the reference gives the same matrix.
#include "engine.cuh"
#include "devutil.cuh"

#define RS_THREADS 256
#define RS_ITEMS 16
#define RS_TILE (RS_THREADS * RS_ITEMS)
#define RS_WARPS (RS_THREADS / 32)

// ---------------------------------------------------------------------------- exclusive scan of u32 (in place)
__global__ void __launch_bounds__(RS_THREADS) k_sc_tile_sums(const u32 *__restrict__ in, u64 n, u32 *__restrict__ sums) {
    __shared__ u32 s_w[RS_WARPS];
    const u64 base = (u64)blockIdx.x * RS_TILE;
    u32 v = 0;
#pragma unroll
    for (int j = 0; j < RS_ITEMS; j++) { const u64 i = base + (u64)j * RS_THREADS + threadIdx.x; if (i < n) v += in[i]; }
    v = warp_sum_u32(v);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) { u32 t = 0; for (int w = 0; w < RS_WARPS; w++) t += s_w[w]; sums[blockIdx.x] = t; }
}
// one CTA: data[0..n) -> exclusive prefix, data[n] = total
__global__ void __launch_bounds__(1024) k_sc_single(u32 *data, u32 n) {
    __shared__ u32 s_w[33];
    __shared__ u32 s_run;
    const u32 tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) s_run = 0;
    __syncthreads();
    for (u32 base = 0; base < n; base += 1024) {
        const u32 i = base + tid;
        const u32 v = i < n ? data[i] : 0u;
        u32 incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const u32 t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
        if (lane == 31) s_w[w] = incl;
        __syncthreads();
        if (w == 0) {
            u32 x = s_w[lane], xi = x;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const u32 t = __shfl_up_sync(0xFFFFFFFFu, xi, d); if (lane >= d) xi += t; }
            s_w[lane] = xi - x;
            if (lane == 31) s_w[32] = xi;
        }
        __syncthreads();
        const u32 run = s_run;
        if (i < n) data[i] = run + s_w[w] + incl - v;
        __syncthreads();
        if (tid == 0) s_run = run + s_w[32];
        __syncthreads();
    }
    if (tid == 0) data[n] = s_run;
}
// in-place: every tile scans itself (items are strided over the threads, so the order inside a tile is item-major)
__global__ void __launch_bounds__(RS_THREADS) k_sc_apply(u32 *data, u64 n, const u32 *__restrict__ tile_base) {
    __shared__ u32 s_v[RS_TILE];
    __shared__ u32 s_w[RS_WARPS + 1];
    const u64 base = (u64)blockIdx.x * RS_TILE;
    const u32 tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (int j = 0; j < RS_ITEMS; j++) { const u32 l = j * RS_THREADS + tid; s_v[l] = base + l < n ? data[base + l] : 0u; }
    __syncthreads();
    // thread t owns the RS_ITEMS consecutive items t * RS_ITEMS ..
    u32 loc[RS_ITEMS], sum = 0;
#pragma unroll
    for (int j = 0; j < RS_ITEMS; j++) { loc[j] = sum; sum += s_v[tid * RS_ITEMS + j]; }
    u32 incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const u32 t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
    if (lane == 31) s_w[w] = incl;
    __syncthreads();
    u32 wb = 0;
    for (u32 i = 0; i < w; i++) wb += s_w[i];
    const u32 tb = tile_base[blockIdx.x] + wb + incl - sum;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < RS_ITEMS; j++) s_v[tid * RS_ITEMS + j] = tb + loc[j];
    __syncthreads();
    for (int j = 0; j < RS_ITEMS; j++) { const u32 l = j * RS_THREADS + tid; if (base + l < n) data[base + l] = s_v[l]; }
}

// exclusive scan of d[0..n) in place, d[n] = total.  `tmp` holds tiles(n) + tiles(tiles(n)) + 4 words.  n < 2^32.
static int scan_u32(b200_ctx *ctx, u32 *d, u64 n, u32 *tmp, cudaStream_t s) {
    if (n <= (u64)RS_TILE * 4) { k_sc_single<<<1, 1024, 0, s>>>(d, (u32)n); LAUNCH_CHECK(ctx); return B200_OK; }
    const u64 t1 = (n + RS_TILE - 1) / RS_TILE;
    k_sc_tile_sums<<<(unsigned)t1, RS_THREADS, 0, s>>>(d, n, tmp);
    LAUNCH_CHECK(ctx);
    TRY(scan_u32(ctx, tmp, t1, tmp + t1 + 1, s));                       // tmp[t1] = total
    k_sc_apply<<<(unsigned)t1, RS_THREADS, 0, s>>>(d, n, tmp);
    LAUNCH_CHECK(ctx);
    CUDA_TRY(cudaMemcpyAsync(d + n, tmp + t1, 4, cudaMemcpyDeviceToDevice, s));
    return B200_OK;
}
static u64 scan_tmp_words(u64 n) { u64 w = 8; while (n > (u64)RS_TILE * 4) { n = (n + RS_TILE - 1) / RS_TILE; w += n + 2; } return w + 8; }

// ---------------------------------------------------------------------------- radix sort passes
__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const u64 *__restrict__ keys, u64 n, int shift, u32 *__restrict__ hist, u32 nblocks) {
    __shared__ u32 h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const u64 base = (u64)blockIdx.x * RS_TILE;
#pragma unroll
    for (int j = 0; j < RS_ITEMS; j++) {
        const u64 i = base + (u64)j * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(u32)(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(u64)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

// offs: exclusive scan of the digit-major histogram table.  Stable: a warp owns RS_TILE / RS_WARPS consecutive elements
// and walks them 32 at a time; lanes with the same digit (match.any) take consecutive ranks behind the warp's running count.
template <typename VT>
__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const u64 *__restrict__ kin, const VT *__restrict__ vin, u64 *__restrict__ kout,
                                                           VT *__restrict__ vout, u64 n, int shift, const u32 *__restrict__ offs, u32 nblocks) {
    __shared__ u32 wh[RS_WARPS][256];
    const u32 tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (u32 t = tid; t < RS_WARPS * 256; t += RS_THREADS) (&wh[0][0])[t] = 0;
    __syncthreads();
    const u64 wbase = (u64)blockIdx.x * RS_TILE + (u64)w * (RS_TILE / RS_WARPS);
    u64 key[RS_ITEMS]; u32 rank[RS_ITEMS];
#pragma unroll
    for (int j = 0; j < RS_ITEMS; j++) {
        const u64 i = wbase + (u64)j * 32 + lane;
        const bool ok = i < n;
        key[j] = ok ? kin[i] : 0ull;
        const u32 d = ok ? (u32)(key[j] >> shift) & 255u : 256u;           // (out-of-range lanes share a digit nobody counts)
        const u32 peers = __match_any_sync(0xFFFFFFFFu, d);
        const u32 before = __popc(peers & ((1u << lane) - 1u));
        const int leader = __ffs(peers) - 1;
        u32 basec = 0;
        if (ok && (int)lane == leader) { basec = wh[w][d]; wh[w][d] = basec + __popc(peers); }
        basec = __shfl_sync(0xFFFFFFFFu, basec, leader);
        rank[j] = basec + before;
        __syncwarp();
    }
    __syncthreads();
    {   // digit d: where each warp's elements start
        u32 run = offs[(u64)tid * nblocks + blockIdx.x];
#pragma unroll
        for (int i = 0; i < RS_WARPS; i++) { const u32 t = wh[i][tid]; wh[i][tid] = run; run += t; }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < RS_ITEMS; j++) {
        const u64 i = wbase + (u64)j * 32 + lane;
        if (i < n) {
            const u32 d = (u32)(key[j] >> shift) & 255u;
            const u32 pos = wh[w][d] + rank[j];
            kout[pos] = key[j]; vout[pos] = vin[i];
        }
    }
}

// ---------------------------------------------------------------------------- triplets -> keys, runs -> entries
template <typename VT>
__global__ void __launch_bounds__(256) k_coo_pack(u64 n, const u32 *__restrict__ r, const u32 *__restrict__ c, u64 rows, u64 cols, int cbits,
                                                  u64 *__restrict__ keys, u32 *bad) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const u32 rr = r[i], cc = c[i];
        if (rr >= rows || cc >= cols) atomicOr(bad, 1u);
        keys[i] = ((u64)rr << cbits) | cc;
    }
}
// first element of every run of equal keys: the run's sum (wrapping `+=`, or saturating), kept unless it is zero
template <typename VT>
__global__ void __launch_bounds__(256) k_coo_runs(u64 n, const u64 *__restrict__ keys, const VT *__restrict__ vals, int saturating,
                                                  VT *__restrict__ sums, u32 *__restrict__ keep) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const u64 k = keys[i];
        u32 kp = 0;
        if (i == 0 || keys[i - 1] != k) {
            VT sum = vals[i];
            for (u64 t = i + 1; t < n && keys[t] == k; t++) sum = saturating ? sat_add(sum, vals[t]) : (VT)(sum + vals[t]);
            sums[i] = sum;
            kp = sum != 0;
        }
        keep[i] = kp;
    }
}
template <typename VT>
__global__ void __launch_bounds__(256) k_coo_emit(u64 n, const u64 *__restrict__ keys, const VT *__restrict__ sums, const u32 *__restrict__ pos, int cbits,
                                                  u32 *__restrict__ col, VT *__restrict__ val, u32 *__restrict__ row_cnt) {
    const u64 cmask = cbits ? (1ull << cbits) - 1ull : 0ull;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const u32 p = pos[i];
        if (pos[i + 1] != p) {
            const u64 k = keys[i];
            col[p] = (u32)(k & cmask); val[p] = sums[i];
            atomicAdd(&row_cnt[(u32)(k >> cbits)], 1u);
        }
    }
}

// R-MAT edges (host twin: hostgen.rmat): quadrant of edge e at level l from draw e * scale + l of
// splitmix64(seed): u < a -> (0,0); < a+b -> (0,1); < a+b+c -> (1,0); else (1,1); the level sets bit l of row / column.
__device__ __forceinline__ u64 splitmix64_at(u64 seed, u64 idx) {
    u64 z = seed + (idx + 1ull) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
template <typename VT>
__global__ void __launch_bounds__(256) k_rmat_edges(u64 m, int scale, double a, double ab, double abc, u64 seed, u64 *__restrict__ keys, VT *__restrict__ vals) {
    for (u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x; e < m; e += (u64)gridDim.x * blockDim.x) {
        u64 r = 0, c = 0;
        for (int l = 0; l < scale; l++) {
            const double u = (double)(splitmix64_at(seed, e * (u64)scale + (u64)l) >> 11) * (1.0 / 9007199254740992.0);
            c |= (u64)(((u >= a) && (u < ab)) || (u >= abc)) << l;
            r |= (u64)(u >= ab) << l;
        }
        keys[e] = (r << scale) | c;
        vals[e] = (VT)1;
    }
}

// ---------------------------------------------------------------------------- host side
static int bits_for(u64 n) { int b = 0; while (b < 64 && (n - 1) >> b) b++; return n <= 1 ? 0 : b; }

// keys / vals: n packed triplets in device buffers of n entries each (consumed); alt buffers of the same size.
template <typename VT>
static int coo_build(b200_ctx *ctx, u64 rows, u64 cols, u64 n, u64 *keys, VT *vals, u64 *keys2, VT *vals2, int rbits, int cbits,
                     int saturating, b200_csr **out) {
    cudaStream_t s = ctx->stream;
    if (n == 0 || rows == 0) {                                             // no triplets: the empty matrix
        b200_csr *E = nullptr;
        TRY(csr_alloc(ctx, rows, cols, 0, (int)sizeof(VT) * 8, true, &E));
        cudaError_t e = cudaMemsetAsync(E->d_rp, 0, (rows + 1) * 8 + 16, s);
        if (e != cudaSuccess) { b200_csr_free(ctx, E); return set_err(B200_ERR_CUDA, "from_coo: %s", cudaGetErrorString(e)); }
        E->h_maxval = 0; E->h_maxval_known = true;
        *out = E;
        return B200_OK;
    }
    const int g = (int)std::max<u64>(1, std::min<u64>((n + 255) / 256, (u64)ctx->num_sms * 16));
    const u32 nblocks = (u32)((n + RS_TILE - 1) / RS_TILE);
    const u64 table = (u64)256 * nblocks;
    u32 *hist = nullptr, *tmp = nullptr, *keep = nullptr;
    const u64 tmp_words = std::max(scan_tmp_words(table), scan_tmp_words(n + 1));
    TRY(dmalloc(ctx, (void **)&hist, (table + 2) * 4));
    int r = dmalloc(ctx, (void **)&tmp, tmp_words * 4);
    if (r == B200_OK) r = dmalloc(ctx, (void **)&keep, (n + 2) * 4);
    auto cleanup = [&]() { dfree(ctx, hist); dfree(ctx, tmp); dfree(ctx, keep); };
    if (r != B200_OK) { cleanup(); return r; }
    // ---- order by (row, column): as many 8-bit passes as the packed key has bits
    const int kbits = rbits + cbits;
    for (int shift = 0; shift < kbits && n > 1; shift += 8) {
        k_rs_hist<<<nblocks, RS_THREADS, 0, s>>>(keys, n, shift, hist, nblocks);
        ctx->launches++;
        r = scan_u32(ctx, hist, table, tmp, s);
        if (r != B200_OK) { cleanup(); return r; }
        k_rs_scatter<VT><<<nblocks, RS_THREADS, 0, s>>>(keys, vals, keys2, vals2, n, shift, hist, nblocks);
        ctx->launches++;
        std::swap(keys, keys2); std::swap(vals, vals2);
    }
    // ---- runs of equal (row, column): sum, drop zeros, place
    k_coo_runs<VT><<<g, 256, 0, s>>>(n, keys, vals, saturating, vals2, keep);
    ctx->launches++;
    r = scan_u32(ctx, keep, n, tmp, s);
    u32 nnz32 = 0;
    if (r == B200_OK && cudaMemcpyAsync(&nnz32, keep + n, 4, cudaMemcpyDeviceToHost, s) != cudaSuccess) r = set_err(B200_ERR_CUDA, "from_coo: reading the entry count failed");
    if (r == B200_OK && cudaStreamSynchronize(s) != cudaSuccess) r = set_err(B200_ERR_CUDA, "from_coo: %s", cudaGetErrorString(cudaGetLastError()));
    if (r != B200_OK) { cleanup(); return r; }
    b200_csr *C = nullptr;
    r = csr_alloc(ctx, rows, cols, nnz32, (int)sizeof(VT) * 8, true, &C);
    if (r == B200_OK) r = ensure_row_scratch(ctx, rows);
    if (r != B200_OK) { if (C) b200_csr_free(ctx, C); cleanup(); return r; }
    cudaMemsetAsync(ctx->d_nnz_row, 0, rows * 4, s);
    k_coo_emit<VT><<<g, 256, 0, s>>>(n, keys, vals2, keep, cbits, C->d_col, (VT *)C->d_val, ctx->d_nnz_row);
    ctx->launches++;
    u64 total = 0, maxlen = 0;
    r = scan_row_counts(ctx, rows, C->d_rp, &total, &maxlen);
    cleanup();
    if (r == B200_OK && total != nnz32) r = set_err(B200_ERR_CUDA, "from_coo: internal count mismatch (%llu vs %u)", (ull)total, nnz32);
    if (r == B200_OK) { C->max_row_len = maxlen; r = finish_new_csr(ctx, C, true, false); }
    if (r != B200_OK) { b200_csr_free(ctx, C); return r; }
    *out = C;
    return B200_OK;
}

template <typename VT>
static int from_coo_t(b200_ctx *ctx, u64 rows, u64 cols, u64 n, const u32 *r_in, const u32 *c_in, const void *v_in, bool host, int saturating,
                      b200_csr **out) {
    cudaStream_t s = ctx->stream;
    const int rbits = bits_for(rows), cbits = bits_for(cols);
    u64 *keys = nullptr, *keys2 = nullptr; VT *vals = nullptr, *vals2 = nullptr; u32 *dr = nullptr, *dc = nullptr;
    const u64 cap = std::max<u64>(n, 1);
    auto cleanup = [&]() { dfree(ctx, keys); dfree(ctx, keys2); dfree(ctx, vals); dfree(ctx, vals2); if (host) { dfree(ctx, dr); dfree(ctx, dc); } };
    int r = dmalloc(ctx, (void **)&keys, cap * 8);
    if (r == B200_OK) r = dmalloc(ctx, (void **)&keys2, cap * 8);
    if (r == B200_OK) r = dmalloc(ctx, (void **)&vals, cap * sizeof(VT));
    if (r == B200_OK) r = dmalloc(ctx, (void **)&vals2, cap * sizeof(VT));
    if (r == B200_OK && host && n) {
        r = dmalloc(ctx, (void **)&dr, n * 4);
        if (r == B200_OK) r = dmalloc(ctx, (void **)&dc, n * 4);
        if (r == B200_OK && (cudaMemcpyAsync(dr, r_in, n * 4, cudaMemcpyHostToDevice, s) != cudaSuccess || cudaMemcpyAsync(dc, c_in, n * 4, cudaMemcpyHostToDevice, s) != cudaSuccess ||
                             cudaMemcpyAsync(vals, v_in, n * sizeof(VT), cudaMemcpyHostToDevice, s) != cudaSuccess))
            r = set_err(B200_ERR_CUDA, "from_coo: uploading the triplets failed");
    } else if (r == B200_OK && n) {
        dr = const_cast<u32 *>(r_in); dc = const_cast<u32 *>(c_in);
        if (cudaMemcpyAsync(vals, v_in, n * sizeof(VT), cudaMemcpyDeviceToDevice, s) != cudaSuccess) r = set_err(B200_ERR_CUDA, "from_coo: copying the values failed");
    }
    if (r == B200_OK && n) {
        cudaMemsetAsync(ctx->d_flag, 0, 4, s);
        const int g = (int)std::max<u64>(1, std::min<u64>((n + 255) / 256, (u64)ctx->num_sms * 16));
        k_coo_pack<VT><<<g, 256, 0, s>>>(n, dr, dc, rows, cols, cbits, keys, ctx->d_flag);
        ctx->launches++;
        if (cudaMemcpyAsync(ctx->h_flag, ctx->d_flag, 4, cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess)
            r = set_err(B200_ERR_CUDA, "from_coo: %s", cudaGetErrorString(cudaGetLastError()));
        else if (ctx->h_flag[0]) r = set_err(B200_ERR_BADARG, "from_coo: triplet index out of range");
    }
    if (r == B200_OK) r = coo_build<VT>(ctx, rows, cols, n, keys, vals, keys2, vals2, rbits, cbits, saturating, out);
    cleanup();
    return r;
}

static int from_coo_common(b200_ctx *ctx, uint64_t rows, uint64_t cols, uint64_t n, const uint32_t *r, const uint32_t *c, const void *v,
                           int val_bits, int saturating, bool host, b200_csr **out) {
    if (!ctx || !out || (n && (!r || !c || !v))) return set_err(B200_ERR_BADARG, "NULL argument");
    if (val_bits != 32 && val_bits != 64) return set_err(B200_ERR_BADARG, "val_bits must be 32 or 64");
    if (rows > 0xFFFFFFFFull || cols > 0xFFFFFFFFull) return set_err(B200_ERR_BADARG, "from_coo: more than 2^32-1 rows or columns (NodeId is u32)");
    if (n >= 0xFFFFFFF0ull) return set_err(B200_ERR_BADARG, "from_coo: at most 2^32-17 triplets a call");
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (val_bits == 32) return from_coo_t<u32>(ctx, rows, cols, n, r, c, v, host, saturating, out);
    return from_coo_t<u64>(ctx, rows, cols, n, r, c, v, host, saturating, out);
}
extern "C" int b200_csr_from_coo(b200_ctx *ctx, uint64_t rows, uint64_t cols, uint64_t n, const uint32_t *r, const uint32_t *c, const void *v,
                                 int val_bits, int saturating, b200_csr **out) {
    return from_coo_common(ctx, rows, cols, n, r, c, v, val_bits, saturating, true, out);
}
extern "C" int b200_csr_from_coo_device(b200_ctx *ctx, uint64_t rows, uint64_t cols, uint64_t n, const uint32_t *d_r, const uint32_t *d_c, const void *d_v,
                                        int val_bits, int saturating, b200_csr **out) {
    return from_coo_common(ctx, rows, cols, n, d_r, d_c, d_v, val_bits, saturating, false, out);
}

template <typename VT>
static int rmat_t(b200_ctx *ctx, int scale, u64 m, double a, double b, double c, u64 seed, b200_csr **out) {
    cudaStream_t s = ctx->stream;
    u64 *keys = nullptr, *keys2 = nullptr; VT *vals = nullptr, *vals2 = nullptr;
    auto cleanup = [&]() { dfree(ctx, keys); dfree(ctx, keys2); dfree(ctx, vals); dfree(ctx, vals2); };
    int r = dmalloc(ctx, (void **)&keys, m * 8);
    if (r == B200_OK) r = dmalloc(ctx, (void **)&keys2, m * 8);
    if (r == B200_OK) r = dmalloc(ctx, (void **)&vals, m * sizeof(VT));
    if (r == B200_OK) r = dmalloc(ctx, (void **)&vals2, m * sizeof(VT));
    if (r != B200_OK) { cleanup(); return r; }
    const double ab = a + b, abc = ab + c;                                   // (the host twin adds in the same order)
    const int g = (int)std::max<u64>(1, std::min<u64>((m + 255) / 256, (u64)ctx->num_sms * 16));
    k_rmat_edges<VT><<<g, 256, 0, s>>>(m, scale, a, ab, abc, seed, keys, vals);
    ctx->launches++;
    const u64 n = 1ull << scale;
    r = coo_build<VT>(ctx, n, n, m, keys, vals, keys2, vals2, scale, scale, 0, out);
    cleanup();
    return r;
}
extern "C" int b200_rmat(b200_ctx *ctx, int scale, uint64_t edge_factor, double a, double b, double c, uint64_t seed, int val_bits, b200_csr **out) {
    if (!ctx || !out) return set_err(B200_ERR_BADARG, "NULL argument");
    if (scale < 1 || scale > 31) return set_err(B200_ERR_BADARG, "rmat: scale must be 1..31 (NodeId is u32)");
    if (val_bits != 32 && val_bits != 64) return set_err(B200_ERR_BADARG, "val_bits must be 32 or 64");
    if (!(a >= 0 && b >= 0 && c >= 0 && a + b + c <= 1.0)) return set_err(B200_ERR_BADARG, "rmat: quadrant probabilities must be >= 0 and sum to at most 1");
    const unsigned __int128 m128 = (unsigned __int128)edge_factor << scale;
    if (m128 == 0 || m128 >= 0xFFFFFFF0ull) return set_err(B200_ERR_BADARG, "rmat: edge_factor * 2^scale must be in 1..2^32-17");
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (val_bits == 32) return rmat_t<u32>(ctx, scale, (u64)m128, a, b, c, seed, out);
    return rmat_t<u64>(ctx, scale, (u64)m128, a, b, c, seed, out);
}

// =======================================================================================
// Locality pre-pass: CsrMatrix::rcm / permute / bandwidth_stats, /root/reference/src/graph_csr.rs:663-818.
//   permute          on the device: every entry is relabelled (inv[row], inv[col]) and the triplets go through the same
//                    radix-sort assembly as from_coo (a permutation creates no duplicates), so rows come out sorted;
//   bandwidth_stats  one reduction kernel (max and sum of |r - c|);
//   rcm_order        the Cuthill-McKee queue is sequential by definition (a node's place depends on every node queued
//                    before it); the ordering runs on the host inside the library, on a downloaded pattern, exactly as
//                    the reference's loop does -- applying it (permute) and measuring it (bandwidth) are device work.
// =======================================================================================
template <typename VT>
__global__ void __launch_bounds__(256) k_perm_keys(u64 rows, const u64 *__restrict__ rp, const u32 *__restrict__ col, const u32 *__restrict__ inv_r,
                                                   const u32 *__restrict__ inv_c, int cbits, u64 *__restrict__ keys) {
    // a warp per row: entries of a row are contiguous
    const u32 lane = threadIdx.x & 31;
    const u64 wpb = blockDim.x >> 5;
    for (u64 r = (u64)blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += (u64)gridDim.x * wpb) {
        const u64 s = rp[r], e = rp[r + 1];
        const u64 hi = (u64)inv_r[r] << cbits;
        for (u64 i = s + lane; i < e; i += 32) keys[i] = hi | inv_c[col[i]];
    }
}
__global__ void __launch_bounds__(256) k_invert_perm(u64 n, const u32 *__restrict__ perm, u32 *__restrict__ inv, u32 *bad) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const u32 o = perm[i];
        if (o >= n) { atomicOr(bad, 1u); continue; }
        if (atomicExch(&inv[o], (u32)i) != 0xFFFFFFFFu) atomicOr(bad, 2u);     // an old index named twice: not a permutation
    }
}
__global__ void __launch_bounds__(256) k_bandwidth(u64 rows, const u64 *__restrict__ rp, const u32 *__restrict__ col, ull *out /* max, sum */) {
    const u32 lane = threadIdx.x & 31;
    const u64 wpb = blockDim.x >> 5;
    u64 mx = 0, sum = 0;
    for (u64 r = (u64)blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += (u64)gridDim.x * wpb) {
        const u64 s = rp[r], e = rp[r + 1];
        for (u64 i = s + lane; i < e; i += 32) { const u64 c = col[i], d = r > c ? r - c : c - r; mx = mx > d ? mx : d; sum += d; }
    }
    mx = warp_max_u64(mx); sum = warp_sum_u64(sum);
    if (lane == 0) { if (mx) atomicMax(&out[0], (ull)mx); if (sum) atomicAdd(&out[1], (ull)sum); }
}

extern "C" int b200_csr_bandwidth_stats(b200_ctx *ctx, const b200_csr *m, uint64_t *max_bw, double *avg_bw) {
    if (!ctx || !m || !max_bw || !avg_bw) return set_err(B200_ERR_BADARG, "NULL argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    RESOLVE(ctx, m);
    cudaStream_t s = ctx->stream;
    ull *d = reinterpret_cast<ull *>(ctx->d_flag + 16);                    // 16 bytes of the context's flag words (8-byte aligned)
    CUDA_TRY(cudaMemsetAsync(d, 0, 16, s));
    if (m->rows && m->nnz) {
        const int g = (int)std::max<u64>(1, std::min<u64>((m->rows + 7) / 8, (u64)ctx->num_sms * 16));
        k_bandwidth<<<g, 256, 0, s>>>(m->rows, m->d_rp, m->d_col, d);
        LAUNCH_CHECK(ctx);
    }
    ull h[2] = {0, 0};
    CUDA_TRY(cudaMemcpyAsync(h, d, 16, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    *max_bw = h[0];
    *avg_bw = (double)h[1] / (double)(m->nnz ? m->nnz : 1);
    return B200_OK;
}

template <typename VT>
static int permute_t(b200_ctx *ctx, const b200_csr *A, const u32 *d_inv, b200_csr **out) {
    cudaStream_t s = ctx->stream;
    const u64 n = A->nnz, cap = std::max<u64>(n, 1);
    const int bits = bits_for(A->rows);
    u64 *keys = nullptr, *keys2 = nullptr; VT *vals = nullptr, *vals2 = nullptr;
    auto cleanup = [&]() { dfree(ctx, keys); dfree(ctx, keys2); dfree(ctx, vals); dfree(ctx, vals2); };
    int r = dmalloc(ctx, (void **)&keys, cap * 8);
    if (r == B200_OK) r = dmalloc(ctx, (void **)&keys2, cap * 8);
    if (r == B200_OK) r = dmalloc(ctx, (void **)&vals, cap * sizeof(VT));
    if (r == B200_OK) r = dmalloc(ctx, (void **)&vals2, cap * sizeof(VT));
    if (r == B200_OK && n) {
        const int g = (int)std::max<u64>(1, std::min<u64>((A->rows + 7) / 8, (u64)ctx->num_sms * 16));
        k_perm_keys<VT><<<g, 256, 0, s>>>(A->rows, A->d_rp, A->d_col, d_inv, d_inv, bits, keys);
        ctx->launches++;
        if (cudaMemcpyAsync(vals, A->d_val, n * sizeof(VT), cudaMemcpyDeviceToDevice, s) != cudaSuccess) r = set_err(B200_ERR_CUDA, "permute: copying the values failed");
    }
    if (r == B200_OK) r = coo_build<VT>(ctx, A->rows, A->cols, n, keys, vals, keys2, vals2, bits, bits, 0, out);
    cleanup();
    return r;
}

// perm[new] = old (host array of A.rows entries); A must be square.
extern "C" int b200_csr_permute(b200_ctx *ctx, const b200_csr *A, const uint32_t *perm, b200_csr **out) {
    if (!ctx || !A || !out || (A->rows && !perm)) return set_err(B200_ERR_BADARG, "NULL argument");
    if (A->rows != A->cols) return set_err(B200_ERR_SHAPE, "permute: the matrix must be square (%llux%llu)", (ull)A->rows, (ull)A->cols);
    CUDA_TRY(cudaSetDevice(ctx->device));
    RESOLVE(ctx, A);
    cudaStream_t s = ctx->stream;
    const u64 n = A->rows;
    u32 *d_perm = nullptr, *d_inv = nullptr;
    TRY(dmalloc(ctx, (void **)&d_perm, std::max<u64>(n, 1) * 4));
    int r = dmalloc(ctx, (void **)&d_inv, std::max<u64>(n, 1) * 4);
    if (r == B200_OK && n) {
        cudaMemcpyAsync(d_perm, perm, n * 4, cudaMemcpyHostToDevice, s);
        cudaMemsetAsync(d_inv, 0xFF, n * 4, s);
        cudaMemsetAsync(ctx->d_flag, 0, 4, s);
        k_invert_perm<<<(int)std::max<u64>(1, std::min<u64>((n + 255) / 256, (u64)ctx->num_sms * 16)), 256, 0, s>>>(n, d_perm, d_inv, ctx->d_flag);
        ctx->launches++;
        if (cudaMemcpyAsync(ctx->h_flag, ctx->d_flag, 4, cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess)
            r = set_err(B200_ERR_CUDA, "permute: %s", cudaGetErrorString(cudaGetLastError()));
        else if (ctx->h_flag[0]) r = set_err(B200_ERR_BADARG, "permute: perm is not a permutation of 0..n-1");
    }
    if (r == B200_OK) r = A->val_bits == 32 ? permute_t<u32>(ctx, A, d_inv, out) : permute_t<u64>(ctx, A, d_inv, out);
    dfree(ctx, d_perm); dfree(ctx, d_inv);
    return r;
}

// Reverse Cuthill-McKee order of A's pattern, perm[new] = old (host array of A.rows entries).  CsrMatrix::rcm,
// /root/reference/src/graph_csr.rs:663-723, step for step: per unvisited seed in index order a plain BFS whose last
// dequeued node becomes the start, then a BFS from the start appending each node's unvisited neighbours by ascending
// degree (equal degrees keep adjacency order: Rust's sort_unstable is an insertion sort on the short slices this sees),
// the whole order reversed.
extern "C" int b200_csr_rcm_order(b200_ctx *ctx, const b200_csr *A, uint32_t *perm_out) {
    if (!ctx || !A || (A->rows && !perm_out)) return set_err(B200_ERR_BADARG, "NULL argument");
    if (A->rows != A->cols) return set_err(B200_ERR_SHAPE, "rcm: the matrix must be square (%llux%llu)", (ull)A->rows, (ull)A->cols);
    CUDA_TRY(cudaSetDevice(ctx->device));
    RESOLVE(ctx, A);
    const u64 n = A->rows;
    std::vector<u64> rp(n + 1);
    std::vector<u32> col(std::max<u64>(A->nnz, 1));
    CUDA_TRY(cudaMemcpyAsync(rp.data(), A->d_rp, (n + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (A->nnz) CUDA_TRY(cudaMemcpyAsync(col.data(), A->d_col, A->nnz * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    std::vector<unsigned char> visited(n, 0);
    std::vector<u32> stamp(n, 0), queue(std::max<u64>(n, 1)), order, nbrs;
    order.reserve(n);
    u32 epoch = 0;
    auto deg = [&](u32 v) { return rp[v + 1] - rp[v]; };
    for (u64 seed = 0; seed < n; seed++) {
        if (visited[seed]) continue;
        // (the first BFS's own visited set is an epoch stamp: the reference allocates a fresh vector per seed)
        if (++epoch == 0) { std::fill(stamp.begin(), stamp.end(), 0u); epoch = 1; }
        u64 qh = 0, qt = 0; u32 last = (u32)seed;
        queue[qt++] = (u32)seed; stamp[seed] = epoch;
        while (qh < qt) {
            const u32 u = queue[qh++]; last = u;
            for (u64 i = rp[u]; i < rp[u + 1]; i++) { const u32 v = col[i]; if (stamp[v] != epoch) { stamp[v] = epoch; queue[qt++] = v; } }
        }
        qh = qt = 0; queue[qt++] = last; visited[last] = 1;
        while (qh < qt) {
            const u32 u = queue[qh++]; order.push_back(u);
            nbrs.clear();
            for (u64 i = rp[u]; i < rp[u + 1]; i++) { const u32 v = col[i]; if (!visited[v]) nbrs.push_back(v); }
            std::stable_sort(nbrs.begin(), nbrs.end(), [&](u32 a, u32 b) { return deg(a) < deg(b); });
            for (u32 v : nbrs) if (!visited[v]) { visited[v] = 1; queue[qt++] = v; }
        }
    }
    for (u64 i = 0; i < n; i++) perm_out[i] = order[n - 1 - i];
    return B200_OK;
}
