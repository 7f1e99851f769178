// api.cu -- extern "C" entry points (include/b200_spgemm.h) and host orchestration of the kernels.
// No CPU fallback anywhere in this file: every compute entry point runs CUDA kernels or fails.
#include "engine.cuh"
#include "kernels.cuh"
#include "gen.cuh"

// ---------------------------------------------------------------------------- error plumbing
static thread_local std::string g_last_error;
int set_err(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    g_last_error = buf;
    return code;
}

extern "C" const char *b200_last_error(void) { return g_last_error.c_str(); }

extern "C" int b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

void trace_mark(b200_ctx *ctx, int line) {
    cudaEvent_t e; cudaEventCreate(&e);
    cudaEventRecord(e, ctx->cur_stream ? ctx->cur_stream : ctx->stream);
    ctx->marks->push_back(std::make_pair(line, e));
    ctx->cur_stream = nullptr;
}
void trace_dump(b200_ctx *ctx, const char *what) {
    if (!ctx->trace || ctx->marks->empty()) return;
    cudaDeviceSynchronize();
    fprintf(stderr, "[b200 trace] %s\n", what);
    for (size_t i = 0; i < ctx->marks->size(); i++) {
        float ms = 0; cudaEventElapsedTime(&ms, (*ctx->marks)[0].second, (*ctx->marks)[i].second);
        fprintf(stderr, "   line %4d done at %8.1f us\n", (*ctx->marks)[i].first, ms * 1e3f);
    }
    for (auto &m : *ctx->marks) cudaEventDestroy(m.second);
    ctx->marks->clear();
}


template <typename VT>
static CsrView<VT> view(const b200_csr *m) {
    CsrView<VT> v; v.rows = m->rows; v.cols = m->cols; v.nnz = m->nnz; v.rp = m->d_rp; v.col = m->d_col; v.val = (const VT *)m->d_val;
    return v;
}


static void entry_cache_flush(b200_ctx *ctx);
int dmalloc(b200_ctx *ctx, void **p, size_t bytes) {
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMallocAsync(p, bytes, ctx->stream);
    if (e != cudaSuccess && ctx->entry_cache && !ctx->entry_cache->empty()) {   // out of memory with blocks parked in the entry cache: release them, retry
        cudaGetLastError();
        entry_cache_flush(ctx);
        e = cudaMallocAsync(p, bytes, ctx->stream);
    }
    if (e != cudaSuccess) { cudaGetLastError(); return set_err(B200_ERR_ALLOC, "cudaMallocAsync(%zu bytes) failed: %s", bytes, cudaGetErrorString(e)); }
    return B200_OK;
}
void dfree(b200_ctx *ctx, void *p) { if (p) cudaFreeAsync(p, ctx->stream); }

#define B200_CTRL_BYTES 512
int ensure_row_scratch(b200_ctx *ctx, u64 rows) {
    if (rows > ctx->cap_rows) {
        dfree(ctx, ctx->d_prod); dfree(ctx, ctx->d_nnz_row); dfree(ctx, ctx->d_bin_rows); dfree(ctx, ctx->d_tmp_ptr); dfree(ctx, ctx->d_win);
        ctx->d_prod = nullptr; ctx->d_nnz_row = nullptr; ctx->d_bin_rows = nullptr; ctx->d_tmp_ptr = nullptr; ctx->d_win = nullptr; ctx->cap_rows = 0;
        u64 cap = rows + rows / 8 + 1024;
        TRY(dmalloc(ctx, (void **)&ctx->d_prod, cap * 8));
        TRY(dmalloc(ctx, (void **)&ctx->d_tmp_ptr, (cap + 1) * 8));
        TRY(dmalloc(ctx, (void **)&ctx->d_nnz_row, cap * 4));
        TRY(dmalloc(ctx, (void **)&ctx->d_bin_rows, cap * 4 * (B200_BIN_WIDE0 + B200_NUM_HASH_BINS)));   // every bin has its own list
        TRY(dmalloc(ctx, (void **)&ctx->d_win, cap * sizeof(uint4)));
        ctx->cap_rows = cap;
    }
    const u64 tiles = (rows + SCAN_TILE - 1) / SCAN_TILE + 1, tiles_pre = rows / 8 + 2;  // pre-pass: >= 8 rows per CTA
    if (!ctx->d_scan || tiles > ctx->cap_tiles || tiles_pre > ctx->cap_tiles_pre) {
        dfree(ctx, ctx->d_scan); ctx->d_scan = nullptr; ctx->scan_clean_bytes = 0;
        const u64 ct = tiles + 64, cp = tiles_pre + tiles_pre / 8 + 1024;
        TRY(dmalloc(ctx, (void **)&ctx->d_scan, B200_CTRL_BYTES + (ct + cp) * 8));
        ctx->d_ctrl = (B200Ctrl *)ctx->d_scan;
        ctx->d_tile_status = (u64 *)(ctx->d_scan + B200_CTRL_BYTES);
        ctx->d_tile_pre = ctx->d_tile_status + ct;
        ctx->cap_tiles = ct; ctx->cap_tiles_pre = cp;
    }
    return B200_OK;
}
// zero the control block and the scan status words a multiply over `rows` rows will use
// (skipped when the previous multiply's compaction kernel already left them zeroed: scan_clean_bytes)
static cudaError_t reset_scan(b200_ctx *ctx, u64 tiles_pre) {
    const size_t need = B200_CTRL_BYTES + (ctx->cap_tiles + tiles_pre) * 8;
    const bool clean = ctx->scan_clean_bytes >= need;
    ctx->scan_clean_bytes = 0;                                             // whoever asked is about to use the area
    return clean ? cudaSuccess : cudaMemsetAsync(ctx->d_scan, 0, need, ctx->stream);
}
static int ensure_tmp(b200_ctx *ctx, size_t col_bytes, size_t val_bytes) {
    if (col_bytes > ctx->cap_tmp_col) {
        dfree(ctx, ctx->d_tmp_col); ctx->d_tmp_col = nullptr; ctx->cap_tmp_col = 0;
        const size_t want = col_bytes + col_bytes / 4;
        TRY(dmalloc(ctx, &ctx->d_tmp_col, want));
        ctx->cap_tmp_col = want;
    }
    if (val_bytes > ctx->cap_tmp_val) {
        dfree(ctx, ctx->d_tmp_val); ctx->d_tmp_val = nullptr; ctx->cap_tmp_val = 0;
        const size_t want = val_bytes + val_bytes / 4;
        TRY(dmalloc(ctx, &ctx->d_tmp_val, want));
        ctx->cap_tmp_val = want;
    }
    return B200_OK;
}
static int ensure_heavy_scratch(b200_ctx *ctx, size_t bytes) {
    if (bytes > ctx->cap_heavy) {
        dfree(ctx, ctx->d_heavy); ctx->d_heavy = nullptr; ctx->cap_heavy = 0;
        TRY(dmalloc(ctx, &ctx->d_heavy, bytes));
        ctx->cap_heavy = bytes;
    }
    return B200_OK;
}

// ---------------------------------------------------------------------------- kernel attribute setup
template <typename K>
static void allow_big_smem(K kernel, size_t optin) {
    // static shared memory counts against the opt-in limit
    cudaFuncAttributes fa;
    size_t stat = 1024;
    if (cudaFuncGetAttributes(&fa, kernel) == cudaSuccess) stat = fa.sharedSizeBytes; else cudaGetLastError();
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(optin - stat));
    if (e != cudaSuccess) { fprintf(stderr, "b200: cudaFuncSetAttribute(max dynamic smem) failed: %s\n", cudaGetErrorString(e)); cudaGetLastError(); }
}

template <typename VT>
static void setup_kernels_vt(size_t optin) {
    allow_big_smem(k_num_cta<VT, 0>, optin); allow_big_smem(k_num_cta<VT, 1>, optin);
    allow_big_smem(k_num_rank<VT, 0>, optin); allow_big_smem(k_num_rank<VT, 1>, optin);
    allow_big_smem(k_num_warp<VT, 0>, optin); allow_big_smem(k_num_warp<VT, 1>, optin);
    allow_big_smem(k_num_expand<VT, 0, false, false, false>, optin); allow_big_smem(k_num_expand<VT, 0, false, false, true>, optin); allow_big_smem(k_num_expand<VT, 0, false, true, false>, optin); allow_big_smem(k_num_expand<VT, 0, false, true, true>, optin);
    allow_big_smem(k_num_expand<VT, 0, true, false, false>, optin); allow_big_smem(k_num_expand<VT, 0, true, false, true>, optin); allow_big_smem(k_num_expand<VT, 0, true, true, false>, optin); allow_big_smem(k_num_expand<VT, 0, true, true, true>, optin);
    allow_big_smem(k_num_expand<VT, 1, false, false, false>, optin); allow_big_smem(k_num_expand<VT, 1, false, false, true>, optin); allow_big_smem(k_num_expand<VT, 1, false, true, false>, optin); allow_big_smem(k_num_expand<VT, 1, false, true, true>, optin);
    allow_big_smem(k_num_expand<VT, 1, true, false, false>, optin); allow_big_smem(k_num_expand<VT, 1, true, false, true>, optin); allow_big_smem(k_num_expand<VT, 1, true, true, false>, optin); allow_big_smem(k_num_expand<VT, 1, true, true, true>, optin);
}

// ---------------------------------------------------------------------------- configuration
// Tuning switches live in a plain struct passed through the ABI (b200_config, the MagnusConfig::default() analogue of
// /root/reference/src/graph_magnus.rs:227): b200_config_default fills the defaults, b200_ctx_configure installs a copy.
// The B200_* environment variables of round 1 are read ONCE, when a context is created, as developer overrides of the
// defaults; nothing on the multiply path touches the environment.
extern "C" int b200_config_default(b200_config *c) {
    if (!c) return set_err(B200_ERR_BADARG, "config is NULL");
    memset(c, 0, sizeof(*c));
    c->struct_bytes = (uint32_t)sizeof(b200_config);
    c->pipeline = 0; c->placement = -1; c->exact_limit_mb = -1; c->force_acc_mode = -1; c->window_cap_groups = -1; c->window_mul = 3;
    c->circular_windows = 1; c->arc_window = 1; c->touched_span = 1; c->narrow_scratch = 1; c->expand_kernel = 1; c->pack_b = -1;
    c->lanes_per_entry_lg = -1; c->expand_div = 8; c->hash_div = 32; c->grid_div = 8; c->grid_mul = 4; c->aux_streams = 1;
    c->fused_threads = 0; c->fused_window_cols = 0; c->fused_dense_pmax = 0; c->heavy_chunk_cols = 0; c->heavy_kernel = 1; c->heavy_min_products = 0; c->heavy_unit_products = 0; c->narrow_download = 0;
    c->commute_swap = 1;
    return B200_OK;
}
static void config_from_env(b200_config *c) {
    struct { const char *name; int32_t *field; } tab[] = {
        {"B200_PIPELINE", &c->pipeline}, {"B200_EXACT", &c->placement}, {"B200_EXACT_MB", &c->exact_limit_mb}, {"B200_FORCE_MODE", &c->force_acc_mode},
        {"B200_WINCAP", &c->window_cap_groups}, {"B200_WINMUL", &c->window_mul}, {"B200_CIRCULAR", &c->circular_windows}, {"B200_ARC", &c->arc_window},
        {"B200_SPAN", &c->touched_span}, {"B200_NARROW", &c->narrow_scratch}, {"B200_EXPAND", &c->expand_kernel}, {"B200_PACK", &c->pack_b},
        {"B200_LG", &c->lanes_per_entry_lg}, {"B200_EDIV", &c->expand_div}, {"B200_TDIV", &c->hash_div}, {"B200_GDIV", &c->grid_div},
        {"B200_GMUL", &c->grid_mul}, {"B200_NAUX", &c->aux_streams}, {"B200_FUSED_THREADS", &c->fused_threads}, {"B200_FUSED_WINDOW", &c->fused_window_cols},
        {"B200_FUSED_PMAX", &c->fused_dense_pmax}, {"B200_FUSED_RING", &c->fused_ring_slots}, {"B200_FUSED_PBUF", &c->fused_product_slots}, {"B200_HEAVY_CHUNK", &c->heavy_chunk_cols}, {"B200_HEAVY_KERNEL", &c->heavy_kernel}, {"B200_HEAVY_PMIN", &c->heavy_min_products}, {"B200_HEAVY_UNIT", &c->heavy_unit_products}, {"B200_NARROW_DL", &c->narrow_download},
        {"B200_RW_CAP", &c->rw_cap_percent}, {"B200_COMMUTE", &c->commute_swap},
    };
    for (auto &t : tab) { const char *v = getenv(t.name); if (v && *v) *t.field = atoi(v); }
}
extern "C" int b200_ctx_configure(b200_ctx *ctx, const b200_config *c) {
    if (!ctx || !c) return set_err(B200_ERR_BADARG, "NULL argument");
    if (c->struct_bytes != sizeof(b200_config)) return set_err(B200_ERR_BADARG, "b200_config.struct_bytes is %u, this library expects %zu (call b200_config_default first)", c->struct_bytes, sizeof(b200_config));
    ctx->cfg = *c;
    ctx->naux_enabled = std::max(0, std::min(B200_NAUX, (int)c->aux_streams));
    return B200_OK;
}
extern "C" int b200_ctx_get_config(b200_ctx *ctx, b200_config *c) {
    if (!ctx || !c) return set_err(B200_ERR_BADARG, "NULL argument");
    *c = ctx->cfg;
    return B200_OK;
}

static int env_int_early(const char *name) { const char *v = getenv(name); return v && *v ? atoi(v) : 0; }
extern "C" int b200_ctx_destroy(b200_ctx *ctx);
static void narrow_destroy(b200_ctx *ctx);
static void narrow_reap(b200_ctx *ctx);
// (inside b200_ctx_create, once the context exists: a failing call releases what has been set up so far)
#define CUDA_TRY_X(expr)                                                                           \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            const int _c = set_err(_e == cudaErrorMemoryAllocation ? B200_ERR_ALLOC : B200_ERR_CUDA, \
                                   "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            b200_ctx_destroy(ctx); cudaGetLastError();                                             \
            return _c;                                                                             \
        }                                                                                          \
    } while (0)
extern "C" int b200_ctx_create(int device, void *cuda_stream, b200_ctx **out) {
    if (!out) return set_err(B200_ERR_BADARG, "b200_ctx_create: out is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return set_err(B200_ERR_CUDA, "no CUDA device available (%s); this engine has no CPU fallback",
                       e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (device < 0 || device >= n) return set_err(B200_ERR_BADARG, "device %d out of range (0..%d)", device, n - 1);
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return set_err(B200_ERR_CUDA, "device %d is sm_%d%d; the kernels are built for sm_100a only", device, prop.major, prop.minor);
    b200_ctx *ctx = new b200_ctx();
    memset(ctx, 0, sizeof(*ctx));
    ctx->device = device; ctx->num_sms = prop.multiProcessorCount; ctx->smem_optin = prop.sharedMemPerBlockOptin; ctx->total_mem = prop.totalGlobalMem;
    if (cuda_stream) { ctx->stream = (cudaStream_t)cuda_stream; ctx->own_stream = false; }
    else { CUDA_TRY_X(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)); ctx->own_stream = true; }
    cudaMemPool_t pool;
    CUDA_TRY_X(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t thresh = UINT64_MAX;
    CUDA_TRY_X(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
    static_assert(sizeof(B200Ctrl) <= B200_CTRL_BYTES, "control block outgrew its slot");
    CUDA_TRY_X(cudaMallocHost((void **)&ctx->h_ctrl, sizeof(B200Ctrl) + 64));
    memset(ctx->h_ctrl, 0, sizeof(B200Ctrl) + 64);
    CUDA_TRY_X(cudaMallocHost((void **)&ctx->h_report, sizeof(B200Ctrl) * 2));
    memset(ctx->h_report, 0, sizeof(B200Ctrl) * 2);
    CUDA_TRY_X(cudaMalloc((void **)&ctx->d_flag, 128));
    CUDA_TRY_X(cudaMallocHost((void **)&ctx->h_flag, 128));
    for (int i = 0; i < 4; i++) CUDA_TRY_X(cudaEventCreate(&ctx->ev[i]));
    for (int i = 0; i < B200_NAUX; i++) { CUDA_TRY_X(cudaStreamCreateWithFlags(&ctx->aux[i], cudaStreamNonBlocking)); CUDA_TRY_X(cudaEventCreateWithFlags(&ctx->ev_join[i], cudaEventDisableTiming)); }
    CUDA_TRY_X(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    CUDA_TRY_X(cudaStreamCreateWithFlags(&ctx->copy, cudaStreamNonBlocking));
    CUDA_TRY_X(cudaEventCreateWithFlags(&ctx->ev_ready, cudaEventDisableTiming));
    // one auxiliary stream measured as good as three, with half the event calls
    b200_config_default(&ctx->cfg);
    config_from_env(&ctx->cfg);
    ctx->naux_enabled = std::max(0, std::min(B200_NAUX, (int)ctx->cfg.aux_streams));
    CUDA_TRY_X(cudaMallocHost((void **)&ctx->h_freport, (size_t)B200_REPORT_SLOTS * sizeof(B200Ctrl) * 2));
    memset(ctx->h_freport, 0, (size_t)B200_REPORT_SLOTS * sizeof(B200Ctrl) * 2);
    for (int i = 0; i < B200_REPORT_SLOTS; i++) for (int j = 0; j < 3; j++) CUDA_TRY_X(cudaEventCreate(&ctx->f_ev[i][j]));
    ctx->f_dirty = true;
    fz_setup(ctx);
    rw_setup(ctx);
    rwf_setup(ctx);
    hv_setup(ctx);
    dn_setup(ctx);
    lm_setup(ctx);
    ctx->lm_min_list = env_int_early("B200_LM_MINLIST") > 0 ? (double)env_int_early("B200_LM_MINLIST") : 48.0;
    ctx->cap_cta_tot = 8192;
    CUDA_TRY_X(cudaMalloc((void **)&ctx->d_cta_tot, ctx->cap_cta_tot * 8));
    CUDA_TRY_X(cudaMalloc((void **)&ctx->d_lm_tot, ctx->cap_cta_tot * 8));   // slice totals of the left multiply: zero between multiplies
    CUDA_TRY_X(cudaMemset(ctx->d_lm_tot, 0, ctx->cap_cta_tot * 8)); ctx->lm_tot_dirty = false;
    ctx->timing = true;
    ctx->hosttime = env_int_early("B200_HOSTTIME") != 0;
    ctx->trace = env_int_early("B200_TRACE") != 0; ctx->marks = new std::vector<std::pair<int, cudaEvent_t>>();
    ctx->entry_cache = new std::vector<std::pair<void *, size_t>>(); ctx->entry_cache_bytes = 0;
    setup_kernels_vt<u32>(ctx->smem_optin);
    setup_kernels_vt<u64>(ctx->smem_optin);
    allow_big_smem(k_sym_cta<false>, ctx->smem_optin); allow_big_smem(k_sym_cta<true>, ctx->smem_optin);
    allow_big_smem(k_num_cta<u64, 2>, ctx->smem_optin); allow_big_smem(k_num_rank<u64, 2>, ctx->smem_optin);
    allow_big_smem(k_num_warp<u64, 2>, ctx->smem_optin);
    allow_big_smem(k_sym_expand<false>, ctx->smem_optin); allow_big_smem(k_sym_expand<true>, ctx->smem_optin);
    allow_big_smem(k_num_expand<u64, 2, false, false, false>, ctx->smem_optin); allow_big_smem(k_num_expand<u64, 2, false, false, true>, ctx->smem_optin); allow_big_smem(k_num_expand<u64, 2, false, true, false>, ctx->smem_optin); allow_big_smem(k_num_expand<u64, 2, false, true, true>, ctx->smem_optin);
    allow_big_smem(k_num_expand<u64, 2, true, false, false>, ctx->smem_optin); allow_big_smem(k_num_expand<u64, 2, true, false, true>, ctx->smem_optin); allow_big_smem(k_num_expand<u64, 2, true, true, false>, ctx->smem_optin); allow_big_smem(k_num_expand<u64, 2, true, true, true>, ctx->smem_optin);
    cudaGetLastError();
    *out = ctx;
    return B200_OK;
}

extern "C" int b200_ctx_destroy(b200_ctx *ctx) {
    if (!ctx) return B200_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->copy);
    cudaStreamSynchronize(ctx->stream);
    narrow_destroy(ctx);
    if (ctx->entry_cache) { for (auto &b : *ctx->entry_cache) dfree(ctx, b.first); delete ctx->entry_cache; ctx->entry_cache = nullptr; }
    dfree(ctx, ctx->d_prod); dfree(ctx, ctx->d_tmp_ptr); dfree(ctx, ctx->d_nnz_row); dfree(ctx, ctx->d_bin_rows); dfree(ctx, ctx->d_win); dfree(ctx, ctx->d_scan); dfree(ctx, ctx->d_heavy); dfree(ctx, ctx->d_tmp_col); dfree(ctx, ctx->d_tmp_val); dfree(ctx, ctx->d_hv);
    dfree(ctx, ctx->d_fz); dfree(ctx, ctx->d_units); dfree(ctx, ctx->d_rowwin); dfree(ctx, ctx->d_rowclass); dfree(ctx, ctx->d_spill_acc); dfree(ctx, ctx->d_spill_col);
    cudaStreamSynchronize(ctx->stream);
    cudaFreeHost(ctx->h_freport);
    for (int i = 0; i < B200_REPORT_SLOTS; i++) for (int j = 0; j < 3; j++) cudaEventDestroy(ctx->f_ev[i][j]);
    if (ctx->d_rowstat) cudaFree(ctx->d_rowstat);
    cudaFreeHost(ctx->h_ctrl); cudaFreeHost(ctx->h_report); cudaFree(ctx->d_flag); cudaFreeHost(ctx->h_flag); cudaFree(ctx->d_cta_tot); cudaFree(ctx->d_lm_tot);
    for (int i = 0; i < 4; i++) cudaEventDestroy(ctx->ev[i]);
    for (int i = 0; i < B200_NAUX; i++) { cudaStreamDestroy(ctx->aux[i]); cudaEventDestroy(ctx->ev_join[i]); }
    cudaEventDestroy(ctx->ev_fork);
    cudaStreamDestroy(ctx->copy); cudaEventDestroy(ctx->ev_ready);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return B200_OK;
}

extern "C" int b200_ctx_synchronize(b200_ctx *ctx) {
    if (!ctx) return set_err(B200_ERR_BADARG, "ctx is NULL");
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->copy));
    if (ctx->post) { CUDA_TRY(cudaStreamSynchronize(ctx->post)); narrow_reap(ctx); }
    return B200_OK;
}
extern "C" int b200_ctx_kernel_launches(b200_ctx *ctx, uint64_t *out) {
    if (!ctx || !out) return set_err(B200_ERR_BADARG, "NULL argument");
    *out = ctx->launches; return B200_OK;
}
extern "C" int b200_ctx_set_timing(b200_ctx *ctx, int enabled) {
    if (!ctx) return set_err(B200_ERR_BADARG, "ctx is NULL");
    ctx->timing = enabled != 0; return B200_OK;
}

// ---------------------------------------------------------------------------- CSR handles
// col_idx and values of m->nnz entries in ONE allocation (values behind the columns, 256-byte aligned)
static void entry_cache_flush(b200_ctx *ctx) {
    if (!ctx->entry_cache) return;
    for (auto &b : *ctx->entry_cache) dfree(ctx, b.first);
    ctx->entry_cache->clear(); ctx->entry_cache_bytes = 0;
    cudaStreamSynchronize(ctx->stream);                                   // the pool can hand the memory out again
}
int alloc_entries(b200_ctx *ctx, b200_csr *m) {
    const u64 cap = std::max(m->nnz, m->cap_entries);                     // products of the fused path are allocated from a bound
    m->cap_entries = cap;
    const size_t col_bytes = ((size_t)cap * 4 + 255) & ~(size_t)255;
    // (+32: the chunked heavy-row kernel fetches B-row segments as 16-byte aligned bulk copies, which may run a few entries past nnz)
    const size_t need = col_bytes + (size_t)cap * (size_t)(m->val_bits / 8) + 32;
    m->d_col = nullptr;
    if (ctx->entry_cache && need >= ((size_t)1 << 20)) {
        // smallest cached block that holds the request without wasting more than a quarter of it
        int best = -1;
        for (size_t i = 0; i < ctx->entry_cache->size(); i++) {
            const size_t b = (*ctx->entry_cache)[i].second;
            if (b >= need && b <= need + need / 4 && (best < 0 || b < (*ctx->entry_cache)[best].second)) best = (int)i;
        }
        if (best >= 0) {
            m->d_col = (u32 *)(*ctx->entry_cache)[best].first; m->entry_bytes = (*ctx->entry_cache)[best].second;
            ctx->entry_cache_bytes -= m->entry_bytes;
            ctx->entry_cache->erase(ctx->entry_cache->begin() + best);
        }
    }
    if (!m->d_col) {
        TRY(dmalloc(ctx, (void **)&m->d_col, need));
        m->entry_bytes = need;
    }
    m->d_val = (unsigned char *)m->d_col + col_bytes;
    m->val_shares_col = true;
    return B200_OK;
}
int csr_alloc(b200_ctx *ctx, u64 rows, u64 cols, u64 nnz, int val_bits, bool alloc_arrays, b200_csr **out) {
    b200_csr *m = new b200_csr();
    memset(m, 0, sizeof(*m));
    m->rows = rows; m->cols = cols; m->nnz = nnz; m->val_bits = val_bits; m->ctx = ctx;
    m->cr_start = 0; m->cr_len = cols;
    m->ar_start = 0; m->ar_len = rows;
    m->pending_slot = -1; m->max_row_span = cols;
    static u64 next_uid = 0;
    m->uid = __sync_add_and_fetch(&next_uid, 1); m->lin_base = m->uid; m->lin_pow = 1;
    int r = dmalloc(ctx, (void **)&m->d_rp, (rows + 1) * 8 + 16);           // row_ptr + the max-value scalar
    if (r == B200_OK) m->d_maxval = (ull *)(m->d_rp + rows + 1);
    if (r == B200_OK && alloc_arrays) r = alloc_entries(ctx, m);
    if (r != B200_OK) { dfree(ctx, m->d_rp); delete m; return r; }
    *out = m;
    return B200_OK;
}

extern "C" int b200_csr_free(b200_ctx *ctx, b200_csr *m) {
    if (!m) return B200_OK;
    if (!ctx) ctx = m->ctx;
    if (m->pending_slot >= 0 && ctx->slot_owner[m->pending_slot] == m) ctx->slot_owner[m->pending_slot] = nullptr;   // its report is simply never read
    delete m->stats;
    if (m->ev_copy) { cudaStreamWaitEvent(ctx->stream, m->ev_copy, 0); cudaEventDestroy(m->ev_copy); }   // frees are ordered after a pending download
    dfree(ctx, m->d_rp);
    // large entry arrays go to the context's cache (ordered on ctx->stream like a free); small ones back to the pool
    if (m->d_col && m->val_shares_col && m->entry_bytes >= ((size_t)1 << 20) && ctx->entry_cache &&
        ctx->entry_cache->size() < 16 && ctx->entry_cache_bytes + m->entry_bytes <= ctx->total_mem / 3) {
        ctx->entry_cache->push_back({(void *)m->d_col, m->entry_bytes});
        ctx->entry_cache_bytes += m->entry_bytes;
    } else dfree(ctx, m->d_col);
    if (!m->val_shares_col) dfree(ctx, m->d_val);
    dfree(ctx, m->d_desc); dfree(ctx, m->d_span); dfree(ctx, m->d_cspan); dfree(ctx, m->d_pack);
    delete m;
    return B200_OK;
}

static int grid_for(u64 n, int threads, int cap) { u64 g = (n + threads - 1) / threads; if (g < 1) g = 1; if (g > (u64)cap) g = cap; return (int)g; }

// value max + format check (explicit zeros, column range); synchronises when `check`.  `device_rowptr`: the
// row_ptr never passed through the host, so its sanity and the longest row are established on the device too.
int finish_new_csr(b200_ctx *ctx, b200_csr *m, bool check, bool device_rowptr) {
    CUDA_TRY(cudaMemsetAsync(m->d_maxval, 0, 16, ctx->stream));
    CUDA_TRY(cudaMemsetAsync(ctx->d_flag, 0, 32, ctx->stream));
    CUDA_TRY(cudaMemsetAsync(ctx->d_flag + 16, 0, 48, ctx->stream));
    if (device_rowptr && m->rows) {
        k_rowptr_stats<<<grid_for(m->rows, 256, ctx->num_sms * 8), 256, 0, ctx->stream>>>(m->rows, m->nnz, m->d_rp, ctx->d_flag + 4, ctx->d_flag);
        LAUNCH_CHECK(ctx);
    }
    if (m->nnz) {
        int g = grid_for(m->nnz, 256, ctx->num_sms * 8);
        if (m->val_bits == 32) k_value_stats<u32><<<g, 256, 0, ctx->stream>>>(m->nnz, (const u32 *)m->d_val, m->d_col, m->cols, m->d_maxval, ctx->d_flag, ctx->d_flag + 16);
        else k_value_stats<u64><<<g, 256, 0, ctx->stream>>>(m->nnz, (const u64 *)m->d_val, m->d_col, m->cols, m->d_maxval, ctx->d_flag, ctx->d_flag + 16);
        LAUNCH_CHECK(ctx);
    }
    if (m->rows && m->nnz) {
        // columns strictly ascending inside every row (the kernels take a row's first / last column as its min / max and
        // merge or binary-search sorted rows); flag bit 4
        k_rows_sorted<<<grid_for(m->rows, 256, ctx->num_sms * 8), 256, 0, ctx->stream>>>(m->rows, m->d_rp, m->d_col, m->nnz, ctx->d_flag);
        LAUNCH_CHECK(ctx);
        TRY(fz_row_span(ctx, m));
    }
    if (check) {
        CUDA_TRY(cudaMemcpyAsync(ctx->h_flag, ctx->d_flag, 128, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (ctx->h_flag[0] & 2u) return set_err(B200_ERR_FORMAT, "row_ptr is not monotone from 0 to nnz");
        if (ctx->h_flag[0] & 4u) return set_err(B200_ERR_FORMAT, "column indices are not strictly ascending inside every row");
        if (ctx->h_flag[0]) return set_err(B200_ERR_FORMAT, "CSR holds an explicit zero value or a column index >= cols");
        m->max_row_span = m->nnz ? (u64)ctx->h_flag[26] : 0;
        if (device_rowptr) m->max_row_len = ctx->h_flag[4];
        if (m->nnz && m->cols) {                                            // circular column range (see k_value_stats)
            const long long n = (long long)m->cols, half = n / 2, quarter = n / 4, ref = (long long)ctx->h_flag[24];
            m->cr_start = 0; m->cr_len = m->cols;
            for (int f = 0; f < 4; f++) {                                   // shortest arc over the four cuts
                const long long omin = (long long)(~ctx->h_flag[16 + 2 * f]), omax = (long long)ctx->h_flag[17 + 2 * f];
                if (omax < omin || (u64)(omax - omin + 1) >= m->cr_len) continue;
                long long start = ref + omin - half + (long long)f * quarter;
                start %= n; if (start < 0) start += n;
                m->cr_start = (u32)start; m->cr_len = (u64)(omax - omin + 1);
            }
        }
    }
    return B200_OK;
}

// (also finds the active row arc of the handle: the complement of the longest run of empty rows on the row circle)
static int check_host_rowptr(u64 rows, const uint64_t *row_ptr, u64 *nnz, u64 *max_len, u64 *ar_start, u64 *ar_len) {
    if (row_ptr[0] != 0) return set_err(B200_ERR_FORMAT, "row_ptr[0] must be 0");
    u64 ml = 0;
    const u64 none = ~0ull;
    u64 first = none, last = none, gap_len = 0, gap_end = 0;              // longest run of empty rows between two non-empty ones
    for (u64 i = 0; i < rows; i++) {
        if (row_ptr[i + 1] < row_ptr[i]) return set_err(B200_ERR_FORMAT, "row_ptr is not monotone at row %llu", (ull)i);
        u64 l = row_ptr[i + 1] - row_ptr[i]; if (l > ml) ml = l;
        if (l) {
            if (first == none) first = i;
            else if (i - last - 1 > gap_len) { gap_len = i - last - 1; gap_end = i; }
            last = i;
        }
    }
    *nnz = row_ptr[rows]; *max_len = ml;
    *ar_start = 0; *ar_len = rows;
    if (first != none) {
        const u64 wrap = (rows - 1 - last) + first;                         // the run of empty rows through the end of the matrix
        if (wrap >= gap_len) { *ar_start = first; *ar_len = last - first + 1; }
        else { *ar_start = gap_end; *ar_len = rows - gap_len; }
    }
    return B200_OK;
}

static int upload_common(b200_ctx *ctx, uint64_t rows, uint64_t cols, const uint64_t *row_ptr, const void *col_idx, bool idx64,
                         const void *values, int val_bits, b200_csr **out) {
    if (!ctx || !row_ptr || !out) return set_err(B200_ERR_BADARG, "NULL argument");
    if (val_bits != 32 && val_bits != 64) return set_err(B200_ERR_BADARG, "val_bits must be 32 or 64, got %d", val_bits);
    if (rows >= 0xFFFFFFFFull || cols >= 0xFFFFFFFFull) return set_err(B200_ERR_BADARG, "rows/cols must fit in u32 (NodeId)");
    CUDA_TRY(cudaSetDevice(ctx->device));
    u64 nnz = 0, max_len = 0, ar_start = 0, ar_len = rows;
    TRY(check_host_rowptr(rows, row_ptr, &nnz, &max_len, &ar_start, &ar_len));
    if (nnz && (!col_idx || !values)) return set_err(B200_ERR_BADARG, "NULL col_idx/values with nnz > 0");
    b200_csr *m = nullptr;
    TRY(csr_alloc(ctx, rows, cols, nnz, val_bits, true, &m));
    m->max_row_len = max_len; m->ar_start = ar_start; m->ar_len = ar_len;
    cudaError_t e = cudaMemcpyAsync(m->d_rp, row_ptr, (rows + 1) * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && nnz) e = cudaMemcpyAsync(m->d_val, values, nnz * (size_t)(val_bits / 8), cudaMemcpyHostToDevice, ctx->stream);
    u64 *tmp = nullptr;
    if (e == cudaSuccess && nnz) {
        if (!idx64) e = cudaMemcpyAsync(m->d_col, col_idx, nnz * 4, cudaMemcpyHostToDevice, ctx->stream);
        else {
            int r = dmalloc(ctx, (void **)&tmp, nnz * 8);
            if (r != B200_OK) { b200_csr_free(ctx, m); return r; }
            e = cudaMemcpyAsync(tmp, col_idx, nnz * 8, cudaMemcpyHostToDevice, ctx->stream);
            if (e == cudaSuccess) {
                cudaMemsetAsync(ctx->d_flag + 8, 0, 4, ctx->stream);
                k_narrow_idx<<<grid_for(nnz, 256, ctx->num_sms * 8), 256, 0, ctx->stream>>>(nnz, tmp, m->d_col, ctx->d_flag + 8);
                ctx->launches++;
                e = cudaGetLastError();
            }
        }
    }
    if (e != cudaSuccess) { dfree(ctx, tmp); b200_csr_free(ctx, m); return set_err(B200_ERR_CUDA, "upload copy failed: %s", cudaGetErrorString(e)); }
    int r = finish_new_csr(ctx, m, true);
    dfree(ctx, tmp);
    if (r == B200_OK && idx64 && nnz && ctx->h_flag[8]) r = set_err(B200_ERR_FORMAT, "column index does not fit in u32 (NodeId)");
    if (r != B200_OK) { b200_csr_free(ctx, m); return r; }
    *out = m;
    return B200_OK;
}

extern "C" int b200_csr_upload(b200_ctx *ctx, uint64_t rows, uint64_t cols, const uint64_t *row_ptr, const uint32_t *col_idx,
                               const void *values, int val_bits, b200_csr **out) {
    return upload_common(ctx, rows, cols, row_ptr, col_idx, false, values, val_bits, out);
}
extern "C" int b200_csr_upload_idx64(b200_ctx *ctx, uint64_t rows, uint64_t cols, const uint64_t *row_ptr, const uint64_t *col_idx,
                                     const void *values, int val_bits, b200_csr **out) {
    return upload_common(ctx, rows, cols, row_ptr, col_idx, true, values, val_bits, out);
}

extern "C" int b200_csr_from_device(b200_ctx *ctx, uint64_t rows, uint64_t cols, uint64_t nnz, const void *d_row_ptr,
                                    const void *d_col_idx, const void *d_values, int val_bits, b200_csr **out) {
    if (!ctx || !d_row_ptr || !out) return set_err(B200_ERR_BADARG, "NULL argument");
    if (val_bits != 32 && val_bits != 64) return set_err(B200_ERR_BADARG, "val_bits must be 32 or 64");
    if (rows >= 0xFFFFFFFFull || cols >= 0xFFFFFFFFull) return set_err(B200_ERR_BADARG, "rows/cols must fit in u32 (NodeId)");
    CUDA_TRY(cudaSetDevice(ctx->device));
    b200_csr *m = nullptr;
    TRY(csr_alloc(ctx, rows, cols, nnz, val_bits, true, &m));
    cudaError_t e = cudaMemcpyAsync(m->d_rp, d_row_ptr, (rows + 1) * 8, cudaMemcpyDeviceToDevice, ctx->stream);
    if (e == cudaSuccess && nnz) e = cudaMemcpyAsync(m->d_col, d_col_idx, nnz * 4, cudaMemcpyDeviceToDevice, ctx->stream);
    if (e == cudaSuccess && nnz) e = cudaMemcpyAsync(m->d_val, d_values, nnz * (size_t)(val_bits / 8), cudaMemcpyDeviceToDevice, ctx->stream);
    if (e != cudaSuccess) { b200_csr_free(ctx, m); return set_err(B200_ERR_CUDA, "device copy failed: %s", cudaGetErrorString(e)); }
    m->max_row_len = std::min<u64>(nnz, cols);          // bound; replaced by the exact longest row below
    int r = finish_new_csr(ctx, m, true, true);
    if (r != B200_OK) { b200_csr_free(ctx, m); return r; }
    *out = m;
    return B200_OK;
}

extern "C" int b200_csr_info(const b200_csr *m, uint64_t *rows, uint64_t *cols, uint64_t *nnz, int *val_bits) {
    if (!m) return set_err(B200_ERR_BADARG, "matrix is NULL");
    if (nnz) RESOLVE((b200_ctx *)nullptr, m);                                            // a fused product learns its size from the device report
    if (rows) *rows = m->rows; if (cols) *cols = m->cols; if (nnz) *nnz = m->nnz; if (val_bits) *val_bits = m->val_bits;
    return B200_OK;
}
extern "C" int b200_csr_device_ptrs(const b200_csr *m, void **d_row_ptr, void **d_col_idx, void **d_values) {
    if (!m) return set_err(B200_ERR_BADARG, "matrix is NULL");
    if (d_row_ptr) *d_row_ptr = m->d_rp; if (d_col_idx) *d_col_idx = m->d_col; if (d_values) *d_values = m->d_val;
    return B200_OK;
}
extern "C" int b200_csr_max_value(b200_ctx *ctx, const b200_csr *m, uint64_t *out) {
    if (!ctx || !m || !out) return set_err(B200_ERR_BADARG, "NULL argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    TRY(host_maxval(ctx, m));
    *out = m->h_maxval;
    return B200_OK;
}

// ---- narrow downloads.  The end-to-end chain is PCIe-bound (291.6 MB of results at 57 GB/s); u64 path counts that are proven
// to fit 32 bits (the report of the multiply that produced them carries their maximum) are narrowed by a kernel on the copy
// stream, cross the bus as u32 in chunks, and a small pool of host threads widens every chunk into the caller's u64 array
// while the next chunk is in flight.
#include <thread>
#include <mutex>
#include <condition_variable>
#include <immintrin.h>
#define NARROW_CHUNK ((size_t)4 << 20)        // elements per chunk (16 MiB on the bus)
// u32 -> u64 with streaming stores: the destination is written once and not read here, so no read-for-ownership traffic
__attribute__((target("avx2"))) static void widen_avx2(const u32 *src, u64 *dst, size_t n) {
    size_t j = 0;
    while (j < n && ((uintptr_t)(dst + j) & 31)) { dst[j] = (u64)src[j]; j++; }
    for (; j + 8 <= n; j += 8) {
        const __m256i v = _mm256_loadu_si256((const __m256i *)(src + j));
        _mm256_stream_si256((__m256i *)(dst + j), _mm256_cvtepu32_epi64(_mm256_castsi256_si128(v)));
        _mm256_stream_si256((__m256i *)(dst + j + 4), _mm256_cvtepu32_epi64(_mm256_extracti128_si256(v, 1)));
    }
    for (; j < n; j++) dst[j] = (u64)src[j];
    _mm_sfence();
}
static void widen_plain(const u32 *src, u64 *dst, size_t n) { for (size_t j = 0; j < n; j++) dst[j] = (u64)src[j]; }
#define NARROW_MIN ((u64)1 << 18)             // smaller arrays are copied as they are
struct NarrowStage { u32 *h; size_t cap; cudaEvent_t done; bool used; };
struct NarrowJob { struct NarrowState *ns; const u32 *src; u64 *dst; size_t n; };
struct NarrowState {
    std::vector<std::thread> th; std::mutex m; std::condition_variable cv, cv_done;
    const u32 *src = nullptr; u64 *dst = nullptr; size_t n = 0; long gen = 0; int remaining = 0; bool stop = false; int T = 1;
    bool avx2 = __builtin_cpu_supports("avx2");
    std::vector<NarrowStage> stages; std::vector<cudaEvent_t> ev_pool; size_t ev_next = 0;
    std::vector<NarrowJob *> jobs;           // callback arguments, freed at synchronize / destroy
    void worker(int i) {
        long seen = 0;
        while (true) {
            std::unique_lock<std::mutex> l(m);
            cv.wait(l, [&] { return stop || gen != seen; });
            if (stop) return;
            seen = gen;
            const u32 *s_ = src; u64 *d_ = dst; const size_t n_ = n;
            l.unlock();
            const size_t lo = n_ * (size_t)i / (size_t)T, hi = n_ * (size_t)(i + 1) / (size_t)T;
            if (avx2) widen_avx2(s_ + lo, d_ + lo, hi - lo); else widen_plain(s_ + lo, d_ + lo, hi - lo);
            l.lock();
            if (--remaining == 0) cv_done.notify_all();
        }
    }
    void run(const u32 *s_, u64 *d_, size_t n_) {          // called from one CUDA callback thread at a time (one post stream)
        std::unique_lock<std::mutex> l(m);
        src = s_; dst = d_; n = n_; remaining = T; gen++;
        cv.notify_all();
        cv_done.wait(l, [&] { return remaining == 0; });
    }
};
static void CUDART_CB narrow_cb(void *arg) { NarrowJob *j = (NarrowJob *)arg; j->ns->run(j->src, j->dst, j->n); }
__global__ void __launch_bounds__(256) k_narrow_vals(u64 n, const u64 *__restrict__ in, u32 *__restrict__ out) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) out[i] = (u32)in[i];
}
static NarrowState *narrow_state(b200_ctx *ctx) {
    if (ctx->narrow) return ctx->narrow;
    NarrowState *ns = new NarrowState();
    int cores = (int)std::thread::hardware_concurrency();
    cpu_set_t set; CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof(set), &set) == 0) cores = CPU_COUNT(&set);
    ns->T = std::max(2, std::min(16, cores / 2));
    for (int i = 0; i < ns->T; i++) ns->th.emplace_back([ns, i] { ns->worker(i); });
    if (cudaStreamCreateWithFlags(&ctx->post, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); ctx->post = nullptr; }
    ctx->narrow = ns;
    return ns;
}
static void narrow_reap(b200_ctx *ctx) {                    // after the post stream has drained
    if (!ctx->narrow) return;
    for (NarrowJob *j : ctx->narrow->jobs) delete j;
    ctx->narrow->jobs.clear();
    for (auto &st : ctx->narrow->stages) st.used = false;
    ctx->narrow->ev_next = 0;
}
static void narrow_destroy(b200_ctx *ctx) {
    NarrowState *ns = ctx->narrow;
    if (!ns) return;
    if (ctx->post) cudaStreamSynchronize(ctx->post);
    narrow_reap(ctx);
    { std::lock_guard<std::mutex> l(ns->m); ns->stop = true; }
    ns->cv.notify_all();
    for (auto &t : ns->th) t.join();
    for (auto &st : ns->stages) { cudaFreeHost(st.h); cudaEventDestroy(st.done); }
    for (auto e : ns->ev_pool) cudaEventDestroy(e);
    if (ctx->post) cudaStreamDestroy(ctx->post);
    delete ns; ctx->narrow = nullptr; ctx->post = nullptr;
}
// pinned staging of at least `elems` u32: a stage not in use since the last synchronize, else a new one
static NarrowStage *narrow_stage(NarrowState *ns, size_t elems) {
    NarrowStage *best = nullptr;
    for (auto &st : ns->stages) if (!st.used && st.cap >= elems && (!best || st.cap < best->cap)) best = &st;
    if (!best) {
        NarrowStage st; st.cap = elems + elems / 8; st.used = false; st.h = nullptr;
        if (cudaMallocHost((void **)&st.h, st.cap * 4) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        if (cudaEventCreateWithFlags(&st.done, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); cudaFreeHost(st.h); return nullptr; }
        ns->stages.push_back(st);
        best = &ns->stages.back();
    }
    best->used = true;
    return best;
}
static cudaEvent_t narrow_event(NarrowState *ns) {
    if (ns->ev_next == ns->ev_pool.size()) { cudaEvent_t e; if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return nullptr; } ns->ev_pool.push_back(e); }
    return ns->ev_pool[ns->ev_next++];
}

extern "C" int b200_csr_download_async(b200_ctx *ctx, const b200_csr *m, uint64_t *row_ptr, uint32_t *col_idx, void *values) {
    if (!ctx || !m) return set_err(B200_ERR_BADARG, "NULL argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    RESOLVE(ctx, m);
    // the copies run on the context's copy stream, after everything queued on the compute stream so far
    b200_csr *mm = const_cast<b200_csr *>(m);
    if (!mm->ev_copy) CUDA_TRY(cudaEventCreateWithFlags(&mm->ev_copy, cudaEventDisableTiming));
    // values that fit 32 bits: narrow on the device, widen on the host (the staging is allocated before anything is queued)
    NarrowState *ns = nullptr; NarrowStage *stage = nullptr; u32 *d_narrow = nullptr;
    if (values && m->val_bits == 64 && ctx->cfg.narrow_download && m->nnz >= NARROW_MIN && m->h_maxval_known && m->h_maxval <= 0xFFFFFFFFull) {
        ns = narrow_state(ctx);
        if (ctx->post) stage = narrow_stage(ns, m->nnz);
        if (stage && dmalloc(ctx, (void **)&d_narrow, m->nnz * 4) != B200_OK) { stage->used = false; stage = nullptr; }
    }
    CUDA_TRY(cudaEventRecord(ctx->ev_ready, ctx->stream));
    CUDA_TRY(cudaStreamWaitEvent(ctx->copy, ctx->ev_ready, 0));
    if (row_ptr) CUDA_TRY(cudaMemcpyAsync(row_ptr, m->d_rp, (m->rows + 1) * 8, cudaMemcpyDeviceToHost, ctx->copy));
    if (col_idx && m->nnz) CUDA_TRY(cudaMemcpyAsync(col_idx, m->d_col, m->nnz * 4, cudaMemcpyDeviceToHost, ctx->copy));
    if (stage) {
        k_narrow_vals<<<grid_for(m->nnz, 256, ctx->num_sms * 4), 256, 0, ctx->copy>>>(m->nnz, (const u64 *)m->d_val, d_narrow);
        ctx->launches++;
        for (size_t off = 0; off < m->nnz; off += NARROW_CHUNK) {
            const size_t n = std::min<size_t>(NARROW_CHUNK, m->nnz - off);
            CUDA_TRY(cudaMemcpyAsync(stage->h + off, d_narrow + off, n * 4, cudaMemcpyDeviceToHost, ctx->copy));
            cudaEvent_t e = narrow_event(ns);
            if (!e) return set_err(B200_ERR_CUDA, "download: event creation failed");
            CUDA_TRY(cudaEventRecord(e, ctx->copy));
            CUDA_TRY(cudaStreamWaitEvent(ctx->post, e, 0));
            NarrowJob *job = new NarrowJob{ns, stage->h + off, (u64 *)values + off, n};
            ns->jobs.push_back(job);
            CUDA_TRY(cudaLaunchHostFunc(ctx->post, narrow_cb, job));
        }
        CUDA_TRY(cudaFreeAsync(d_narrow, ctx->copy));
    } else if (values && m->nnz) CUDA_TRY(cudaMemcpyAsync(values, m->d_val, m->nnz * (size_t)(m->val_bits / 8), cudaMemcpyDeviceToHost, ctx->copy));
    CUDA_TRY(cudaEventRecord(mm->ev_copy, ctx->copy));
    return B200_OK;
}
extern "C" int b200_csr_download(b200_ctx *ctx, const b200_csr *m, uint64_t *row_ptr, uint32_t *col_idx, void *values) {
    TRY(b200_csr_download_async(ctx, m, row_ptr, col_idx, values));
    CUDA_TRY(cudaStreamSynchronize(ctx->copy));
    if (ctx->post) CUDA_TRY(cudaStreamSynchronize(ctx->post));             // (narrow downloads: the host threads' widening runs behind the copies)
    return B200_OK;
}
extern "C" int b200_csr_download_idx64(b200_ctx *ctx, const b200_csr *m, uint64_t *row_ptr, uint64_t *col_idx, void *values) {
    if (!ctx || !m) return set_err(B200_ERR_BADARG, "NULL argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    RESOLVE(ctx, m);
    u64 *tmp = nullptr;
    if (col_idx && m->nnz) {
        TRY(dmalloc(ctx, (void **)&tmp, m->nnz * 8));
        k_widen_idx<<<grid_for(m->nnz, 256, ctx->num_sms * 8), 256, 0, ctx->stream>>>(m->nnz, m->d_col, tmp);
        LAUNCH_CHECK(ctx);
        CUDA_TRY(cudaMemcpyAsync(col_idx, tmp, m->nnz * 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    TRY(b200_csr_download_async(ctx, m, row_ptr, nullptr, values));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->copy));
    dfree(ctx, tmp);
    return B200_OK;
}

// ---------------------------------------------------------------------------- SpGEMM

// rows of the handle's active arc (mean row lengths in the heuristics: see b200_csr::ar_len)
static inline u64 act_rows(const b200_csr *m) { return std::max<u64>(1, std::min<u64>(m->ar_len, m->rows)); }
// lanes cooperating on one A entry while walking its B row: largest power of two <= mean B row length / 2
int pick_lg(const b200_ctx *ctx, const b200_csr *B, int max_lg) {
    const int forced = ctx->cfg.lanes_per_entry_lg;
    if (forced >= 0) return std::min(forced, max_lg);
    const double avg = B->rows ? (double)B->nnz / (double)act_rows(B) : 1.0;
    int lg = 0;
    while (lg < max_lg && (double)(2 << lg) <= avg / 2.0) lg++;
    return lg;
}
// threads that own one row of hash bin hb: ~one group per 4 A entries, assuming deg_A ~ cap/2
static int bin_threads(const b200_ctx *ctx, int hb, int lg) {
    const int div = std::max(1, (int)ctx->cfg.hash_div);
    long t = ((long)b200_hash_cap(hb) << lg) / div;
    t = std::max(32L, std::min(1024L, t));
    long need = (long)b200_hash_slots(hb) / 16;                           // sort path keeps <= 16 slots per thread
    return (int)std::max(t, std::min(1024L, need));
}

// B row descriptors {start,len}: built once per right operand and cached in the handle
int ensure_desc(b200_ctx *ctx, const b200_csr *B) {
    if (B->d_desc) return B200_OK;
    if (B->nnz >= 0xFFFFFFFFull) return set_err(B200_ERR_BADARG, "right operand with >= 2^32 stored entries is not supported");
    b200_csr *Bm = const_cast<b200_csr *>(B);
    TRY(dmalloc(ctx, (void **)&Bm->d_desc, (B->rows + 1) * sizeof(uint2)));
    TRY(dmalloc(ctx, (void **)&Bm->d_span, (B->rows + 1) * sizeof(uint4)));
    k_build_desc<<<grid_for(B->rows, 256, ctx->num_sms * 8), 256, 0, ctx->stream>>>(B->rows, B->d_rp, B->d_col, Bm->d_desc, Bm->d_span);
    LAUNCH_CHECK(ctx);
    return B200_OK;
}

// circular column spans of a square right operand (see k_build_cspan), cached like the descriptors
static int ensure_cspan(b200_ctx *ctx, const b200_csr *B) {
    if (B->d_cspan || B->rows != B->cols) return B200_OK;
    b200_csr *Bm = const_cast<b200_csr *>(B);
    TRY(dmalloc(ctx, (void **)&Bm->d_cspan, (B->rows + 1) * sizeof(uint4)));
    k_build_cspan<<<grid_for(B->rows, 256, ctx->num_sms * 8), 256, 0, ctx->stream>>>(B->rows, B->d_rp, B->d_col, Bm->d_cspan);
    LAUNCH_CHECK(ctx);
    return B200_OK;
}

// operand-wide circular column offsets of a square right operand (k_cspan_bounds): one reduction, read back and cached
int ensure_cs_bounds(b200_ctx *ctx, const b200_csr *B) {
    if (B->cs_state) return B200_OK;
    b200_csr *Bm = const_cast<b200_csr *>(B);
    if (B->rows != B->cols || B->nnz == 0) { Bm->cs_state = 2; return B200_OK; }
    CUDA_TRY(cudaMemsetAsync(ctx->d_flag + 12, 0, 12, ctx->stream));
    k_cspan_bounds<<<grid_for(B->rows, 256, ctx->num_sms * 8), 256, 0, ctx->stream>>>(B->rows, B->d_rp, B->d_col, ctx->d_flag + 12);
    LAUNCH_CHECK(ctx);
    CUDA_TRY(cudaMemcpyAsync(ctx->h_flag + 12, ctx->d_flag + 12, 12, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    const long long half = (long long)(B->cols / 2);
    Bm->cs_lo = (long long)(~ctx->h_flag[12]) - half; Bm->cs_hi = (long long)ctx->h_flag[13] - half;
    Bm->cs_state = ctx->h_flag[14] == 0 && Bm->cs_lo <= Bm->cs_hi ? 1 : 2;   // rows too long to scan: no bound
    return B200_OK;
}

// sector-packed records for low-degree right operands (mean row length <= 4)
bool want_pack(const b200_ctx *ctx, const b200_csr *B) {
    const int forced = ctx->cfg.pack_b;
    if (forced >= 0) return forced != 0;
    return B->rows && (double)B->nnz / (double)act_rows(B) <= 4.0;
}
int ensure_pack(b200_ctx *ctx, const b200_csr *B) {
    if (B->d_pack) return B200_OK;
    b200_csr *Bm = const_cast<b200_csr *>(B);
    TRY(dmalloc(ctx, (void **)&Bm->d_pack, (B->rows + 1) * 2 * sizeof(uint4)));
    k_build_pack<<<grid_for(B->rows, 256, ctx->num_sms * 8), 256, 0, ctx->stream>>>(B->rows, B->d_rp, B->d_col, Bm->d_pack);
    LAUNCH_CHECK(ctx);
    return B200_OK;
}

// host copy of a handle's largest value (read back once; products get it from their own final read-back)
int host_maxval(b200_ctx *ctx, const b200_csr *m) {
    RESOLVE(ctx, m);
    if (m->h_maxval_known) return B200_OK;
    b200_csr *mm = const_cast<b200_csr *>(m);
    CUDA_TRY(cudaMemcpyAsync(ctx->h_flag + 2, m->d_maxval, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    memcpy(&mm->h_maxval, ctx->h_flag + 2, 8);
    mm->h_maxval_known = true;
    return B200_OK;
}

static int launch_row_products(b200_ctx *ctx, const b200_csr *A, const b200_csr *B) {
    const double avg = A->rows ? (double)A->nnz / (double)A->rows : 0.0;
    const u64 rows = A->rows;
#define RP_LAUNCH(G) k_row_products<G><<<(unsigned)((rows * G + 255) / 256), 256, 0, ctx->stream>>>(rows, A->d_rp, A->d_col, B->d_desc, ctx->d_prod)
    if (avg <= 2.0) RP_LAUNCH(1);
    else if (avg <= 6.0) RP_LAUNCH(4);
    else if (avg <= 24.0) RP_LAUNCH(8);
    else RP_LAUNCH(32);
#undef RP_LAUNCH
    LAUNCH_CHECK(ctx);
    return B200_OK;
}

// The final scan's last CTA writes the control block into pinned host memory as {word, epoch} chunks; the host waits until
// every chunk carries this multiply's epoch and reassembles the block into ctx->h_ctrl.  This costs one PCIe write latency
// instead of a memcpy + stream synchronise.  A stream query every so often catches a failed launch (the chunks would
// never arrive).
static int wait_for_report(b200_ctx *ctx, u32 epoch) {
    volatile u64 *chunk = ctx->h_report;
    u32 *out = reinterpret_cast<u32 *>(ctx->h_ctrl);
    const u32 n = sizeof(B200Ctrl) / 4;
    u64 spins = 0;
    for (u32 i = 0; i < n; i++) {
        u64 c;
        while ((u32)((c = chunk[i]) >> 32) != epoch) {
            if ((++spins & 0xFFF) == 0) {
                const cudaError_t q = cudaStreamQuery(ctx->stream);
                if (q == cudaErrorNotReady) { cudaGetLastError(); continue; }
                if (q != cudaSuccess) return set_err(B200_ERR_CUDA, "multiply failed on the device: %s", cudaGetErrorString(q));
                CUDA_TRY(cudaStreamSynchronize(ctx->stream));              // finished: the report must be there now
                if ((u32)(chunk[i] >> 32) != epoch) return set_err(B200_ERR_CUDA, "the row_ptr scan finished without reporting (epoch %u)", epoch);
            }
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        }
        out[i] = (u32)c;
    }
    __sync_synchronize();
    return B200_OK;
}

// row_ptr scan: 2048-row tiles either way; up to 512 K rows as 1024 threads x 2 rows (coalesced stores, all tiles
// co-resident), beyond that as 256 threads x 8 rows
static void launch_scan_rowptr(b200_ctx *ctx, u64 rows, u64 *rp, cudaStream_t s, u64 *host_mirror, u32 epoch) {
    const unsigned tiles = (unsigned)((rows + SCAN_TILE - 1) / SCAN_TILE);
    static_assert(SCAN_TILE == 1024 * 2, "both variants cut the rows into SCAN_TILE pieces");
    if (rows <= (1ull << 19)) k_scan_rowptr<1024, 2><<<tiles, 1024, 0, s>>>(rows, ctx->d_nnz_row, rp, ctx->d_tile_status, ctx->d_ctrl, host_mirror, epoch);
    else k_scan_rowptr<SCAN_THREADS, SCAN_ITEMS><<<tiles, SCAN_THREADS, 0, s>>>(rows, ctx->d_nnz_row, rp, ctx->d_tile_status, ctx->d_ctrl, host_mirror, epoch);
}

static int ctas_per_sm(const b200_ctx *ctx, int threads, size_t smem) {
    return std::max(1, std::min(32, std::min(2048 / threads, (int)(ctx->smem_optin / (smem + 1024)))));
}

// accumulator width from a proof that no row sum can overflow: rows_bound * max(A) * max(B)
int pick_mode_bits(const b200_ctx *ctx, int val_bits, u64 max_row_products, u64 maxA, u64 maxB) {
    unsigned __int128 bound = (unsigned __int128)max_row_products * maxA;
    bool over64 = (bound >> 64) != 0;
    if (!over64) { bound *= maxB; over64 = (bound >> 64) != 0; }
    int mode;
    if (!over64 && (u64)bound < (1ull << 32)) mode = 0;
    else if (val_bits == 32) mode = 1;                                    // clamped 32-bit products, < 2^32 of them per row
    else mode = over64 ? 2 : 1;
    const int forced = ctx->cfg.force_acc_mode;                           // testing hook: a wider mode is always valid
    if (forced > mode && forced <= (val_bits == 64 ? 2 : 1)) mode = forced;
    return mode;
}

// k_num_expand's compile-time switches (packed B records, pattern-only B, touched-span tracking) from run-time flags
template <typename VT, int MODE>
static void expand_dispatch(bool packed, bool bpat, bool span, int eg, int et, size_t smem, cudaStream_t bs, const NumArgs<VT> &na,
                            const uint4 *pack, const u32 *bin_rows, B200Ctrl *ctrl, int bin, int nb, u32 pcap, u32 ncap, u32 nw4,
                            const uint4 *win, u32 ncols, const OutArgs<VT> &o) {
#define K(P, Q, S) k_num_expand<VT, MODE, P, Q, S><<<eg, et, smem, bs>>>(na, P ? pack : nullptr, bin_rows, ctrl, bin, nb, pcap, ncap, nw4, win, ncols, o)
    switch ((packed ? 4 : 0) | (bpat ? 2 : 0) | (span ? 1 : 0)) {
        case 0: K(false, false, false); break;
        case 1: K(false, false, true); break;
        case 2: K(false, true, false); break;
        case 3: K(false, true, true); break;
        case 4: K(true, false, false); break;
        case 5: K(true, false, true); break;
        case 6: K(true, true, false); break;
        default: K(true, true, true); break;
    }
#undef K
}

// Rows the row-per-warp kernels (rowwarp.cu) take in the exact placement: bitmap lists of hash bins 0..hb_max.
struct RwPlan { int hb_max; u32 nw, cap; bool all_fit; };                   // hb_max < 0: none; all_fit: no row can land on a hash list of bins <= hb_max
static const RwPlan kNoRw = {-1, 0u, 0u, false};

// Last kernel of an exact-placement multiply whose result was handed out before its size was known: the control block
// goes to the pinned report ring, the product's largest value to its handle, and the control block and scan status words
// are left zeroed for the next multiply.
__global__ void __launch_bounds__(256) k_finish_exact(B200Ctrl *ctrl, u64 *host_mirror, u32 epoch, ull *maxval_dst, u64 *scan_area,
                                                      u32 ctrl_words, u64 scan_words) {
    if (blockIdx.x == 0) {
        const volatile u32 *src = reinterpret_cast<const volatile u32 *>(ctrl);
        for (u32 i = threadIdx.x; i < sizeof(B200Ctrl) / 4; i += blockDim.x) st_volatile_u64(host_mirror + i, ((u64)epoch << 32) | (u64)src[i]);
        if (threadIdx.x == 0) *maxval_dst = *reinterpret_cast<volatile ull *>(&ctrl->max_val_out);
        __syncthreads();
        for (u32 i = threadIdx.x; i < ctrl_words; i += blockDim.x) scan_area[i] = 0;
    }
    for (u64 i = ctrl_words + (u64)blockIdx.x * blockDim.x + threadIdx.x; i < scan_words; i += (u64)gridDim.x * blockDim.x) scan_area[i] = 0;
}

// Launch the numeric kernels of every list the pre-pass filled.  The list sizes live on the device only, so grids are
// sized from `rows` and bins that no row can reach (p_bound) are skipped.
template <typename VT>
static int launch_numeric(b200_ctx *ctx, const b200_csr *A, const b200_csr *B, u64 rows, u64 p_bound, u64 heavy_cap,
                          int mode, bool packed, bool bpat, int lg, OutArgs<VT> o, Fan &fan, const WinCaps &caps,
                          B200Ctrl *ctrl = nullptr, bool wide_only = false, const RwPlan &rw = kNoRw, b200_csr *C = nullptr,
                          const HvPlan *hv = nullptr) {
    if (!ctrl) ctrl = ctx->d_ctrl;
    const u32 nwords = (u32)((B->cols + 31) / 32);
    const size_t smem_max = ctx->smem_optin - 1024;
    const bool hv_on = hv && hv->on;
    const HvSkip skip{ctx->d_prod, hv_on ? hv->pmin : ~0ull, hv_on ? hv->cap_li : 0u};
    NumArgs<VT> na{A->d_rp, A->d_col, (const VT *)A->d_val, B->d_desc, B->d_col, (const VT *)B->d_val};
    NumArgs<u64> na64{A->d_rp, A->d_col, (const u64 *)A->d_val, B->d_desc, B->d_col, (const u64 *)B->d_val};
    OutArgs<u64> o64{o.base, o.col, (u64 *)o.val, o.nnz_out, o.bin_cnt, o.bin_stride, o.narrow};
    const size_t accb = mode == 0 ? 4 : 8;
    auto reachable = [&](int hb) { return hb == 0 ? (p_bound > 32 || A->max_row_len > 32) : p_bound > (u64)b200_hash_cap(hb - 1); };
    auto do_tiny = [&]() -> int {
        if (wide_only) return B200_OK;
        const int g = (int)std::min<u64>((rows + 31) / 32, (u64)ctx->num_sms * 32);   // >= 4 rows per warp: the software pipeline needs a row stream
        k_num_tiny<VT><<<g, 256, 0, fan.pick()>>>(na, ctx->d_bin_rows, ctrl, o, bpat, B->cols < (1ull << 27), mode == 0);
        LAUNCH_CHECK(ctx);
        return B200_OK;
    };
    // one-pass bins are cut on the product count P <= cap: expand the products into shared memory once and run every
    // later phase one product per thread (k_num_expand).  Returns false when the bin's buffers do not fit shared memory.
    const u32 nw4_full = caps.full;
    auto launch_expand = [&](int bin, int nb, u32 cap, u32 nw4, u64 n, cudaStream_t bs) -> bool {
        if (nw4 == 0 || wide_only || bin - B200_BIN_HASH0 <= rw.hb_max) return false;   // (rows of bins <= rw.hb_max: row-per-warp kernel)
        const u32 pcap = cap, ncap = (u32)std::min<u64>(cap, B->cols);
        const size_t pvb = (mode == 0 || sizeof(VT) == 4) ? 4 : 8;
        const size_t ex_smem = (size_t)nw4 * 24 + (size_t)pcap * (4 + pvb) + (size_t)ncap * (4 + accb);
        if (ex_smem + (packed ? 0 : sizeof(EnumSmem)) > smem_max) return false;
        // threads: ~8 products or ~8 bitmap groups each, whichever asks for more
        const int et = std::max(32, std::min(512, (int)std::max<u32>(pcap, nw4) / std::max(1, (int)ctx->cfg.expand_div) / 32 * 32));
        // a CTA should see several rows (one bitmap/accumulator clear and one pipeline fill per CTA): at most rows/8 CTAs
        const u64 gdiv = (u64)std::max(1, (int)ctx->cfg.grid_div), gmul = (u64)std::max(1, (int)ctx->cfg.grid_mul);
        const int eg = (int)std::max<u64>(1, std::min<u64>((n + gdiv - 1) / gdiv, (u64)ctx->num_sms * ctas_per_sm(ctx, et, ex_smem) * gmul));
        if (ctx->trace) { cudaStream_t keep = ctx->cur_stream; trace_mark(ctx, -(int)pcap); ctx->cur_stream = keep; }
        const bool span = (int)nw4 > et && ctx->cfg.touched_span;       // more bitmap groups than threads: walk only the groups a row touches
#define EXPAND(MODE, VTT, NA, OO)                                                                                                     \
        expand_dispatch<VTT, MODE>(packed, bpat, span, eg, et, ex_smem, bs, NA, B->d_pack, ctx->d_bin_rows, ctrl, bin, nb, pcap, ncap, nw4, ctx->d_win, (u32)B->cols, OO)
        if (mode == 0) EXPAND(0, VT, na, o);
        else if (mode == 1) EXPAND(1, VT, na, o);
        else EXPAND(2, u64, na64, o64);
#undef EXPAND
        return true;
    };
    // hash + in-row sort kernels (any column space): warp per row for the two smallest bins, CTA per row above
    auto launch_warp_hash = [&](int first_bin, int nb, u64 n, cudaStream_t bs) -> int {
        const size_t smem = 8 * (accb * B200_WARP_SLOTS + (size_t)B200_WARP_SLOTS * 4 + B200_WARP_ORDER_BYTES);
        const int g = (int)std::min<u64>((n + 7) / 8, (u64)ctx->num_sms * 16);
        const int wlg = std::min(lg, 5);
        if (mode == 0) k_num_warp<VT, 0><<<g, 256, smem, bs>>>(na, ctx->d_bin_rows, ctrl, first_bin, nb, wlg, o);
        else if (mode == 1) k_num_warp<VT, 1><<<g, 256, smem, bs>>>(na, ctx->d_bin_rows, ctrl, first_bin, nb, wlg, o);
        else k_num_warp<u64, 2><<<g, 256, smem, bs>>>(na64, ctx->d_bin_rows, ctrl, first_bin, nb, wlg, o64);
        LAUNCH_CHECK(ctx);
        return B200_OK;
    };
    auto launch_cta_hash = [&](int bin, int hb, u64 n, cudaStream_t bs) -> int {
        const u32 slots = b200_hash_slots(hb);
        const int threads = bin_threads(ctx, hb, lg);
        // table + the ordering step's bucket counters + the enumeration's per-thread arrays
        const size_t smem = (size_t)slots * (4 + accb) + (((size_t)b200_order_buckets(slots) + 1 + 3) & ~(size_t)3) * 4 + EnumPtrs::bytes((u32)threads);
        if (smem > smem_max) return set_err(B200_ERR_CUDA, "hash bin %d needs %zu B of shared memory", hb, smem);
        const int g = (int)std::min<u64>(n, (u64)ctx->num_sms * ctas_per_sm(ctx, threads, smem) * 4);
        if (mode == 0) k_num_cta<VT, 0><<<g, threads, smem, bs>>>(na, ctx->d_bin_rows, ctrl, bin, slots, lg, o);
        else if (mode == 1) k_num_cta<VT, 1><<<g, threads, smem, bs>>>(na, ctx->d_bin_rows, ctrl, bin, slots, lg, o);
        else k_num_cta<u64, 2><<<g, threads, smem, bs>>>(na64, ctx->d_bin_rows, ctrl, bin, slots, lg, o64);
        LAUNCH_CHECK(ctx);
        return B200_OK;
    };
    auto do_small = [&]() -> int {
        if (!(reachable(0) || reachable(1))) return B200_OK;
        // bins 0 and 1 share list HASH0+1; rows whose column window exceeds the bitmap are on the wide list
        if (launch_expand(B200_BIN_HASH0, 2, b200_hash_cap(1), caps.cap[1], rows, fan.pick())) LAUNCH_CHECK(ctx);
        if (caps.cap[1] < nw4_full && !(rw.hb_max >= 1 && rw.all_fit)) TRY(launch_warp_hash(B200_BIN_WIDE0 + 1, 1, rows, fan.pick()));
        return B200_OK;
    };
    auto do_bin = [&](int hb) -> int {
        if (!reachable(hb)) return B200_OK;
        // rows whose column window fits the bin's bitmap -> k_num_expand, the others (wide list) -> hash + sort
        cudaStream_t bs = fan.pick();
        bool used = false;
        if (launch_expand(B200_BIN_HASH0 + hb, 1, b200_hash_cap(hb), caps.cap[hb], rows, bs)) { LAUNCH_CHECK(ctx); used = true; }
        if (caps.cap[hb] < nw4_full && !(hb <= rw.hb_max && rw.all_fit)) TRY(launch_cta_hash(B200_BIN_WIDE0 + hb, hb, rows, used ? fan.pick() : bs));
        return B200_OK;
    };
    // the longest rows first: the big kernels start while the host is still queueing the small ones
    const bool heavy = p_bound > (u64)b200_hash_cap(B200_NUM_HASH_BINS - 1);
    if (rw.hb_max >= 1 && (reachable(0) || reachable(1)) && !wide_only)
        TRY(rw_launch(ctx, A, B, ctrl, B200_BIN_HASH0 + 1, rw.hb_max, false, mode, packed, bpat, rw.nw, rw.cap, C, fan.pick()));
    if (heavy) {
        const u64 n = rows;
        const size_t heavy_rank_smem = (size_t)nwords * 6 + 16 + (size_t)heavy_cap * (4 + accb);
        if (!hv_on && heavy_cap < 65536 && heavy_rank_smem <= smem_max) {
            // heavy rows over a small column space: the rank kernel with accumulators sized for the longest row
            const int g = (int)std::min<u64>(n, (u64)ctx->num_sms * 2);
            cudaStream_t bs = fan.pick();
            const u32 cap = (u32)heavy_cap;
            if (mode == 0) k_num_rank<VT, 0><<<g, 1024, heavy_rank_smem, bs>>>(na, ctx->d_bin_rows, ctrl, B200_BIN_HEAVY, cap, nwords, 5, o);
            else if (mode == 1) k_num_rank<VT, 1><<<g, 1024, heavy_rank_smem, bs>>>(na, ctx->d_bin_rows, ctrl, B200_BIN_HEAVY, cap, nwords, 5, o);
            else k_num_rank<u64, 2><<<g, 1024, heavy_rank_smem, bs>>>(na64, ctx->d_bin_rows, ctrl, B200_BIN_HEAVY, cap, nwords, 5, o64);
            LAUNCH_CHECK(ctx);
        } else {
            u64 max_slots = 2; while (max_slots < 2 * heavy_cap) max_slots <<= 1;
            // per CTA: table (keys u32, sums u64), a place and an order word per entry, the ordering step's bucket counters
            const size_t per_cta = (size_t)max_slots * 12 + (size_t)max_slots * 4 + ((size_t)max_slots / 2 + 1) * 4 + ((size_t)B200_HEAVY_NB_MAX + 1) * 4 + 256;
            const size_t budget = (size_t)8 << 30;
            const int g = (int)std::min<u64>(std::min<u64>(n, (u64)ctx->num_sms), std::max<u64>(1, budget / per_cta));
            TRY(ensure_heavy_scratch(ctx, per_cta * g));
            unsigned char *base = (unsigned char *)ctx->d_heavy;
            u64 *s_vals = (u64 *)base;                                            // g * max_slots u64
            u32 *s_keys = (u32 *)(base + (size_t)g * max_slots * 8);              // g * max_slots u32
            u32 *s_place = s_keys + (size_t)g * max_slots;                        // g * max_slots u32
            u32 *s_order = s_place + (size_t)g * max_slots;                       // g * (max_slots / 2 + 1) u32
            u32 *s_cnt = s_order + (size_t)g * (max_slots / 2 + 1);               // g * (B200_HEAVY_NB_MAX + 1) u32
            if (mode == 2) k_num_heavy<u64, 2><<<g, 1024, 0, ctx->stream>>>(na64, ctx->d_bin_rows, ctrl, ctx->d_nnz_row, max_slots, s_place, s_order, s_cnt, s_keys, s_vals, o64, skip);
            else k_num_heavy<VT, 1><<<g, 1024, 0, ctx->stream>>>(na, ctx->d_bin_rows, ctrl, ctx->d_nnz_row, max_slots, s_place, s_order, s_cnt, s_keys, s_vals, o, skip);
            LAUNCH_CHECK(ctx);
            // the rows with enough products per column chunk: dense accumulation chunk by chunk (heavy.cu)
            if (hv_on) TRY(hv_numeric(ctx, A, B, ctrl, *hv, mode, bpat, o.base, o.col, (void *)o.val, o.narrow, ctx->stream));
        }
    }
    // launch order: the bin that holds the row of mean size first (it carries most of the work), then outwards from
    // it, larger bins before smaller; the host-side estimate of the mean is nnz(A)/rows * nnz(B)/rows(B)
    const double meanP = (A->rows ? (double)A->nnz / (double)act_rows(A) : 0.0) * (B->rows ? (double)B->nnz / (double)act_rows(B) : 0.0);
    int center = -1;                                                       // -1: tiny, 1: small (hash bins 0-1), 2..: hash bin
    if (meanP > 32.0) { center = 1; while (center < B200_NUM_HASH_BINS - 1 && meanP > (double)b200_hash_cap(center)) center++; }
    auto run = [&](int id) -> int { return id < 0 ? do_tiny() : id <= 1 ? do_small() : do_bin(id); };
    const int ids[B200_NUM_HASH_BINS] = {-1, 1, 2, 3, 4, 5, 6, 7};          // ascending row size
    int ci = 0;
    for (int i = 0; i < B200_NUM_HASH_BINS; i++) if (ids[i] == center) ci = i;
    TRY(run(ids[ci]));
    for (int d = 1; d < B200_NUM_HASH_BINS; d++) {
        if (ci + d < B200_NUM_HASH_BINS) TRY(run(ids[ci + d]));
        if (ci - d >= 0) TRY(run(ids[ci - d]));
    }
    return B200_OK;
}

static int launch_sym_heavy(b200_ctx *ctx, const SymArgs &sa, u64 rows, u32 nwords, Fan &fan, B200Ctrl *ctrl = nullptr,
                            const b200_csr *A = nullptr, const b200_csr *B = nullptr, const HvPlan *hv = nullptr) {
    if (!ctrl) ctrl = ctx->d_ctrl;
    const size_t smem_max = ctx->smem_optin - 1024;
    const bool hv_on = hv && hv->on && A && B;
    const HvSkip skip{ctx->d_prod, hv_on ? hv->pmin : ~0ull, hv_on ? hv->cap_li : 0u};
    const size_t bm_smem = (((size_t)nwords + 3) & ~(size_t)3) * 4 + EnumPtrs::bytes(1024);
    if (bm_smem <= smem_max) {
        const int g = (int)std::min<u64>(rows, (u64)ctx->num_sms * 2);
        k_sym_cta<true><<<g, 1024, bm_smem, fan.pick()>>>(sa, ctx->d_bin_rows, ctrl, B200_BIN_HEAVY, 2, nwords, 5, ctx->d_nnz_row, (u32)ctx->cap_rows, skip);
    } else {
        const int g = (int)std::min<u64>(rows, (u64)ctx->num_sms);
        TRY(ensure_heavy_scratch(ctx, (size_t)g * nwords * 4));
        k_sym_heavy<<<g, 1024, 0, ctx->stream>>>(sa, ctx->d_bin_rows, ctrl, nwords, (u32 *)ctx->d_heavy, ctx->d_nnz_row, (u32)ctx->cap_rows, skip);
    }
    LAUNCH_CHECK(ctx);
    if (hv_on) TRY(hv_count(ctx, A, B, ctrl, *hv, ctx->stream));
    return B200_OK;
}

// Exact mode, first half: distinct-column counts for every list the one-pass pre-pass produced (the lists and the
// kernels mirror launch_numeric's: tiny / window bitmap / hash, heavy), so that C can be allocated at its exact size.
static int launch_counts(b200_ctx *ctx, const b200_csr *A, const b200_csr *B, const SymArgs &sa, u64 rows, u64 p_bound, bool packed, int lg,
                         Fan &fan, const WinCaps &caps, B200Ctrl *ctrl = nullptr, bool wide_only = false, const RwPlan &rw = kNoRw,
                         const HvPlan *hv = nullptr) {
    if (!ctrl) ctrl = ctx->d_ctrl;
    const u32 nwords = (u32)((B->cols + 31) / 32), nw4_full = caps.full;
    const u32 bstride = (u32)ctx->cap_rows;
    const size_t smem_max = ctx->smem_optin - 1024;
    auto reachable = [&](int hb) { return hb == 0 ? (p_bound > 32 || A->max_row_len > 32) : p_bound > (u64)b200_hash_cap(hb - 1); };
    auto count_expand = [&](int bin, int nb, u32 pcap, u32 nw4) -> int {
        if (nw4 == 0 || wide_only || bin - B200_BIN_HASH0 <= rw.hb_max) return B200_OK;
        const size_t smem = (size_t)nw4 * 16;
        const int t = std::max(32, std::min(512, (int)std::max<u32>(pcap / 2, nw4) / 8 / 32 * 32));
        const int g = (int)std::max<u64>(1, std::min<u64>((rows + 7) / 8, (u64)ctx->num_sms * ctas_per_sm(ctx, t, smem) * 4));
        cudaStream_t bs = fan.pick();
        if (packed) k_sym_expand<true><<<g, t, smem, bs>>>(sa, B->d_pack, ctx->d_bin_rows, ctrl, bin, nb, nw4, ctx->d_win, (u32)B->cols, ctx->d_nnz_row, bstride);
        else k_sym_expand<false><<<g, t, smem, bs>>>(sa, nullptr, ctx->d_bin_rows, ctrl, bin, nb, nw4, ctx->d_win, (u32)B->cols, ctx->d_nnz_row, bstride);
        LAUNCH_CHECK(ctx);
        return B200_OK;
    };
    if (rw.hb_max >= 1 && (reachable(0) || reachable(1)) && !wide_only)
        TRY(rw_launch(ctx, A, B, ctrl, B200_BIN_HASH0 + 1, rw.hb_max, true, 0, packed, true, rw.nw, rw.cap, nullptr, fan.pick()));
    if (p_bound > (u64)b200_hash_cap(B200_NUM_HASH_BINS - 1)) TRY(launch_sym_heavy(ctx, sa, rows, nwords, fan, ctrl, A, B, hv));
    for (int hb = B200_NUM_HASH_BINS - 1; hb >= 2; hb--) {
        if (!reachable(hb)) continue;
        TRY(count_expand(B200_BIN_HASH0 + hb, 1, b200_hash_cap(hb), caps.cap[hb]));
        if (caps.cap[hb] < nw4_full && !(hb <= rw.hb_max && rw.all_fit)) {
            const u32 slots = b200_hash_slots(hb);
            const int threads = bin_threads(ctx, hb, lg);
            const size_t smem = (size_t)slots * 4 + EnumPtrs::bytes((u32)threads);   // key table + the enumeration's per-thread arrays
            if (smem > smem_max) return set_err(B200_ERR_CUDA, "hash bin %d needs %zu B of shared memory", hb, smem);
            const int g = (int)std::min<u64>(rows, (u64)ctx->num_sms * ctas_per_sm(ctx, threads, smem) * 2);
            k_sym_cta<false><<<g, threads, smem, fan.pick()>>>(sa, ctx->d_bin_rows, ctrl, B200_BIN_WIDE0 + hb, slots, nwords, lg, ctx->d_nnz_row, bstride, HvSkip{nullptr, ~0ull, 0u});
            LAUNCH_CHECK(ctx);
        }
    }
    if (reachable(0) || reachable(1)) {
        TRY(count_expand(B200_BIN_HASH0, 2, b200_hash_cap(1), caps.cap[1]));
        if (caps.cap[1] < nw4_full && !(rw.hb_max >= 1 && rw.all_fit)) {
            const int g = (int)std::min<u64>((rows + 7) / 8, (u64)ctx->num_sms * 16);
            k_sym_warp<<<g, 256, 0, fan.pick()>>>(sa, ctx->d_bin_rows, ctrl, B200_BIN_WIDE0 + 1, 1, std::min(lg, 5), ctx->d_nnz_row, bstride);
            LAUNCH_CHECK(ctx);
        }
    }
    if (!wide_only) {
        const int g = (int)std::min<u64>((rows + 31) / 32, (u64)ctx->num_sms * 32);   // >= 4 rows per warp: the software pipeline needs a row stream
        k_sym_tiny<<<g, 256, 0, fan.pick()>>>(sa, ctx->d_bin_rows, ctrl, ctx->d_nnz_row);
        LAUNCH_CHECK(ctx);
    }
    return B200_OK;
}

// The same kernels for the fused pipeline's "other" rows (hash and heavy lists only; fused.cu places the rows in between)
int legacy_counts(b200_ctx *ctx, const b200_csr *A, const b200_csr *B, B200Ctrl *ctrl, u64 p_bound, int lg, Fan &fan) {
    SymArgs sa{A->d_rp, A->d_col, B->d_desc, B->d_col};
    WinCaps caps; for (int hb = 0; hb < B200_NUM_HASH_BINS; hb++) caps.cap[hb] = 0;
    caps.heavy_from = 0xFFFFFFFFu; caps.full = 1;                          // every bin has a hash ("wide") list, none a bitmap list
    return launch_counts(ctx, A, B, sa, A->rows, p_bound, false, lg, fan, caps, ctrl, true);
}
int legacy_numeric(b200_ctx *ctx, const b200_csr *A, const b200_csr *B, B200Ctrl *ctrl, b200_csr *C, u64 p_bound, u64 heavy_cap, int mode,
                   bool packed, bool bpat, int lg, Fan &fan) {
    WinCaps caps; for (int hb = 0; hb < B200_NUM_HASH_BINS; hb++) caps.cap[hb] = 0;
    caps.heavy_from = 0xFFFFFFFFu; caps.full = 1;
    const u32 bstride = (u32)ctx->cap_rows;
    if (A->val_bits == 32) {
        OutArgs<u32> o{C->d_rp, C->d_col, (u32 *)C->d_val, nullptr, ctrl->sym_bin_count, bstride, 0u};
        return launch_numeric<u32>(ctx, A, B, A->rows, p_bound, heavy_cap, mode, packed, bpat, lg, o, fan, caps, ctrl, true);
    }
    OutArgs<u64> o{C->d_rp, C->d_col, (u64 *)C->d_val, nullptr, ctrl->sym_bin_count, bstride, 0u};
    return launch_numeric<u64>(ctx, A, B, A->rows, p_bound, heavy_cap, mode, packed, bpat, lg, o, fan, caps, ctrl, true);
}

template <typename VT>
// (inside spgemm_typed: a failing CUDA call must not leak the half-built product)
#define CUDA_TRY_C(expr)                                                                           \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            b200_csr_free(ctx, C);                                                                 \
            return set_err(_e == cudaErrorMemoryAllocation ? B200_ERR_ALLOC : B200_ERR_CUDA,       \
                           "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
        }                                                                                          \
    } while (0)
static int spgemm_typed(b200_ctx *ctx, const b200_csr *A, const b200_csr *B, b200_csr **out, b200_stats *st) {
    const u64 rows = A->rows, ncols = B->cols;
    cudaStream_t s = ctx->stream;
    const u64 launches0 = ctx->launches;
    const bool timing = ctx->timing && st;
    b200_csr *C = nullptr;
    TRY(csr_alloc(ctx, rows, ncols, 0, A->val_bits, false, &C));
    C->ar_start = A->ar_start; C->ar_len = A->ar_len;                       // rows of C outside A's active arc are empty
    // mean row lengths for the heuristics below count the rows of the active arcs only (whole-size handles that hold a row
    // block plus a halo: the empty rows outside say nothing about the lists a kernel will see)
    const u64 actA = std::max<u64>(1, std::min<u64>(A->ar_len, rows)), actB = std::max<u64>(1, std::min<u64>(B->ar_len, B->rows));
    if (st) { memset(st, 0, sizeof(*st)); st->rows = rows; st->cols = ncols; st->nnz_a = A->nnz; st->nnz_b = B->nnz; }
    if (rows == 0 || A->nnz == 0 || B->nnz == 0) {
        CUDA_TRY_C(cudaMemsetAsync(C->d_rp, 0, (rows + 1) * 8 + 16, s));   // row_ptr and the max-value scalar behind it
        TRY(alloc_entries(ctx, C));
        C->h_maxval = 0; C->h_maxval_known = true;
        if (st) st->bytes_algorithmic = (A->nnz + B->nnz) * (4 + sizeof(VT)) + (A->rows + B->rows + rows + 3) * 8;
        *out = C;
        return B200_OK;
    }
    // (the left multiply reads B's row_ptr directly: where it is the likely choice, B's per-row descriptors are built only if
    //  the multiply falls through to the other pipelines -- in a swapped power chain B is a fresh product every step)
    const bool lm_try = ctx->cfg.pipeline == 6 || (ctx->cfg.pipeline == 0 && A->max_row_len <= 32 && (double)B->nnz / (double)actB >= ctx->lm_min_list &&
                                                   (double)B->nnz / (double)actB >= 4.0 * ((double)A->nnz / (double)actA));
    int r = ensure_row_scratch(ctx, rows);
    if (r == B200_OK && !lm_try) r = ensure_desc(ctx, B);
    const bool packed = want_pack(ctx, B);
    if (r == B200_OK && packed && !lm_try) r = ensure_pack(ctx, B);
    if (r == B200_OK) r = host_maxval(ctx, A);
    if (r == B200_OK) r = host_maxval(ctx, B);
    if (r != B200_OK) { b200_csr_free(ctx, C); return r; }
    SymArgs sa{A->d_rp, A->d_col, B->d_desc, B->d_col};
    const u32 nwords = (u32)((ncols + 31) / 32);
    const int lg = pick_lg(ctx, B, 5);
    const u64 p_bound = A->max_row_len * B->max_row_len;                 // host-known bound of the largest P_i
    const size_t smem_max = ctx->smem_optin - 1024;
    const u64 maxA = A->h_maxval, maxB = B->h_maxval;
    const bool bpat = maxB == 1;
    Fan fan(ctx);
    void *tmp_col = nullptr, *tmp_val = nullptr;

    // Two ways to place the rows of C (both start with the one-launch pre-pass):
    //   scratch (default): numeric kernels write every row at its bound offset prefix(min(P_i, cols)) in a scratch CSR
    //     and report its exact length; a scan of the lengths gives row_ptr and a compaction kernel moves the rows;
    //   exact: a count pass gives every row's exact length first, C is allocated at its exact size and written once.
    // The scratch needs sum(min(P_i, cols)) entries; its host-known bound nnz(A) * maxlen(B) decides (see below).
    const size_t total_b = ctx->total_mem;
    const size_t esz = 4 + sizeof(VT);
    unsigned __int128 hb128 = (unsigned __int128)A->nnz * B->max_row_len;
    const unsigned __int128 dense128 = (unsigned __int128)rows * ncols;
    if (dense128 < hb128) hb128 = dense128;
    const bool cheap_bound = hb128 * esz <= (unsigned __int128)(total_b / 16);

    // this multiply's slot in the pinned report ring (a slot still owned by an unread product is read first)
    u32 epoch = 0; int slot = 0; u64 *mirror = nullptr;
    r = claim_report_slot(ctx, &epoch, &slot, &mirror);
    if (r != B200_OK) { b200_csr_free(ctx, C); return r; }
    if (timing) { cudaEventRecord(ctx->ev[0], s); cudaEventRecord(ctx->f_ev[slot][0], s); }
    if (ctx->hosttime) ctx->ht[0] = host_now_us();
    if (ctx->trace) trace_mark(ctx, __LINE__);
    const u32 bstride = (u32)ctx->cap_rows;                               // every bin's row list has room for all rows
    u64 tmp_entries = 0;
    size_t scan_bytes = 0;                                                // control block + scan status words this multiply uses
    const int mode1 = pick_mode_bits(ctx, (int)sizeof(VT) * 8, p_bound, maxA, maxB);                 // one-pass accumulator width (host-side bound)
    // Per hash bin: the largest column window (in 128-column groups) its k_num_expand bitmap holds.  Narrow column
    // spaces fit whole; otherwise a bin with capacity P gets a window of 384*P columns (at most 512 K columns): on the
    // 100^3 torus wider windows lost to hash + sort (the per-row prefix over a mostly empty bitmap dominates).
    // Rows beyond the window take the hash + sort kernels.
    WinCaps caps;
    HvPlan hv;                                                            // chunked kernels for the heaviest rows (heavy.cu), where they apply
    r = hv_plan(ctx, A, B, mode1, p_bound, &hv);
    if (r != B200_OK) { b200_csr_free(ctx, C); return r; }
    caps.heavy_from = hv.on ? hv.heavy_from : 0xFFFFFFFFu;               // (over few column chunks the chunked kernels also beat hash + sort on the largest hash bins)
    // The arc of the index circle this multiply can touch: (column range of A, known on the host for every handle) +
    // (offsets c - k of B's entries, a per-operand constant for a square B).  When the arc is short -- a GPU's row block of
    // a torus or banded matrix, whatever the size of the whole matrix -- ONE window serves every row: the pre-pass needs
    // no per-row window (WMODE 0), every row fits the bitmap, and the product's own column range is the arc.
    u32 all_groups = (nwords + 3) / 4, all_rot = 0;
    u64 arc_start = 0, arc_len = ncols;
    const bool want_rw = ctx->cfg.pipeline == 3;                          // (auto: one-launch pipeline 4 where it applies, else the binned kernels)
    if (B->rows == B->cols && ((A->cr_len < ncols && ctx->cfg.arc_window) || (want_rw && ctx->cfg.circular_windows))) {
        r = ensure_cs_bounds(ctx, B);
        if (r != B200_OK) { b200_csr_free(ctx, C); return r; }
        if (B->cs_state == 1) {
            const long long n = (long long)ncols;
            const unsigned __int128 len = (unsigned __int128)A->cr_len + (unsigned __int128)(B->cs_hi - B->cs_lo);
            if (len < (unsigned __int128)ncols) {
                long long st = ((long long)A->cr_start + B->cs_lo) % n; if (st < 0) st += n;
                arc_start = (u64)st; arc_len = (u64)len;
            }
        }
    }
    const bool use_arc = arc_len < ncols && (arc_len + 127) / 128 + 1 < (u64)all_groups && (arc_len + 127) / 128 <= 1024;
    if (use_arc) { all_groups = (u32)((arc_len + 127) / 128); all_rot = (u32)arc_start; }
    C->cr_start = (u32)arc_start; C->cr_len = arc_len;                     // a product's columns stay inside the arc
    if (B->rows == B->cols && B->cs_state == 1 && A->max_row_span < ncols) {
        const u64 sb = A->max_row_span + (u64)(B->cs_hi - B->cs_lo);       // a row's arc grows by B's offset range
        C->max_row_span = sb <= ncols / 2 ? sb : ncols;
    }
    // Row-per-warp kernels (pipeline 3, the default): rows of 33..4096 products whose window fits a warp's bitmap.  The
    // bitmap is sized from what the host knows about the operands: the longest arc a row of A covers plus the offset range of
    // B's entries (square B), the operand-level arc, or the whole column space -- at most B200_RW_MAX_GROUPS groups.
    RwPlan rw = kNoRw;
    if (want_rw) {
        u64 groups = all_groups;
        bool all_fit = true;                                               // no row's window can exceed the bitmap: no hash lists
        if (B->rows == B->cols && ctx->cfg.circular_windows && B->cs_state == 1 && A->max_row_span < ncols) {
            const u64 width = A->max_row_span + (u64)(B->cs_hi - B->cs_lo) + 1;
            if (width + 2 < ncols / 2) groups = std::min<u64>(groups, (width + 127 + 127) / 128);
        }
        if (groups > B200_RW_MAX_GROUPS) { groups = B200_RW_MAX_GROUPS; all_fit = false; }
        if (ctx->cfg.window_cap_groups >= 0 && (u64)ctx->cfg.window_cap_groups < groups) { groups = std::max<u64>(1, (u64)ctx->cfg.window_cap_groups); all_fit = false; }
        const double meanP = ((double)A->nnz / (double)actA) * ((double)B->nnz / (double)actB);
        const double factor = ctx->cfg.rw_cap_percent > 0 ? 0.01 * ctx->cfg.rw_cap_percent : 1.4;
        u64 cap = (u64)(factor * meanP) + 32;
        cap = std::min<u64>(cap, std::min<u64>(std::min<u64>(p_bound, 8192), groups * 128));
        cap = std::max<u64>(64, std::min<u64>(2048, (cap + 31) / 32 * 32));
        rw.hb_max = B200_RW_MAX_HB; rw.nw = ((u32)groups * 4u + 31u) & ~31u; rw.cap = (u32)cap; rw.all_fit = all_fit;   // (a lane owns nw / 32 consecutive words)
    }
    // Left multiply (pipeline 6, leftmul.cu): A's rows are short, B's long -- a row of C is the union of a few long sorted rows of
    // B, streamed as contiguous lists.  Window: for square operands whose entry offsets (c - row) are bounded on the index
    // circle the window travels with the row (the bound of a product is the sum of its operands' bounds, carried on the
    // handles); else the operand-level arc or the whole column space, as pipeline 4.
    if (B->rows == B->cols && A->rows == A->cols) {
        if (A->cs_state == 0 && A->max_row_len <= 64) { r = ensure_cs_bounds(ctx, A); if (r != B200_OK) { b200_csr_free(ctx, C); return r; } }
        if (B->cs_state == 0 && B->max_row_len <= 64) { r = ensure_cs_bounds(ctx, B); if (r != B200_OK) { b200_csr_free(ctx, C); return r; } }
        if (A->cs_state == 1 && B->cs_state == 1) {
            const long long lo = A->cs_lo + B->cs_lo, hi = A->cs_hi + B->cs_hi, half = (long long)(ncols / 2);
            if (lo > -half && hi < half - 1) { C->cs_lo = lo; C->cs_hi = hi; C->cs_state = 1; }
        }
    }
    {
        const double meanA = (double)A->nnz / (double)actA, meanB = (double)B->nnz / (double)actB;
        const bool shape_ok = A->max_row_len <= 32 && meanB >= ctx->lm_min_list && meanB >= 4.0 * meanA;
        if ((ctx->cfg.pipeline == 6 || (ctx->cfg.pipeline == 0 && shape_ok)) && cheap_bound && ncols < 0xFFFF0000ull && rows < 0xFFFF0000ull &&
            ctx->cfg.window_cap_groups < 0 && ctx->cfg.placement < 0) {
            u64 words = (u64)all_groups * 4; u32 org = all_rot; bool per_row = false;
            if (C->cs_state == 1 && ctx->cfg.circular_windows) {
                const u64 w = ((u64)(C->cs_hi - C->cs_lo + 1) + 31) / 32;
                if (w < words) { words = w; per_row = true; org = (u32)(C->cs_lo < 0 ? (long long)ncols + C->cs_lo : C->cs_lo); }
            }
            u32 nw = (u32)((words + 31) & ~31ull);
            if (per_row) { nw = (u32)((words + 127) / 128); nw = (nw | 1u) * 128u; }   // 128 x odd words: conflict-free 128-bit sweeps (leftmul.cu)
            const double meanP = meanA * meanB;
            const double factor = ctx->cfg.rw_cap_percent > 0 ? 0.01 * ctx->cfg.rw_cap_percent : 1.0;
            u64 cap = (u64)(factor * meanP) + 32;
            // a row of C is a union of rows of B: when B is itself a product, its own compression (entries per intermediate
            // product) and the skew of its rows (longest / mean) predict the longest row of C better than the product count
            if (ctx->cfg.rw_cap_percent <= 0 && B->stats && B->stats->products && B->stats->nnz_c && meanB > 0.0) {
                const double cr = (double)B->stats->nnz_c / (double)B->stats->products, skew = (double)B->max_row_len / meanB;
                cap = std::min<u64>(cap, (u64)(1.1 * meanP * cr * skew) + 32);
            }
            cap = std::min<u64>(cap, std::min<u64>(p_bound, words * 32));
            cap = std::max<u64>(64, std::min<u64>(4096, (cap + 31) / 32 * 32));
            const size_t per_warp = lm_smem_per_warp(mode1, nw, (u32)cap);
            if (words <= 2048 && per_warp * 4 + 1024 <= ctx->smem_optin && (ctx->cfg.pipeline == 6 || per_warp <= 24 * 1024)) {
                if (ctx->scan_clean_bytes < B200_CTRL_BYTES) { CUDA_TRY_C(cudaMemsetAsync(ctx->d_ctrl, 0, B200_CTRL_BYTES, s)); ctx->scan_clean_bytes = B200_CTRL_BYTES; }
                if (timing) cudaEventRecord(ctx->f_ev[slot][1], s);
                C->cap_entries = std::max<u64>((u64)hb128, 1);
                r = alloc_entries(ctx, C);
                if (r == B200_OK) r = lm_launch(ctx, A, B, C, ctx->d_ctrl, mode1, org, per_row, nw, (u32)cap, ctx->cfg.fused_threads == 4 ? 0 : ctx->cfg.fused_threads == 6 ? 1 : ctx->cfg.fused_threads == 5 ? 2 : -1, mirror, epoch, s);
                if (r != B200_OK) { b200_csr_free(ctx, C); return r; }
                if (timing) cudaEventRecord(ctx->f_ev[slot][2], s);
                trace_dump(ctx, "left multiply");
                mark_pending(ctx, C, slot, epoch, A, B, mode1, 6, (int32_t)(ctx->launches - launches0), timing, std::min<u64>(p_bound, ncols));
                *out = C;
                if (st) { TRY(resolve_pending(ctx, C)); *st = *C->stats; }
                return B200_OK;
            }
        }
    }
    if (lm_try) {                                                         // not taken after all: the other pipelines need B's descriptors
        r = ensure_desc(ctx, B);
        if (r == B200_OK && packed) r = ensure_pack(ctx, B);
        if (r != B200_OK) { b200_csr_free(ctx, C); return r; }
    }
    // One pass over the products (pipeline 5, dense.cu): 32-bit sums proven, a square low-degree right operand with a known offset
    // range, C allocated from the host-known bound.
    if (ctx->cfg.pipeline == 5 && mode1 == 0 && packed && B->rows == B->cols && cheap_bound && B->nnz < 0xFFFFFFFFull && ncols < 0x7FFFFFFFull) {
        r = ensure_cs_bounds(ctx, B);
        if (r != B200_OK) { b200_csr_free(ctx, C); return r; }
        if (B->cs_state == 1) {
            if (ctx->scan_clean_bytes < B200_CTRL_BYTES) { CUDA_TRY_C(cudaMemsetAsync(ctx->d_ctrl, 0, B200_CTRL_BYTES, s)); ctx->scan_clean_bytes = B200_CTRL_BYTES; }
            if (timing) cudaEventRecord(ctx->f_ev[slot][1], s);
            C->cap_entries = std::max<u64>((u64)hb128, 1);
            r = alloc_entries(ctx, C);
            if (r == B200_OK) r = dn_launch(ctx, A, B, C, ctx->d_ctrl, bpat, ctx->cfg.fused_threads > 0 && ctx->cfg.fused_threads <= 8 ? ctx->cfg.fused_threads : 3, mirror, epoch, s);
            if (r != B200_OK) { b200_csr_free(ctx, C); return r; }
            if (timing) cudaEventRecord(ctx->f_ev[slot][2], s);
            trace_dump(ctx, "one-pass multiply (dense windows)");
            mark_pending(ctx, C, slot, epoch, A, B, mode1, 5, (int32_t)(ctx->launches - launches0), timing, std::min<u64>(p_bound, ncols));
            *out = C;
            if (st) { TRY(resolve_pending(ctx, C)); *st = *C->stats; }
            return B200_OK;
        }
    }
    // One cooperative launch for the whole multiply (pipeline 4, the default where it applies): every row shares one window
    // (the whole column space or the operand-level arc) that fits a warp's bitmap, rows are short enough for a single warp,
    // and C can be allocated from the host-known bound.  Otherwise the listed pipelines below.
    if ((ctx->cfg.pipeline == 4 || ctx->cfg.pipeline == 0) && all_groups <= B200_RW_MAX_GROUPS && p_bound <= 16384 && cheap_bound &&
        B->nnz < 0xFFFFFFFFull && ctx->cfg.window_cap_groups < 0 && ctx->cfg.placement < 0) {
        const u32 nw = (all_groups * 4u + 31u) & ~31u;
        const double meanP = ((double)A->nnz / (double)actA) * ((double)B->nnz / (double)actB);
        const double factor = ctx->cfg.rw_cap_percent > 0 ? 0.01 * ctx->cfg.rw_cap_percent : 1.4;
        u64 cap = (u64)(factor * meanP) + 32;
        cap = std::min<u64>(cap, std::min<u64>(p_bound, (u64)all_groups * 128));
        cap = std::max<u64>(64, std::min<u64>(2048, (cap + 31) / 32 * 32));
        // auto: only while a warp's share of shared memory leaves >= 16 warps per SM (wide arcs of an N-GPU row block do not:
        // the binned kernels, whose bitmap is per bin, are faster there -- measured on rank 0's block of the 8-GPU torus)
        // ... and while rows average >= 48 intermediate products: below that (the first two powers of the torus chain) a row is a
        // handful of products, the warp-per-row phases are all latency, and the binned kernels' tiny-row path is 20-30 % faster
        // (30^3: A^2 51 vs 72 us, A^3 68 vs 82 us; from A^4 on the one-launch multiply is level or ahead)
        const size_t per_warp = rw_smem_per_warp(false, mode1, nw, (u32)cap);
        // ... and while most rows are inside A's active arc: the kernel cuts the rows into static per-CTA ranges, so a whole-size
        // left operand that holds one GPU's rows keeps a fraction of the grid busy (measured: 0.71 ms against 0.10 for a rank of 8)
        if (per_warp * 4 + 1024 <= ctx->smem_optin && (ctx->cfg.pipeline == 4 || (per_warp <= 14 * 1024 && meanP >= 48.0 && actA * 2 > rows))) {
            if (ctx->scan_clean_bytes < B200_CTRL_BYTES) { CUDA_TRY_C(cudaMemsetAsync(ctx->d_ctrl, 0, B200_CTRL_BYTES, s)); ctx->scan_clean_bytes = B200_CTRL_BYTES; }
            if (timing) cudaEventRecord(ctx->f_ev[slot][1], s);
            C->cap_entries = std::max<u64>((u64)hb128, 1);
            r = alloc_entries(ctx, C);
            if (r == B200_OK) r = rw_fused_launch(ctx, A, B, C, ctx->d_ctrl, mode1, packed, bpat, all_rot, all_groups * 4u, nw, (u32)cap, mirror, epoch, s);
            if (r != B200_OK) { b200_csr_free(ctx, C); return r; }
            if (timing) cudaEventRecord(ctx->f_ev[slot][2], s);
            trace_dump(ctx, "one-launch multiply");
            mark_pending(ctx, C, slot, epoch, A, B, mode1, 4, (int32_t)(ctx->launches - launches0), timing, std::min<u64>(p_bound, ncols));
            *out = C;
            if (st) { TRY(resolve_pending(ctx, C)); *st = *C->stats; }
            return B200_OK;
        }
    }
    {
        const u32 nw4_full = all_groups;
        const size_t accb1 = mode1 == 0 ? 4 : 8, pvb1 = (mode1 == 0 || sizeof(VT) == 4) ? 4 : 8;
        const int forced = ctx->cfg.window_cap_groups;                    // testing hook: force a small window (0: hash only)
        for (int hb = 0; hb < B200_NUM_HASH_BINS; hb++) {
            const u64 pcap = b200_hash_cap(hb == 0 ? 1 : hb), ncap = std::min<u64>(pcap, ncols);
            const size_t fixed = pcap * (4 + pvb1) + ncap * (4 + accb1);
            // (with a common arc of at most 512 groups every bin takes the whole arc: one window, no wide lists)
            u64 want = std::min<u64>(nw4_full, std::max<u64>(use_arc ? 512 : 256, std::min<u64>(4096, (u64)ctx->cfg.window_mul * pcap)));
            if (forced >= 0) want = std::min<u64>(want, (u64)forced);
            const size_t avail = smem_max - (packed ? 0 : sizeof(EnumSmem));    // the balanced expansion keeps its tile in static shared memory
            const u64 fit = fixed + 24 * 32 <= avail ? (avail - fixed) / 24 : 0;
            caps.cap[hb] = ctx->cfg.expand_kernel ? (u32)std::min<u64>(want, fit) : 0u;
            if (hb <= rw.hb_max) caps.cap[hb] = std::min<u32>(rw.nw / 4, all_groups);
        }
    }
    {
        // one launch: product counts, column windows, bins, scratch offsets (look-back scan of min(P_i, cols))
        const double avgA = (double)A->nnz / (double)rows;
        const int G = avgA <= 2.0 ? 1 : avgA <= 6.0 ? 4 : avgA <= 24.0 ? 8 : 32;
        const u64 tile_rows = std::max(256 / G, B200_PREPASS_MIN_ROWS);   // B200_PREPASS_ROWS(G)
        const u64 tiles_pre = (rows + tile_rows - 1) / tile_rows;
        scan_bytes = B200_CTRL_BYTES + (ctx->cap_tiles + tiles_pre) * 8;
        CUDA_TRY_C(reset_scan(ctx, tiles_pre));
        // column windows only matter when some bin's bitmap is narrower than B; square operands get circular windows
        bool windows = false;
        for (int hb = 0; hb < B200_NUM_HASH_BINS; hb++) windows |= caps.cap[hb] < all_groups;
        if (windows && use_arc) { all_groups = (nwords + 3) / 4; all_rot = 0; }   // a bin cannot hold the arc: per-row windows over the plain column space
        caps.full = all_groups;
        if (ctx->trace) fprintf(stderr, "[b200 trace] rw: hb_max %d nw %u cap %u all_fit %d A.max_row_span %llu p_bound %llu\n", rw.hb_max, rw.nw, rw.cap, (int)rw.all_fit, (ull)A->max_row_span, (ull)p_bound);
        if (ctx->trace) fprintf(stderr, "[b200 trace] arc: A.cr=(%u,%llu) B.cs=[%lld,%lld] state %d arc=(%llu,%llu) use_arc=%d windows=%d groups=%u caps=%u %u %u %u %u %u %u %u\n",
                                A->cr_start, (ull)A->cr_len, B->cs_lo, B->cs_hi, B->cs_state, (ull)arc_start, (ull)arc_len, (int)use_arc, (int)windows, all_groups,
                                caps.cap[0], caps.cap[1], caps.cap[2], caps.cap[3], caps.cap[4], caps.cap[5], caps.cap[6], caps.cap[7]);
        const bool circular = windows && B->rows == B->cols && ctx->cfg.circular_windows;
        if (circular) { r = ensure_cspan(ctx, B); if (r != B200_OK) { b200_csr_free(ctx, C); return r; } }
        const int wmode = !windows ? 0 : circular ? 2 : 1;
#define PREPASS1(GG, WW) k_prepass<GG, WW><<<(unsigned)tiles_pre, 256, 0, s>>>(rows, A->d_rp, A->d_col, WW == 2 ? B->d_cspan : B->d_span, B->d_desc, ncols,   \
                                                                             ctx->d_prod, ctx->d_nnz_row, ctx->d_tmp_ptr, ctx->d_tile_pre, ctx->d_ctrl,   \
                                                                             ctx->d_bin_rows, bstride, ctx->d_win, caps, all_groups, all_rot)
#define PREPASS(GG) do { if (wmode == 2) PREPASS1(GG, 2); else if (wmode == 1) PREPASS1(GG, 1); else PREPASS1(GG, 0); } while (0)
        if (G == 1) PREPASS(1); else if (G == 4) PREPASS(4); else if (G == 8) PREPASS(8); else PREPASS(32);
#undef PREPASS1
#undef PREPASS
        LAUNCH_CHECK(ctx);
        // Exact mode: count every row's distinct columns first (same lists, count-only kernels), allocate C at its exact
        // size and let the numeric kernels write it once at the final offsets -- no scratch CSR, no compaction, DRAM
        // traffic close to the algorithmic bytes.  Measured (30^3 A^7: count pass ~100 us vs ~64 us of compaction; 200^3
        // A^5: 135 vs 101 ms) the scratch path is faster, so it stays the default while the scratch fits: decided from the
        // host-known bound nnz(A) * maxlen(B) when that is small (1/16 of device memory), else from the pre-pass's exact
        // figure.  B200_EXACT / B200_EXACT_MB override.
        const int exact_env = ctx->cfg.placement;
        const int exact_mb = ctx->cfg.exact_limit_mb;
        bool exact = exact_env >= 0 ? exact_env != 0 : (exact_mb >= 0 && hb128 * esz > (unsigned __int128)((u64)exact_mb << 20));
        if (rw.hb_max >= 0) exact = true;                                  // the row-per-warp kernels write C once, at its final offsets
        tmp_entries = (u64)hb128;
        if (exact_env < 0 && !exact && !cheap_bound) {
            // the host-side bound is too loose to decide: read the pre-pass's exact scratch size sum(min(P_i, cols)) (one
            // small synchronous copy; multiplies of this size run for milliseconds) and keep the faster scratch path while
            // it fits a third of the free memory
            CUDA_TRY_C(cudaMemcpyAsync(ctx->h_ctrl, ctx->d_ctrl, sizeof(B200Ctrl), cudaMemcpyDeviceToHost, s));
            CUDA_TRY_C(cudaStreamSynchronize(s));
            tmp_entries = ctx->h_ctrl->total_bound;
            // scratch the context already owns decides first (steady state of a loop: no driver query at all)
            if ((size_t)tmp_entries * 4 > ctx->cap_tmp_col || (size_t)tmp_entries * sizeof(VT) > ctx->cap_tmp_val) {
                size_t free_b = 0, tot_b = 0;
                CUDA_TRY_C(cudaMemGetInfo(&free_b, &tot_b));
                const size_t reusable = ctx->cap_tmp_col + ctx->cap_tmp_val;
                exact = (unsigned __int128)tmp_entries * esz > (unsigned __int128)((free_b + reusable) / 3);
            }
        }
        if (exact) {
            r = launch_counts(ctx, A, B, sa, rows, p_bound, packed, lg, fan, caps, nullptr, false, rw, &hv);
            fan.join();
            if (r != B200_OK) { b200_csr_free(ctx, C); return r; }
            if (rw.hb_max >= 0 && cheap_bound) {
                // ---- nothing waits for the device: C is allocated from the host-known bound nnz(A) * maxlen(B), the scan writes
                //      row_ptr, the numeric kernels write every row once at its final offset, and the last kernel reports the
                //      control block (nnz, longest row, largest value) into this multiply's slot of the pinned report ring.  The
                //      handle is handed out "pending" (resolve_pending reads the slot when somebody needs the numbers).
                launch_scan_rowptr(ctx, rows, C->d_rp, s, nullptr, 0);
                LAUNCH_CHECK(ctx);
                if (timing) cudaEventRecord(ctx->f_ev[slot][1], s);
                C->cap_entries = std::max<u64>(tmp_entries, 1);
                r = alloc_entries(ctx, C);
                if (r != B200_OK) { b200_csr_free(ctx, C); return r; }
                OutArgs<VT> o{C->d_rp, C->d_col, (VT *)C->d_val, nullptr, ctx->d_ctrl->sym_bin_count, bstride};
                r = launch_numeric<VT>(ctx, A, B, rows, p_bound, std::max<u64>(1, std::min<u64>(p_bound, ncols)), mode1, packed, bpat, lg, o, fan, caps, nullptr, false, rw, C, &hv);
                fan.join();
                if (r != B200_OK) { b200_csr_free(ctx, C); return r; }
                const unsigned fg = (unsigned)std::max<u64>(1, std::min<u64>(64, scan_bytes / 8 / 1024));
                k_finish_exact<<<fg, 256, 0, s>>>(ctx->d_ctrl, mirror, epoch, C->d_maxval, (u64 *)ctx->d_scan, B200_CTRL_BYTES / 8, scan_bytes / 8);
                LAUNCH_CHECK(ctx);
                ctx->scan_clean_bytes = scan_bytes;
                if (timing) cudaEventRecord(ctx->f_ev[slot][2], s);
                trace_dump(ctx, "exact multiply (asynchronous)");
                mark_pending(ctx, C, slot, epoch, A, B, mode1, 3, (int32_t)(ctx->launches - launches0), timing, std::min<u64>(p_bound, ncols));
                *out = C;
                if (st) { TRY(resolve_pending(ctx, C)); *st = *C->stats; }
                return B200_OK;
            }
            const u32 xepoch = ++ctx->epoch ? ctx->epoch : ++ctx->epoch;
            launch_scan_rowptr(ctx, rows, C->d_rp, s, ctx->h_report, xepoch);
            LAUNCH_CHECK(ctx);
            if (timing) cudaEventRecord(ctx->ev[1], s);
            r = wait_for_report(ctx, xepoch);
            if (r != B200_OK) { b200_csr_free(ctx, C); return r; }
            const B200Ctrl hc = *ctx->h_ctrl;
            C->nnz = hc.total_nnz; C->max_row_len = hc.max_row_nnz;
            r = alloc_entries(ctx, C);
            if (r != B200_OK) { b200_csr_free(ctx, C); return r; }
            if (timing) cudaEventRecord(ctx->ev[2], s);
            if (ctx->trace) trace_mark(ctx, __LINE__);
            OutArgs<VT> o{C->d_rp, C->d_col, (VT *)C->d_val, nullptr, ctx->d_ctrl->sym_bin_count, bstride};
            r = launch_numeric<VT>(ctx, A, B, rows, p_bound, std::max<u64>(1, hc.max_row_nnz), mode1, packed, bpat, lg, o, fan, caps, nullptr, false, rw, C, &hv);
            fan.join();
            if (r != B200_OK) { b200_csr_free(ctx, C); return r; }
            CUDA_TRY_C(cudaMemcpyAsync(C->d_maxval, &ctx->d_ctrl->max_val_out, 8, cudaMemcpyDeviceToDevice, s));   // read back lazily (host_maxval)
            if (timing) cudaEventRecord(ctx->ev[3], s);
            {
                // the exact placement has waited for the device anyway: its statistics are complete here
                b200_stats *xs = new b200_stats();
                memset(xs, 0, sizeof(*xs));
                xs->rows = rows; xs->cols = ncols; xs->nnz_a = A->nnz; xs->nnz_b = B->nnz;
                xs->nnz_c = C->nnz; xs->products = hc.total_products; xs->max_row_products = hc.max_row_products; xs->max_row_nnz = hc.max_row_nnz;
                xs->bytes_algorithmic = (A->nnz + B->nnz + C->nnz) * (4 + sizeof(VT)) + (A->rows + B->rows + rows + 3) * 8;
                xs->acc_mode = mode1; xs->kernel_launches = (int32_t)(ctx->launches - launches0);
                for (int i = 0; i < B200_STAT_BINS; i++) xs->sym_bin_rows[i] = hc.sym_bin_count[i];
                xs->pipeline = rw.hb_max >= 0 ? 3 : 2;
                if (timing && st) {
                    CUDA_TRY_C(cudaEventSynchronize(ctx->ev[3]));
                    cudaEventElapsedTime(&xs->ms_symbolic, ctx->ev[0], ctx->ev[1]);   // pre-pass + counts + row_ptr scan
                    cudaEventElapsedTime(&xs->ms_numeric, ctx->ev[2], ctx->ev[3]);    // numeric kernels into the final arrays
                    cudaEventElapsedTime(&xs->ms_total, ctx->ev[0], ctx->ev[3]);
                }
                C->stats = xs;
                if (st) *st = *xs;
            }
            trace_dump(ctx, "exact multiply");
            *out = C;
            return B200_OK;
        }
    }
    {
        r = ensure_tmp(ctx, tmp_entries * 4, tmp_entries * sizeof(VT));
        if (r != B200_OK) { b200_csr_free(ctx, C); return r; }
        tmp_col = ctx->d_tmp_col; tmp_val = ctx->d_tmp_val;
        const int mode = mode1;
        const u64 heavy_cap = std::min<u64>(p_bound, ncols);
        if (p_bound > (u64)b200_hash_cap(B200_NUM_HASH_BINS - 1) && (hv.on || !(heavy_cap < 65536 && (size_t)nwords * 6 + 16 + heavy_cap * 12 <= smem_max))) {
            // heavy rows that need the global table: size it from their exact nnz (bitmap count first); the rows the chunked
            // kernels take get their per-chunk lengths here
            r = launch_sym_heavy(ctx, sa, rows, nwords, fan, nullptr, A, B, &hv);
            fan.join();
        }
        // u64 values that provably stay below 2^32 (mode 0) cross the scratch as u32: 8 instead of 12 bytes per entry, twice
        const bool narrow = sizeof(VT) == 8 && mode == 0 && ctx->cfg.narrow_scratch;
        OutArgs<VT> o{ctx->d_tmp_ptr, (u32 *)tmp_col, (VT *)tmp_val, ctx->d_nnz_row, ctx->d_ctrl->sym_bin_count, bstride, narrow ? 1u : 0u};
        if (ctx->hosttime) ctx->ht[1] = host_now_us();                       // pre-pass enqueued
        if (r == B200_OK) r = launch_numeric<VT>(ctx, A, B, rows, p_bound, heavy_cap, mode, packed, bpat, lg, o, fan, caps, nullptr, false, kNoRw, nullptr, &hv);
        fan.join();
        if (ctx->hosttime) ctx->ht[2] = host_now_us();                       // numeric kernels enqueued
        if (r != B200_OK) { b200_csr_free(ctx, C); return r; }
        // ---- C is allocated from the scratch bound BEFORE its size is known, so nothing below waits for the device: the row_ptr
        //      scan's last CTA reports the control block (nnz, longest row, largest value, product count) into this multiply's
        //      slot of the pinned report ring, and the handle is handed out "pending": whoever needs those numbers first
        //      (b200_csr_info, a download, the next multiply taking C as an operand) reads the slot (resolve_pending).
        C->cap_entries = std::max<u64>(tmp_entries, 1);
        r = alloc_entries(ctx, C);
        if (r != B200_OK) { b200_csr_free(ctx, C); return r; }
        launch_scan_rowptr(ctx, rows, C->d_rp, s, mirror, epoch);
        LAUNCH_CHECK(ctx);
        if (timing) cudaEventRecord(ctx->f_ev[slot][1], s);
        if (ctx->trace) trace_mark(ctx, __LINE__);
        {
            // lanes per row of the compaction from an estimate of the mean row (products / ~1.5); any value is correct
            const double meanP = ((double)A->nnz / (double)actA) * (B->rows ? (double)B->nnz / (double)actB : 0.0);
            const double avg = std::min((double)ncols, meanP / 1.5);
            const int llg = avg <= 2 ? 0 : avg <= 6 ? 2 : avg <= 24 ? 3 : 5;
            const u64 want = (rows << llg) / 256 + 1;
            const unsigned cg = (unsigned)std::min<u64>(want, (u64)ctx->num_sms * 64);
            if (narrow) k_compact_rows<VT, u32><<<cg, 256, 0, s>>>(rows, ctx->d_tmp_ptr, C->d_rp, (const u32 *)tmp_col, (const u32 *)tmp_val, C->d_col, (VT *)C->d_val, llg,
                                                                  &ctx->d_ctrl->max_val_out, C->d_maxval, (u64 *)ctx->d_scan, B200_CTRL_BYTES / 8, scan_bytes / 8);
            else k_compact_rows<VT, VT><<<cg, 256, 0, s>>>(rows, ctx->d_tmp_ptr, C->d_rp, (const u32 *)tmp_col, (const VT *)tmp_val, C->d_col, (VT *)C->d_val, llg,
                                                          &ctx->d_ctrl->max_val_out, C->d_maxval, (u64 *)ctx->d_scan, B200_CTRL_BYTES / 8, scan_bytes / 8);
            LAUNCH_CHECK(ctx);
            ctx->scan_clean_bytes = scan_bytes;
        }
        if (timing) cudaEventRecord(ctx->f_ev[slot][2], s);
        trace_dump(ctx, "one-pass multiply");
        mark_pending(ctx, C, slot, epoch, A, B, mode, 2, (int32_t)(ctx->launches - launches0), timing, std::min<u64>(p_bound, ncols));
        *out = C;
        if (st) { TRY(resolve_pending(ctx, C)); *st = *C->stats; }
        return B200_OK;
    }

    return set_err(B200_ERR_CUDA, "internal: no numeric path selected");
}

extern "C" int b200_spgemm(b200_ctx *ctx, const b200_csr *A, const b200_csr *B, b200_csr **C, b200_stats *stats) {
    if (!ctx || !A || !B || !C) return set_err(B200_ERR_BADARG, "NULL argument");
    if (A->cols != B->rows) return set_err(B200_ERR_SHAPE, "shape mismatch: A is %llux%llu, B is %llux%llu", (ull)A->rows, (ull)A->cols, (ull)B->rows, (ull)B->cols);
    if (A->val_bits != B->val_bits) return set_err(B200_ERR_SHAPE, "value width mismatch: A is u%d, B is u%d", A->val_bits, B->val_bits);
    if (A->ctx != ctx || B->ctx != ctx) return set_err(B200_ERR_BADARG, "operand handles belong to another context");
    CUDA_TRY(cudaSetDevice(ctx->device));
    RESOLVE(ctx, A); RESOLVE(ctx, B);
    // Operands that are powers of one base handle commute, and the saturating path-count semiring is associative (every
    // entry of any product of such powers is min(cap, its exact integer value), whatever the order of evaluation): when A's
    // rows are long and B's short -- a power chain step A^(k-1) x A -- the engine evaluates B x A, whose rows are unions of
    // a few long sorted rows (pipeline 6).  Same matrix, bit for bit.
    const u64 lin_base = A->lin_base == B->lin_base ? A->lin_base : 0, lin_pow = A->lin_pow + B->lin_pow;
    if (lin_base && ctx->cfg.commute_swap && (ctx->cfg.pipeline == 0 || ctx->cfg.pipeline == 6) && A != B && A->rows && A->nnz && B->nnz) {
        const double meanA = (double)A->nnz / (double)A->rows, meanB = (double)B->nnz / (double)B->rows;
        if (B->max_row_len <= 32 && meanA >= ctx->lm_min_list && meanA >= 4.0 * meanB) std::swap(A, B);
    }
    bool handled = false;
    TRY(A->val_bits == 32 ? spgemm_fused<u32>(ctx, A, B, C, stats, &handled) : spgemm_fused<u64>(ctx, A, B, C, stats, &handled));
    if (!handled) TRY(A->val_bits == 32 ? spgemm_typed<u32>(ctx, A, B, C, stats) : spgemm_typed<u64>(ctx, A, B, C, stats));
    if (lin_base) { (*C)->lin_base = lin_base; (*C)->lin_pow = lin_pow; }
    return B200_OK;
}

extern "C" int b200_csr_product_stats(b200_ctx *ctx, const b200_csr *C, b200_stats *stats) {
    if (!ctx || !C || !stats) return set_err(B200_ERR_BADARG, "NULL argument");
    RESOLVE(ctx, C);
    if (!C->stats) return set_err(B200_ERR_BADARG, "the handle is not the product of a multiply (or its multiply kept no statistics)");
    *stats = *C->stats;
    return B200_OK;
}

// ---------------------------------------------------------------------------- sharding helpers
extern "C" int b200_row_products(b200_ctx *ctx, const b200_csr *A, const b200_csr *B, uint64_t *host_out) {
    if (!ctx || !A || !B || !host_out) return set_err(B200_ERR_BADARG, "NULL argument");
    if (A->cols != B->rows) return set_err(B200_ERR_SHAPE, "shape mismatch");
    CUDA_TRY(cudaSetDevice(ctx->device));
    RESOLVE(ctx, A); RESOLVE(ctx, B);
    if (A->rows == 0) return B200_OK;
    TRY(ensure_row_scratch(ctx, A->rows));
    TRY(ensure_desc(ctx, B));
    TRY(launch_row_products(ctx, A, B));
    CUDA_TRY(cudaMemcpyAsync(host_out, ctx->d_prod, A->rows * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return B200_OK;
}

// Cost of a row in the balance: its intermediate products, +1 so that empty rows spread over the parts too, and rows of the
// heavy list (more products than the largest shared-memory table takes) count 2.5 x -- measured on R-MAT row blocks, where
// the global-memory heavy-row kernels run at ~0.4 of the hash kernels' products/s and equal product counts left the rank
// holding the hub rows 1.4 x behind the others.
__host__ __device__ __forceinline__ u64 shard_cost(u64 p) { return p + 1 + (p > (u64)b200_hash_cap(B200_NUM_HASH_BINS - 1) ? p + p / 2 : 0); }
#define SHARD_TILE 4096
__global__ void __launch_bounds__(256) k_shard_tile_sums(u64 rows, const u64 *__restrict__ prod, u64 *__restrict__ sums) {
    __shared__ u64 s_w[8];
    const u64 base = (u64)blockIdx.x * SHARD_TILE;
    u64 v = 0;
    for (u32 j = threadIdx.x; j < SHARD_TILE; j += 256) if (base + j < rows) v += shard_cost(prod[base + j]);
    v = warp_sum_u64(v);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) { u64 t = 0; for (int w = 0; w < 8; w++) t += s_w[w]; sums[blockIdx.x] = t; }
}
// Cut k = first index i of the prefix pre[0..rows] (pre[0] = 0, pre[i] = sum_{j<i} cost(P_j)) with pre[i] >= k * total / nparts.
// The per-row counts stay on the device: the host reads one sum per 4096 rows, then the one tile each cut falls into.
extern "C" int b200_shard_rows_by_products(b200_ctx *ctx, const b200_csr *A, const b200_csr *B, int nparts, uint64_t *cuts) {
    if (!ctx || !A || !B || !cuts || nparts < 1) return set_err(B200_ERR_BADARG, "bad nparts/cuts");
    if (A->cols != B->rows) return set_err(B200_ERR_SHAPE, "shape mismatch");
    CUDA_TRY(cudaSetDevice(ctx->device));
    RESOLVE(ctx, A); RESOLVE(ctx, B);
    const u64 rows = A->rows;
    cuts[0] = 0; cuts[nparts] = rows;
    if (rows == 0) { for (int k = 1; k < nparts; k++) cuts[k] = 0; return B200_OK; }
    TRY(ensure_row_scratch(ctx, rows));
    TRY(ensure_desc(ctx, B));
    TRY(launch_row_products(ctx, A, B));
    const u64 tiles = (rows + SHARD_TILE - 1) / SHARD_TILE;
    u64 *d_sums = nullptr;
    TRY(dmalloc(ctx, (void **)&d_sums, tiles * 8));
    k_shard_tile_sums<<<(unsigned)tiles, 256, 0, ctx->stream>>>(rows, ctx->d_prod, d_sums);
    ctx->launches++;
    std::vector<u64> sums(tiles), tile(SHARD_TILE);
    cudaError_t e = cudaMemcpyAsync(sums.data(), d_sums, tiles * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    dfree(ctx, d_sums);
    if (e != cudaSuccess) return set_err(B200_ERR_CUDA, "shard_rows_by_products: %s", cudaGetErrorString(e));
    std::vector<unsigned __int128> pre(tiles + 1, 0);
    for (u64 t = 0; t < tiles; t++) pre[t + 1] = pre[t] + sums[t];
    const unsigned __int128 total = pre[tiles];
    for (int k = 1; k < nparts; k++) {
        const unsigned __int128 target = total * (unsigned)k / (unsigned)nparts;
        u64 lo = 0;
        if (target > 0) {
            // tile t with pre[t] < target <= pre[t + 1]: the cut is inside it (or at its end)
            const u64 t = (u64)(std::lower_bound(pre.begin(), pre.end(), target) - pre.begin()) - 1;
            const u64 base = t * SHARD_TILE, cnt = std::min<u64>(SHARD_TILE, rows - base);
            CUDA_TRY(cudaMemcpyAsync(tile.data(), ctx->d_prod + base, cnt * 8, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            unsigned __int128 run = pre[t];
            u64 i = 0;
            while (i < cnt && run < target) { run += shard_cost(tile[i]); i++; }
            lo = base + i;
        }
        cuts[k] = std::max<uint64_t>(std::min<u64>(lo, rows), cuts[k - 1]);
    }
    return B200_OK;
}
extern "C" int b200_csr_row_block(b200_ctx *ctx, const b200_csr *A, uint64_t r0, uint64_t r1, b200_csr **out) {
    if (!ctx || !A || !out) return set_err(B200_ERR_BADARG, "NULL argument");
    if (r0 > r1 || r1 > A->rows) return set_err(B200_ERR_BADARG, "row range [%llu,%llu) outside 0..%llu", (ull)r0, (ull)r1, (ull)A->rows);
    CUDA_TRY(cudaSetDevice(ctx->device));
    RESOLVE(ctx, A);
    u64 ends[2];
    CUDA_TRY(cudaMemcpyAsync(&ends[0], A->d_rp + r0, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(&ends[1], A->d_rp + r1, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    const u64 nnz = ends[1] - ends[0], rows = r1 - r0;
    b200_csr *m = nullptr;
    TRY(csr_alloc(ctx, rows, A->cols, nnz, A->val_bits, true, &m));
    k_rebase_rowptr<<<grid_for(rows + 1, 256, ctx->num_sms * 8), 256, 0, ctx->stream>>>(rows + 1, A->d_rp + r0, ends[0], m->d_rp);
    ctx->launches++;
    cudaError_t ce = cudaGetLastError();
    if (ce == cudaSuccess && nnz) {
        ce = cudaMemcpyAsync(m->d_col, A->d_col + ends[0], nnz * 4, cudaMemcpyDeviceToDevice, ctx->stream);
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(m->d_val, (const char *)A->d_val + ends[0] * (A->val_bits / 8), nnz * (size_t)(A->val_bits / 8), cudaMemcpyDeviceToDevice, ctx->stream);
    }
    if (ce != cudaSuccess) { b200_csr_free(ctx, m); return set_err(B200_ERR_CUDA, "row_block: %s", cudaGetErrorString(ce)); }
    m->max_row_len = A->max_row_len;
    int r = finish_new_csr(ctx, m, true, true);
    if (r != B200_OK) { b200_csr_free(ctx, m); return r; }
    *out = m;
    return B200_OK;
}

// ---------------------------------------------------------------------------- add / pattern compare
template <typename VT>
static int add_typed(b200_ctx *ctx, const b200_csr *A, const b200_csr *B, b200_csr **out) {
    const u64 rows = A->rows;
    cudaStream_t s = ctx->stream;
    b200_csr *C = nullptr;
    TRY(csr_alloc(ctx, rows, A->cols, 0, A->val_bits, false, &C));
    CUDA_TRY(cudaMemsetAsync(C->d_maxval, 0, 16, s));
    if (rows == 0) { TRY(alloc_entries(ctx, C)); CUDA_TRY(cudaMemsetAsync(C->d_rp, 0, 8, s)); *out = C; return B200_OK; }
    int r = ensure_row_scratch(ctx, rows);
    if (r != B200_OK) { b200_csr_free(ctx, C); return r; }
    CUDA_TRY(reset_scan(ctx, 0));
    const unsigned g = (unsigned)((rows + 255) / 256);
    k_add_rows<VT, false><<<g, 256, 0, s>>>(view<VT>(A), view<VT>(B), ctx->d_nnz_row, nullptr, nullptr, nullptr, ctx->d_ctrl);
    LAUNCH_CHECK(ctx);
    launch_scan_rowptr(ctx, rows, C->d_rp, s, nullptr, 0);
    LAUNCH_CHECK(ctx);
    CUDA_TRY(cudaMemcpyAsync(ctx->h_ctrl, ctx->d_ctrl, sizeof(B200Ctrl), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    C->nnz = ctx->h_ctrl->total_nnz; C->max_row_len = ctx->h_ctrl->max_row_nnz;
    r = alloc_entries(ctx, C);
    if (r != B200_OK) { b200_csr_free(ctx, C); return r; }
    k_add_rows<VT, true><<<g, 256, 0, s>>>(view<VT>(A), view<VT>(B), ctx->d_nnz_row, C->d_rp, C->d_col, (VT *)C->d_val, ctx->d_ctrl);
    LAUNCH_CHECK(ctx);
    CUDA_TRY(cudaMemcpyAsync(C->d_maxval, &ctx->d_ctrl->max_val_out, 8, cudaMemcpyDeviceToDevice, s));
    *out = C;
    return B200_OK;
}

extern "C" int b200_csr_add(b200_ctx *ctx, const b200_csr *A, const b200_csr *B, b200_csr **C) {
    if (!ctx || !A || !B || !C) return set_err(B200_ERR_BADARG, "NULL argument");
    if (A->rows != B->rows || A->cols != B->cols) return set_err(B200_ERR_SHAPE, "add: shape mismatch");
    if (A->val_bits != B->val_bits) return set_err(B200_ERR_SHAPE, "add: value width mismatch");
    if (A->ctx != ctx || B->ctx != ctx) return set_err(B200_ERR_BADARG, "operand handles belong to another context");
    CUDA_TRY(cudaSetDevice(ctx->device));
    RESOLVE(ctx, A); RESOLVE(ctx, B);
    return A->val_bits == 32 ? add_typed<u32>(ctx, A, B, C) : add_typed<u64>(ctx, A, B, C);
}

extern "C" int b200_csr_same_pattern(b200_ctx *ctx, const b200_csr *A, const b200_csr *B, int *same) {
    if (ctx && A && B && (A->ctx != ctx || B->ctx != ctx)) return set_err(B200_ERR_BADARG, "operand handles belong to another context");
    if (!ctx || !A || !B || !same) return set_err(B200_ERR_BADARG, "NULL argument");
    RESOLVE(ctx, A); RESOLVE(ctx, B);
    if (A->rows != B->rows || A->cols != B->cols || A->nnz != B->nnz) { *same = 0; return B200_OK; }
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaMemsetAsync(ctx->d_flag, 0, 64, ctx->stream));
    k_compare_u64<<<grid_for(A->rows + 1, 256, ctx->num_sms * 8), 256, 0, ctx->stream>>>(A->rows + 1, A->d_rp, B->d_rp, ctx->d_flag);
    LAUNCH_CHECK(ctx);
    if (A->nnz) { k_compare_u32<<<grid_for(A->nnz, 256, ctx->num_sms * 8), 256, 0, ctx->stream>>>(A->nnz, A->d_col, B->d_col, ctx->d_flag); LAUNCH_CHECK(ctx); }
    CUDA_TRY(cudaMemcpyAsync(ctx->h_flag, ctx->d_flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    *same = ctx->h_flag[0] == 0;
    return B200_OK;
}

// ---------------------------------------------------------------------------- fixture generators on the device (gen.cuh)
// exclusive u64 prefix of ctx->d_nnz_row[0..rows) into `out` (rows + 1 words); returns total and the longest row
int scan_row_counts(b200_ctx *ctx, u64 rows, u64 *out, u64 *total, u64 *max_len) {
    CUDA_TRY(reset_scan(ctx, 0));
    launch_scan_rowptr(ctx, rows, out, ctx->stream, nullptr, 0);
    LAUNCH_CHECK(ctx);
    CUDA_TRY(cudaMemcpyAsync(ctx->h_ctrl, ctx->d_ctrl, sizeof(B200Ctrl), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    *total = ctx->h_ctrl->total_nnz; *max_len = ctx->h_ctrl->max_row_nnz;
    return B200_OK;
}

extern "C" int b200_lattice(b200_ctx *ctx, const uint64_t *dims, int ndims, int torus, int val_bits, b200_csr **out) {
    if (!ctx || !out || (ndims > 0 && !dims)) return set_err(B200_ERR_BADARG, "NULL argument");
    if (ndims < 0 || ndims > B200_LATTICE_MAXD) return set_err(B200_ERR_BADARG, "lattice: 0..%d dimensions supported, got %d", B200_LATTICE_MAXD, ndims);
    if (val_bits != 32 && val_bits != 64) return set_err(B200_ERR_BADARG, "val_bits must be 32 or 64");
    CUDA_TRY(cudaSetDevice(ctx->device));
    LatticeDims L; memset(&L, 0, sizeof(L)); L.nd = ndims; L.torus = torus ? 1 : 0;
    unsigned __int128 total128 = 1;
    for (int d = 0; d < ndims; d++) { if (!dims[d]) return set_err(B200_ERR_BADARG, "lattice: dimension %d is 0", d); L.dim[d] = dims[d]; total128 *= dims[d]; }
    if (total128 > 0xFFFFFFFFull) return set_err(B200_ERR_BADARG, "lattice: more than 2^32-1 nodes (NodeId is u32)");
    const u64 total = (u64)total128;
    for (int d = ndims - 1; d >= 0; d--) L.stride[d] = d == ndims - 1 ? 1 : L.stride[d + 1] * L.dim[d + 1];   // last dimension fastest
    b200_csr *C = nullptr;
    TRY(csr_alloc(ctx, total, total, 0, val_bits, false, &C));
    int r = ensure_row_scratch(ctx, total);
    if (r != B200_OK) { b200_csr_free(ctx, C); return r; }
    const unsigned g = (unsigned)((total + 127) / 128);
    cudaStream_t s = ctx->stream;
    if (val_bits == 32) k_lattice<u32, false><<<g, 128, 0, s>>>(total, L, ctx->d_nnz_row, nullptr, nullptr, nullptr);
    else k_lattice<u64, false><<<g, 128, 0, s>>>(total, L, ctx->d_nnz_row, nullptr, nullptr, nullptr);
    LAUNCH_CHECK(ctx);
    u64 nnz = 0, maxlen = 0;
    r = scan_row_counts(ctx, total, C->d_rp, &nnz, &maxlen);
    if (r != B200_OK) { b200_csr_free(ctx, C); return r; }
    C->nnz = nnz; C->max_row_len = maxlen;
    r = alloc_entries(ctx, C);
    if (r != B200_OK) { b200_csr_free(ctx, C); return r; }
    if (nnz) {
        if (val_bits == 32) k_lattice<u32, true><<<g, 128, 0, s>>>(total, L, nullptr, C->d_rp, C->d_col, (u32 *)C->d_val);
        else k_lattice<u64, true><<<g, 128, 0, s>>>(total, L, nullptr, C->d_rp, C->d_col, (u64 *)C->d_val);
        LAUNCH_CHECK(ctx);
    }
    r = finish_new_csr(ctx, C, true, false);
    if (r != B200_OK) { b200_csr_free(ctx, C); return r; }
    *out = C;
    return B200_OK;
}

template <typename VT>
static int thin_typed(b200_ctx *ctx, const b200_csr *A, double density, const ChaChaKey &key, u64 skip, b200_csr **out, u64 *draws) {
    const u64 rows = A->rows;
    cudaStream_t s = ctx->stream;
    b200_csr *C = nullptr;
    TRY(csr_alloc(ctx, rows, A->cols, 0, A->val_bits, false, &C));
    int r = ensure_row_scratch(ctx, rows);
    u64 *base_u = nullptr;
    if (r == B200_OK) r = dmalloc(ctx, (void **)&base_u, (rows + 1) * 8);
    if (r != B200_OK) { b200_csr_free(ctx, C); return r; }
    const unsigned g = (unsigned)((rows + 255) / 256);
    u64 nu = 0, nnz = 0, maxlen = 0;
    k_thin_upper_count<<<g, 256, 0, s>>>(rows, A->d_rp, A->d_col, ctx->d_nnz_row);
    LAUNCH_CHECK(ctx);
    r = scan_row_counts(ctx, rows, base_u, &nu, &maxlen);                  // draw number of every row's first upper entry
    if (r == B200_OK) {
        k_thin_rows<VT, false><<<g, 256, 0, s>>>(rows, A->d_rp, A->d_col, (const VT *)A->d_val, base_u, key, skip, density, ctx->d_nnz_row, nullptr, nullptr, nullptr);
        LAUNCH_CHECK(ctx);
        r = scan_row_counts(ctx, rows, C->d_rp, &nnz, &maxlen);
    }
    if (r == B200_OK) { C->nnz = nnz; C->max_row_len = maxlen; r = alloc_entries(ctx, C); }
    if (r == B200_OK && nnz) {
        k_thin_rows<VT, true><<<g, 256, 0, s>>>(rows, A->d_rp, A->d_col, (const VT *)A->d_val, base_u, key, skip, density, nullptr, C->d_rp, C->d_col, (VT *)C->d_val);
        LAUNCH_CHECK(ctx);
    }
    dfree(ctx, base_u);
    if (r == B200_OK) r = finish_new_csr(ctx, C, true, false);
    if (r != B200_OK) { b200_csr_free(ctx, C); return r; }
    if (draws) *draws = nu;
    *out = C;
    return B200_OK;
}

extern "C" int b200_thin(b200_ctx *ctx, const b200_csr *A, double density, const uint8_t *seed32, uint64_t skip_draws, b200_csr **out,
                         uint64_t *draws_consumed) {
    if (!ctx || !A || !seed32 || !out) return set_err(B200_ERR_BADARG, "NULL argument");
    if (A->rows != A->cols) return set_err(B200_ERR_SHAPE, "thin: the matrix must be square (%llux%llu)", (ull)A->rows, (ull)A->cols);
    CUDA_TRY(cudaSetDevice(ctx->device));
    RESOLVE(ctx, A);
    ChaChaKey key;
    for (int i = 0; i < 8; i++) key.k[i] = (u32)seed32[4 * i] | ((u32)seed32[4 * i + 1] << 8) | ((u32)seed32[4 * i + 2] << 16) | ((u32)seed32[4 * i + 3] << 24);
    if (A->rows == 0) { TRY(csr_alloc(ctx, 0, 0, 0, A->val_bits, true, out)); CUDA_TRY(cudaMemsetAsync((*out)->d_rp, 0, 8 + 16, ctx->stream)); if (draws_consumed) *draws_consumed = 0; return B200_OK; }
    return A->val_bits == 32 ? thin_typed<u32>(ctx, A, density, key, skip_draws, out, draws_consumed)
                             : thin_typed<u64>(ctx, A, density, key, skip_draws, out, draws_consumed);
}

// Host twin of the generator the device kernels use (no GPU needed): StdRng::from_seed(seed).next_u64() numbers
// first .. first+n-1.  Lets a CPU-only test pin the device algorithm's ChaCha12 against the oracle.
extern "C" int b200_stdrng_u64(const uint8_t *seed32, uint64_t first, uint64_t n, uint64_t *out) {
    if (!seed32 || (n && !out)) return set_err(B200_ERR_BADARG, "NULL argument");
    ChaChaKey key;
    for (int i = 0; i < 8; i++) key.k[i] = (u32)seed32[4 * i] | ((u32)seed32[4 * i + 1] << 8) | ((u32)seed32[4 * i + 2] << 16) | ((u32)seed32[4 * i + 3] << 24);
    for (u64 i = 0; i < n; i++) out[i] = stdrng_u64_at(key, first + i);
    return B200_OK;
}
