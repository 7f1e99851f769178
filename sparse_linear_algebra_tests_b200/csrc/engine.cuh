// engine.cuh -- host-side state shared by the engine's translation units (api.cu, fused.cu, heavy.cu, coo.cu, comm.cu):
// the two handle types behind the C ABI, error plumbing, allocation helpers and the per-operand caches.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdarg>
#include <string>
#include <vector>
#include <algorithm>

#include "../../include/b200_spgemm.h"
#include "common.cuh"

#define B200_NAUX 3
#define B200_REPORT_SLOTS 16      // multiplies whose device report may be outstanding at once (fused path)

// ---------------------------------------------------------------------------- error plumbing (api.cu)
int set_err(int code, const char *fmt, ...);
#define CUDA_TRY(expr)                                                                             \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            return set_err(_e == cudaErrorMemoryAllocation ? B200_ERR_ALLOC : B200_ERR_CUDA,       \
                           "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)
#define TRY(expr) do { int _r = (expr); if (_r != B200_OK) return _r; } while (0)
// ---------------------------------------------------------------------------- handles
struct b200_csr {
    u64 rows, cols, nnz;
    int val_bits;
    u64 *d_rp; u32 *d_col; void *d_val;
    ull *d_maxval;          // device scalar: largest stored value (lives behind row_ptr, same allocation)
    bool val_shares_col;    // values live in the col_idx allocation (products: one allocation per multiply)
    u64 max_row_len;        // host-known upper bound of the longest row
    uint2 *d_desc;          // {start,len} per row, built lazily when used as a right operand
    uint4 *d_span;          // {len, first col, last col, -} per row, built with d_desc (pre-pass: plain column windows)
    uint4 *d_cspan;         // square operands: {len, min, max} of (col - row + n/2) mod n (pre-pass: circular windows)
    uint4 *d_pack;          // sector-packed rows (low-degree right operands), built lazily
    u64 h_maxval; bool h_maxval_known;   // host copy of *d_maxval once it has been read back
    // circular column range: every stored column is (cr_start + o) mod cols for some o < cr_len (cr_len = cols: unknown / everything)
    u32 cr_start; u64 cr_len;
    // square operands used on the right: signed offsets (c - k) of all entries lie in [cs_lo, cs_hi]; cs_state 0 unknown, 1 known, 2 none
    long long cs_lo, cs_hi; int cs_state;
    // active row arc: every non-empty row is (ar_start + o) mod rows for some o < ar_len (ar_len = rows: unknown / everything).
    // Known for host uploads (row_ptr passes through the host) and inherited by products from their left operand: a rank of a
    // sharded power chain holds its block of rows plus a halo inside whole-size matrices, the other rows are empty.
    u64 ar_start, ar_len;
    cudaEvent_t ev_copy;    // last asynchronous download of this handle on the copy stream (created on first use)
    b200_ctx *ctx;
    // ---- fused path (fused.cu): a product is returned before the host knows its size
    size_t entry_bytes;     // bytes of the col/val allocation (entry cache)
    u64 cap_entries;        // entries the col/val allocation holds (>= nnz; products are allocated from a bound)
    u64 max_row_span;       // every row's columns lie on an arc of the index circle of at most this many + 1 columns (cols: unknown)
    int pending_slot;       // >= 0: nnz / max_row_len / h_maxval arrive with report slot `pending_slot` (see resolve_pending)
    u32 pending_epoch;
    u64 est_nnz;            // while pending: host-side estimate of nnz (heuristics of a multiply queued behind this one)
    b200_stats *stats;      // measurements of the multiply that produced this handle (filled when the report is read)
    bool stats_timed;       // the report slot's events were recorded for it
    // lineage: this handle holds (base handle `lin_base`) ^ lin_pow.  Every handle is its own base to the power 1; a product
    // of two powers of one base is a power of that base.  Powers of one matrix commute (b200_spgemm may swap them).
    u64 uid, lin_base, lin_pow;
};

struct b200_ctx {
    int device, num_sms;
    size_t smem_optin, total_mem;
    cudaStream_t stream; bool own_stream;
    B200Ctrl *d_ctrl, *h_ctrl;
    u64 *h_report;          // pinned: the final scan's report, one {word, epoch} chunk per 32-bit word of the control block
    // per-row scratch, grown on demand
    u64 cap_rows; u64 *d_prod; u64 *d_tmp_ptr; u32 *d_nnz_row; u32 *d_bin_rows; uint4 *d_win;
    // one buffer, one memset per multiply: control block | status of the row_ptr scan | status of the pre-pass scan
    unsigned char *d_scan; u64 *d_tile_status, *d_tile_pre; u64 cap_tiles, cap_tiles_pre;
    size_t scan_clean_bytes;   // leading bytes of d_scan known to be zero on the stream (left so by k_compact_rows)
    // heavy-row scratch
    void *d_heavy; size_t cap_heavy;
    // one-pass scratch CSR (bound-offset rows), kept across multiplies so the steady state allocates nothing
    void *d_tmp_col, *d_tmp_val; size_t cap_tmp_col, cap_tmp_val;
    u32 *d_flag;            // small device flag word (+ pinned mirror)
    u32 *h_flag;
    cudaEvent_t ev[4];
    cudaStream_t aux[B200_NAUX]; cudaEvent_t ev_fork, ev_join[B200_NAUX]; int naux_enabled;
    cudaStream_t copy; cudaEvent_t ev_ready;   // D2H stream: downloads overlap the next multiply
    bool timing;
    u32 epoch;              // multiplies reported through the pinned mirror so far
    u64 launches;
    // developer timeline (B200_TRACE=1): an event after every launch, on the stream it went to
    bool trace; cudaStream_t cur_stream;
    bool hosttime; double ht[8];   // B200_HOSTTIME=1: host clock at the phase boundaries of a multiply (dev tool)
    std::vector<std::pair<int, cudaEvent_t>> *marks;
    // ---- fused path (fused.cu): pre-pass -> one persistent numeric kernel that also places the rows of C
    unsigned char *d_fz;    // one self-cleaning buffer: control block | pre-pass tile status | unit status
    B200Ctrl *d_fctrl; u64 *d_ftile, *d_fustat; u64 cap_ftile, cap_fustat;
    u32 *d_units; uint4 *d_rowwin; unsigned char *d_rowclass; u64 cap_frows;
    u64 *d_spill_acc; u32 *d_spill_col; u64 cap_spill;   // per-CTA global slots for the upper ranks of rows longer than the shared-memory slots
    bool f_dirty;           // the self-cleaning buffer must be zeroed before its next use (first use, failed multiply)
    u64 *h_freport;         // pinned ring of B200_REPORT_SLOTS reports, {word, epoch} chunks
    b200_csr *slot_owner[B200_REPORT_SLOTS];
    cudaEvent_t f_ev[B200_REPORT_SLOTS][3];
    u32 fepoch;
    u64 *d_lm_tot; bool lm_tot_dirty;  // per-slice totals of the left multiply (leftmul.cu), left zeroed by its last CTA
    u64 *d_cta_tot; u64 cap_cta_tot;   // per-CTA totals of the one-launch multiply (rowwarp.cu)
    // entry arrays (col_idx | values, one allocation) of freed handles, kept for the next product of about that size: a loop
    // of large multiplies (12-GB results on the 200^3 chain) otherwise stalls in the stream-ordered allocator every few steps
    std::vector<std::pair<void *, size_t>> *entry_cache; size_t entry_cache_bytes;
    // narrow downloads: u64 values proven < 2^32 cross PCIe as u32 and are widened by host threads into the caller's array
    cudaStream_t post;                  // host callbacks (the widening) run behind the copy stream's chunks
    struct NarrowState *narrow;
    u64 *d_rowstat; u64 cap_rowstat;   // per-row look-back status of the one-pass multiply (dense.cu)
    void *d_hv; size_t cap_hv;         // chunked heavy-row kernels (heavy.cu): control words | per-(row, chunk) counters | unit lists
    double lm_min_list;     // left multiply (leftmul.cu): shortest mean list (row of B) it is chosen for (48; developer override B200_LM_MINLIST)
    b200_config cfg;        // tuning switches (b200_ctx_configure; the MagnusConfig analogue)
};

// resolve a handle whose size is still on its way from the device (no-op for every other handle)
int resolve_pending(b200_ctx *ctx, const b200_csr *m);
static inline int resolve_if_pending(b200_ctx *c, const b200_csr *m) { return m && m->pending_slot >= 0 ? resolve_pending(c ? c : m->ctx, m) : B200_OK; }
#define RESOLVE(c_, m_) TRY(resolve_if_pending((c_), (m_)))

static inline double host_now_us() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec * 1e6 + t.tv_nsec * 1e-3; }
void trace_mark(b200_ctx *ctx, int line);
void trace_dump(b200_ctx *ctx, const char *what);

#define LAUNCH_CHECK(ctx)                                                                          \
    do { (ctx)->launches++; if ((ctx)->trace) trace_mark(ctx, __LINE__); cudaError_t _e = cudaGetLastError(); \
         if (_e != cudaSuccess) return set_err(B200_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); } while (0)

int dmalloc(b200_ctx *ctx, void **p, size_t bytes);
void dfree(b200_ctx *ctx, void *p);

// Independent per-bin kernels are spread over the main stream and a few auxiliary streams.
struct Fan {
    b200_ctx *ctx; int next; bool used[B200_NAUX]; bool forked;
    explicit Fan(b200_ctx *c) : ctx(c), next(0), forked(false) { for (int i = 0; i < B200_NAUX; i++) used[i] = false; }
    cudaStream_t pick() {
        const int n = ctx->naux_enabled;
        if (n == 0) return ctx->stream;
        const int slot = next++ % (n + 1);
        if (slot == n) return ctx->stream;
        ctx->cur_stream = ctx->aux[slot];
        if (!forked) { cudaEventRecord(ctx->ev_fork, ctx->stream); forked = true; }
        if (!used[slot]) { cudaStreamWaitEvent(ctx->aux[slot], ctx->ev_fork, 0); used[slot] = true; }
        return ctx->aux[slot];
    }
    void join() {
        for (int i = 0; i < B200_NAUX; i++)
            if (used[i]) { cudaEventRecord(ctx->ev_join[i], ctx->aux[i]); cudaStreamWaitEvent(ctx->stream, ctx->ev_join[i], 0); used[i] = false; }
        forked = false; next = 0;
    }
};


// ---- helpers of api.cu that the other translation units use
int ensure_row_scratch(b200_ctx *ctx, u64 rows);
int ensure_desc(b200_ctx *ctx, const b200_csr *B);
int ensure_pack(b200_ctx *ctx, const b200_csr *B);
int ensure_cs_bounds(b200_ctx *ctx, const b200_csr *B);
bool want_pack(const b200_ctx *ctx, const b200_csr *B);
int pick_lg(const b200_ctx *ctx, const b200_csr *B, int max_lg);
int pick_mode_bits(const b200_ctx *ctx, int val_bits, u64 max_row_products, u64 maxA, u64 maxB);
int host_maxval(b200_ctx *ctx, const b200_csr *m);
int csr_alloc(b200_ctx *ctx, u64 rows, u64 cols, u64 nnz, int val_bits, bool alloc_arrays, b200_csr **out);
int alloc_entries(b200_ctx *ctx, b200_csr *m);
// value max + format check of a new handle (synchronises when `check`); row counts in ctx->d_nnz_row -> row_ptr (synchronises)
int finish_new_csr(b200_ctx *ctx, b200_csr *m, bool check, bool device_rowptr = false);
int scan_row_counts(b200_ctx *ctx, u64 rows, u64 *out, u64 *total, u64 *max_len);
// count / numeric kernels of the binned pipeline over the hash ("wide") and heavy lists only, writing C at its final offsets
int legacy_counts(b200_ctx *ctx, const b200_csr *A, const b200_csr *B, B200Ctrl *ctrl, u64 p_bound, int lg, Fan &fan);
int legacy_numeric(b200_ctx *ctx, const b200_csr *A, const b200_csr *B, B200Ctrl *ctrl, b200_csr *C, u64 p_bound, u64 heavy_cap, int mode,
                   bool packed, bool bpat, int lg, Fan &fan);
// ---- rowwarp.cu: row-per-warp count / numeric kernels of the exact placement
#define B200_RW_MAX_HB 7          // hash bins 0..7 (33..8192 intermediate products) are rows a single warp produces
#define B200_RW_MAX_GROUPS 512    // largest window bitmap of a warp, in 128-column groups (window offsets are kept as u16)
void rw_setup(b200_ctx *ctx);
size_t rw_smem_per_warp(bool count, int mode, u32 nw, u32 cap);
int rw_launch(b200_ctx *ctx, const b200_csr *A, const b200_csr *B, B200Ctrl *ctrl, int first_bin, int nbins, bool count, int mode,
              bool packed, bool bpat, u32 nw, u32 cap, b200_csr *C, cudaStream_t s);
void rwf_setup(b200_ctx *ctx);
int rw_fused_launch(b200_ctx *ctx, const b200_csr *A, const b200_csr *B, b200_csr *C, B200Ctrl *ctrl, int mode, bool packed, bool bpat,
                    u32 org, u32 words, u32 nw, u32 cap, u64 *mirror, u32 epoch, cudaStream_t s);
// ---- heavy.cu: column-chunked kernels for the heaviest rows (TMA-staged B-row segments, dense accumulator per chunk)
struct HvPlan {
    bool on;
    u32 W, nchunks, sw, nchunks_c, cap_li;   // numeric chunk columns, chunks, numeric chunks per count chunk, count chunks, list rows covered
    u32 heavy_from;                          // the pre-pass puts rows with at least this many products on the heavy list
    u64 pmin, psplit;                        // rows with at least pmin intermediate products are taken; products per numeric work unit
    void *ctl, *cnt, *units_c, *units_n;     // device: control words, per-(row, chunk) counters, work-unit lists
};
void hv_setup(b200_ctx *ctx);
int hv_plan(b200_ctx *ctx, const b200_csr *A, const b200_csr *B, int mode, u64 p_bound, HvPlan *plan);
int hv_count(b200_ctx *ctx, const b200_csr *A, const b200_csr *B, B200Ctrl *ctrl, const HvPlan &p, cudaStream_t s);
int hv_numeric(b200_ctx *ctx, const b200_csr *A, const b200_csr *B, B200Ctrl *ctrl, const HvPlan &p, int mode, bool bpat,
               const u64 *base, u32 *col, void *val, u32 narrow, cudaStream_t s);
// ---- dense.cu: the multiply as one pass over the products (dense window accumulators, look-back placement over rows)
void dn_setup(b200_ctx *ctx);
int dn_launch(b200_ctx *ctx, const b200_csr *A, const b200_csr *B, b200_csr *C, B200Ctrl *ctrl, bool bpat, int ctas_per_sm, u64 *mirror, u32 epoch,
              cudaStream_t s);
// ---- leftmul.cu: short rows in A, long rows in B -- one cooperative launch, B's rows streamed as contiguous lists
void lm_setup(b200_ctx *ctx);
size_t lm_smem_per_warp(int mode, u32 nw, u32 cap);
int lm_launch(b200_ctx *ctx, const b200_csr *A, const b200_csr *B, b200_csr *C, B200Ctrl *ctrl, int mode, u32 org, bool per_row, u32 nw, u32 cap,
              int tune, u64 *mirror, u32 epoch, cudaStream_t s);
// ---- fused.cu
template <typename VT>
int spgemm_fused(b200_ctx *ctx, const b200_csr *A, const b200_csr *B, b200_csr **out, b200_stats *st_out, bool *handled);
int fz_row_span(b200_ctx *ctx, b200_csr *m);
int claim_report_slot(b200_ctx *ctx, u32 *epoch, int *slot, u64 **mirror);
void mark_pending(b200_ctx *ctx, b200_csr *C, int slot, u32 epoch, const b200_csr *A, const b200_csr *B, int mode, int pipeline,
                  int32_t launches, bool timed, u64 max_row_len_bound);
void fz_setup(b200_ctx *ctx);
