/*
 * b200_spgemm.h -- C ABI of the B200-native SpGEMM engine (libb200spgemm.so / .a).
 *
 * Drop-in boundary for the ONE hot path of imlvts/sparse-linear-algebra-tests:
 * CSR x CSR SpGEMM over saturating unsigned path-count matrices.  The reference has
 * no FFI of its own (pure Rust); each entry point below names the reference item it
 * replaces (paths relative to the reference root).  A Rust `graph_b200.rs` shim binds
 * these with `extern "C"` (see INTEGRATION.md and rust/).
 *
 * Conventions
 *   - plain pointers and sizes only; every call returns an int status (B200_OK == 0);
 *     b200_last_error() gives the text of the last failure on the calling thread.
 *   - the reference panics (assert_eq!) on shape mismatch (src/graph_csr.rs:307,351;
 *     src/graph_magnus.rs:226,236); the ABI returns B200_ERR_SHAPE and the host shim
 *     turns that back into a panic / exception.
 *   - a b200_ctx is bound to one CUDA device and one stream and must be used by one
 *     host thread at a time.  There is NO CPU fallback: without a usable sm_100 device
 *     b200_ctx_create fails with B200_ERR_CUDA.
 *   - matrices are device-resident handles (b200_csr).  Layout = the reference's CSR
 *     (src/graph_csr.rs:42-53): row_ptr u64[rows+1] (Rust usize), col_idx u32[nnz]
 *     sorted and unique per row, values u32[nnz] (graph_csr.rs:17 `Val`) or u64[nnz]
 *     (graph_sprs.rs:16 `Sat64`, linalg/src/csr.rs:75 `Value for u64`).
 *   - explicit zeros are not representable in the reference's constructors
 *     (from_coo drops them, graph_csr.rs:108); upload rejects them.
 */
#ifndef B200_SPGEMM_H
#define B200_SPGEMM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_OK          0
#define B200_ERR_BADARG  1
#define B200_ERR_SHAPE   2   /* operand shapes / value widths do not agree            */
#define B200_ERR_ALLOC   3   /* device or pinned-host allocation failed               */
#define B200_ERR_CUDA    4   /* no device, launch failure, or any other CUDA error    */
#define B200_ERR_FORMAT  5   /* CSR invariant violated (explicit zero, col >= cols, unsorted or repeated column in a row) */
#define B200_ERR_NCCL    6   /* NCCL missing or a collective failed                   */

typedef struct b200_ctx b200_ctx;
typedef struct b200_csr b200_csr;
typedef struct b200_comm b200_comm;   /* one rank (one GPU) of a multi-GPU job: an NCCL communicator bound to a b200_ctx */

/* Per-multiply measurements; device times are CUDA-event times on the ctx stream. */
typedef struct b200_stats {
    uint64_t rows, cols, nnz_a, nnz_b, nnz_c;
    uint64_t products;          /* intermediate products = sum over (i,k) in A of nnz(B[k,:]) */
    uint64_t max_row_products;
    uint64_t max_row_nnz;
    uint64_t bytes_algorithmic; /* (nnzA+nnzB+nnzC)*(4+sizeof(Val)) + (rowsA+rowsB+rowsC+3)*8 */
    float    ms_symbolic;       /* exact mode: pre-pass + count kernels + row_ptr scan;
                                   scratch mode: compaction of the scratch rows into C                          */
    float    ms_numeric;        /* exact mode: numeric kernels writing C;
                                   scratch mode: pre-pass + numeric kernels + row_ptr scan                      */
    float    ms_total;          /* first kernel to last kernel, incl. the host's read of nnz in the middle      */
    int32_t  acc_mode;          /* 0: 32-bit accumulators proved safe, 1: 64-bit, 2: saturating                 */
    int32_t  kernel_launches;   /* kernels launched by this multiply                                            */
    uint32_t sym_bin_rows[16];  /* binned pipeline: rows per pre-pass list: [0] tiny, [1..8] bitmap bins (P <= 64<<hb; 1
                                   and 2 share list 2), [9] heavy, [10..15] hash lists of bins 0..5 (window too wide).
                                   fused pipeline: [0] tiny, [1] dense, [2] other (counted beforehand), [3] empty rows,
                                   [9] heavy and [10..15] hash lists of the "other" rows                          */
    uint32_t pipeline;          /* pipeline that produced the result: 1 fused, 2 binned, 3 row-per-warp, 4 one launch, 5 one pass, 6 left multiply */
    uint32_t reserved[15];
} b200_stats;

/* Tuning switches of a context: the analogue of MagnusConfig::default() (src/graph_magnus.rs:227,237), passed through
 * the ABI instead of being read from the environment.  b200_config_default fills every field; b200_ctx_configure
 * installs a copy (between multiplies).  -1 / 0 mean "engine decides" where noted. */
typedef struct b200_config {
    uint32_t struct_bytes;       /* sizeof(b200_config) as the caller compiled it (checked)                               */
    int32_t pipeline;            /* 0 auto (4 where it applies, else 2); 1 fused look-back pipeline (round-2 experiment, kept
                                    selectable); 2 binned (a kernel per row bin, scratch or exact placement); 3 exact
                                    placement with the row-per-warp count / numeric kernels over the bin lists; 4 the whole
                                    multiply as ONE cooperative launch (count, placement, numeric; C written once, no
                                    host wait) -- needs one window for all rows that fits a warp's bitmap; 5 one pass over
                                    the products (dense window accumulators per row, look-back placement over rows; needs
                                    32-bit sums, a square low-degree right operand; measured slower, kept selectable);
                                    6 left multiply: short rows in A, long rows in B -- a row of C is the union of a few long
                                    sorted rows of B streamed with coalesced loads, one cooperative launch (auto where it
                                    applies)                                                                               */
    int32_t placement;           /* binned pipeline: -1 auto, 0 scratch CSR + compaction, 1 exact (count pass first)       */
    int32_t exact_limit_mb;      /* binned, auto placement: scratch bound above which the exact placement runs; -1 auto    */
    int32_t force_acc_mode;      /* -1 auto (proved from the operands); 1 / 2 force the 64-bit / saturating accumulators   */
    int32_t window_cap_groups;   /* binned: clamp of a bin's bitmap window in 128-column groups (0: hash kernels only); -1 */
    int32_t window_mul;          /* binned: window of a bin = window_mul * 128 * capacity columns (default 3)              */
    int32_t circular_windows;    /* per-row windows measured on the index circle for square right operands (default 1)     */
    int32_t arc_window;          /* operand-level arc windows (default 1)                                                  */
    int32_t touched_span;        /* binned: walk only the touched part of wide bitmaps (default 1)                         */
    int32_t narrow_scratch;      /* binned: u64 values cross the scratch CSR as u32 when 32-bit sums are proved (1)        */
    int32_t expand_kernel;       /* binned: 0 sends every medium row through hash + sort (default 1)                       */
    int32_t pack_b;              /* sector-packed right operand: -1 auto (mean row <= 4), 0 off, 1 on                      */
    int32_t lanes_per_entry_lg;  /* lanes cooperating on one A entry, log2; -1 auto                                        */
    int32_t expand_div, hash_div, grid_div, grid_mul;   /* binned: thread / grid sizing divisors (8, 32, 8, 4)            */
    int32_t aux_streams;         /* auxiliary streams the per-bin kernels fan out over (default 1, at most 3)              */
    int32_t fused_threads;       /* fused: threads per CTA of the numeric kernel (0 auto = 256; multiple of 32, <= 256);
                                    pipeline 5: CTAs per SM the window is cut for (1..8, default 3); pipeline 6: 4 / 5 / 6
                                    pick the kernel build for that many CTAs per SM (0 auto)                                 */
    int32_t fused_window_cols;   /* fused: preferred cap of the dense accumulator window in columns (0 auto)               */
    int32_t fused_dense_pmax;    /* fused: rows with more intermediate products go to the heavy kernel (0 auto = 65536)    */
    int32_t heavy_chunk_cols;    /* heavy rows: columns per chunk of the chunked kernel (0 auto, from shared memory)       */
    int32_t heavy_kernel;        /* heavy rows: 1 chunked TMA kernel (default), 0 single-CTA global-memory table          */
    int32_t fused_ring_slots;    /* fused: accumulator slots per CTA holding finished rows until they are placed (0 auto)  */
    int32_t fused_product_slots; /* fused: product buffer entries per CTA (rows with more generate their products twice)   */
    int32_t rw_cap_percent;      /* row-per-warp kernels: accumulator slots per warp as a percentage of the mean row's
                                    intermediate products (0 auto = 140); longer rows are produced in several passes       */
    int32_t heavy_min_products;  /* heavy rows: rows with at least this many intermediate products take the chunked kernel
                                    (0 auto: 1024 per column chunk, at least 8193)                                         */
    int32_t heavy_unit_products; /* heavy rows: intermediate products per work unit of the chunked kernel -- rows with more are
                                    cut into several units of consecutive chunks, one CTA each (0 auto = 2^20)              */
    int32_t narrow_download;     /* b200_csr_download_async: u64 values whose maximum is known to be < 2^32 cross PCIe as u32 (a
                                    third fewer bytes) and host threads widen them into the caller's array.  Default 0: on a
                                    16-core host the widening (8 threads, streaming stores) costs what the bus saves (5.42 vs
                                    5.53 ms per A^2..A^7 step); worth switching on where host cores are plentiful           */
    int32_t commute_swap;        /* operands known to commute (both are powers of one base handle, e.g. the steps of a power
                                    chain): evaluate B x A instead of A x B when A's rows are long and B's short -- the result is
                                    the same matrix bit for bit (the saturating path-count semiring is associative), but a row
                                    of it is then the union of a few long sorted rows (pipeline 6).  Default 1 (30^3 chain on B200: A^7 0.316 -> 0.25 ms)            */
    int32_t reserved[1];
} b200_config;
int b200_config_default(b200_config *cfg);

const char *b200_last_error(void);
/* Number of CUDA devices visible (0 when there is none or no driver). */
int b200_device_count(void);

/* Context lifecycle.  `cuda_stream` may be NULL (the engine creates its own stream) or a
 * cudaStream_t owned by the caller (e.g. torch's current stream) on `device`. */
int b200_ctx_create(int device, void *cuda_stream, b200_ctx **out);
int b200_ctx_destroy(b200_ctx *ctx);
int b200_ctx_synchronize(b200_ctx *ctx);
/* Total kernels launched through this context so far (bench.py's gpu_launches). */
int b200_ctx_kernel_launches(b200_ctx *ctx, uint64_t *out);
/* Per-phase CUDA-event timing in b200_stats costs a few event records; on by default. */
int b200_ctx_set_timing(b200_ctx *ctx, int enabled);
int b200_ctx_configure(b200_ctx *ctx, const b200_config *cfg);
int b200_ctx_get_config(b200_ctx *ctx, b200_config *cfg);

/* Host CSR -> device handle.  Replaces constructing a CsrMatrix / MagnusMatrix from its three
 * arrays (src/graph_csr.rs:42-53; SparseMatrixCSR::new at src/graph_magnus.rs:74).
 * val_bits is 32 or 64.  rows x cols may be rectangular so that row blocks of a left operand
 * can be held per GPU; the reference itself only multiplies square matrices. */
int b200_csr_upload(b200_ctx *ctx, uint64_t rows, uint64_t cols, const uint64_t *row_ptr,
                    const uint32_t *col_idx, const void *values, int val_bits, b200_csr **out);
/* Same with MAGNUS' usize (u64) column indices (src/graph_magnus.rs:13,52-74). */
int b200_csr_upload_idx64(b200_ctx *ctx, uint64_t rows, uint64_t cols, const uint64_t *row_ptr,
                          const uint64_t *col_idx, const void *values, int val_bits, b200_csr **out);
/* Adopt (copy) arrays that already live on this context's device, e.g. an NCCL-broadcast B. */
int b200_csr_from_device(b200_ctx *ctx, uint64_t rows, uint64_t cols, uint64_t nnz,
                         const void *d_row_ptr, const void *d_col_idx, const void *d_values,
                         int val_bits, b200_csr **out);
int b200_csr_free(b200_ctx *ctx, b200_csr *m);

/* Shape / size queries: CsrMatrix::nnz (src/graph_csr.rs:260), MagnusMatrix::nnz (:220). */
int b200_csr_info(const b200_csr *m, uint64_t *rows, uint64_t *cols, uint64_t *nnz, int *val_bits);
/* Raw device pointers (row_ptr u64*, col_idx u32*, values) for zero-copy consumers. */
int b200_csr_device_ptrs(const b200_csr *m, void **d_row_ptr, void **d_col_idx, void **d_values);
/* Largest stored value (device-side reduction kept with every handle). */
int b200_csr_max_value(b200_ctx *ctx, const b200_csr *m, uint64_t *out);

/* Device handle -> caller-allocated host arrays (sizes from b200_csr_info). */
int b200_csr_download(b200_ctx *ctx, const b200_csr *m, uint64_t *row_ptr, uint32_t *col_idx, void *values);
int b200_csr_download_idx64(b200_ctx *ctx, const b200_csr *m, uint64_t *row_ptr, uint64_t *col_idx, void *values);
/* Asynchronous variant into pinned host memory: ordered after the work queued on the ctx stream so far, run on
 * the context's own copy stream so that it overlaps the next multiply; b200_ctx_synchronize waits for it, and
 * freeing the handle is ordered after it.  The host arrays must not be read before b200_ctx_synchronize returns (u64
 * values proven to fit 32 bits travel narrow and are widened into `values` by the library's host threads). */
int b200_csr_download_async(b200_ctx *ctx, const b200_csr *m, uint64_t *row_ptr, uint32_t *col_idx, void *values);

/* C = A x B.  Replaces CsrMatrix::matmul / matmul_par (src/graph_csr.rs:306-346, 350-484),
 * Csr<I,V>::matmul / matmul_par (linalg/src/csr.rs:308-356, 361-466) and
 * MagnusMatrix::matmul / matmul_seq (src/graph_magnus.rs:225-242).  Result: ascending unique
 * columns per row, saturating sums of saturating products, zeros dropped, row_ptr[0] = 0 --
 * bit-identical to the reference's CSR output.  `stats` may be NULL.
 * The call is asynchronous when `stats` is NULL: the product handle is returned while its kernels are still queued and its
 * size (nnz) reaches the host later through a pinned report; any call that needs the size (b200_csr_info with an nnz
 * pointer, downloads, using the handle as an operand) waits for that report by itself.  With `stats` the call returns
 * after the multiply has finished on the device.  b200_csr_product_stats returns the same numbers later. */
int b200_spgemm(b200_ctx *ctx, const b200_csr *A, const b200_csr *B, b200_csr **C, b200_stats *stats);
/* Measurements of the multiply that produced C (waits for that multiply); B200_ERR_BADARG for handles that are not products. */
int b200_csr_product_stats(b200_ctx *ctx, const b200_csr *C, b200_stats *stats);

/* Per-row intermediate-product counts of A x B (host array of A.rows u64); the quantity the
 * multi-GPU row split is balanced on. */
int b200_row_products(b200_ctx *ctx, const b200_csr *A, const b200_csr *B, uint64_t *host_out);
/* Contiguous row cut points cuts[0..nparts] (cuts[0]=0, cuts[nparts]=A.rows) so that every part
 * holds ~1/nparts of the intermediate products of A x B. */
int b200_shard_rows_by_products(b200_ctx *ctx, const b200_csr *A, const b200_csr *B, int nparts, uint64_t *cuts);
/* Rows [row_begin,row_end) of A as a new (row_end-row_begin) x cols handle. */
int b200_csr_row_block(b200_ctx *ctx, const b200_csr *A, uint64_t row_begin, uint64_t row_end, b200_csr **out);

/* C = A + B, element-wise saturating sorted merge: CsrMatrix::add (src/graph_csr.rs:487-542),
 * MagnusMatrix::add (src/graph_magnus.rs:245-302). */
int b200_csr_add(b200_ctx *ctx, const b200_csr *A, const b200_csr *B, b200_csr **C);
/* 1 if A and B have identical row_ptr and col_idx (the stabilisation test of
 * power_until_stable, src/graph_csr.rs:567-570). */
int b200_csr_same_pattern(b200_ctx *ctx, const b200_csr *A, const b200_csr *B, int *same);

/* ---- multi-GPU: the path shards by rows of the left operand -- row i of C needs row i of A and all of B
 * (src/graph_csr.rs:433-446); the reference itself is single-process (rayon), so these have no counterpart there. ----------------
 * One b200_comm per (process, GPU).  NCCL is loaded at run time; without it these return B200_ERR_NCCL.
 *   one process per GPU:  rank 0 calls b200_comm_unique_id, the launcher's own channel (torch.distributed store, MPI, a file)
 *                         carries the 128 bytes to every rank, every rank calls b200_comm_init_rank with its own context;
 *   one process, N GPUs:  b200_comm_init_all over N contexts (out[i] belongs to ctxs[i]); each communicator is then driven by
 *                         its own host thread (tools/b200_bench.cpp).
 * Typical job: rank `root` builds A; b200_comm_broadcast_csr replicates it; every rank calls b200_shard_rows_by_products
 * (same cuts everywhere), keeps b200_csr_row_block(cuts[rank], cuts[rank+1]) and multiplies its block by the replicated
 * operand -- repeated exponentiation A^k = A^(k-1) x A needs nothing else.  b200_comm_allgather_csr assembles the row blocks
 * (rank order = row order) into the whole matrix on every rank: the optional gather of C, and the per-step exchange of a
 * squaring chain (power_until_stable, src/graph_csr.rs:561-575, where the right operand grows too). */
int b200_comm_unique_id(uint8_t id128[128]);
int b200_comm_init_rank(b200_ctx *ctx, int nranks, int rank, const uint8_t id128[128], b200_comm **out);
int b200_comm_init_all(b200_ctx **ctxs, int ngpus, b200_comm **out);
int b200_comm_destroy(b200_comm *c);
int b200_comm_rank(const b200_comm *c, int *rank, int *size);
int b200_comm_broadcast_csr(b200_comm *c, const b200_csr *src /* NULL off the root */, int root, b200_csr **out);
int b200_comm_allgather_csr(b200_comm *c, const b200_csr *block, b200_csr **out);
/* In-place reduction of n <= 32 host scalars over the ranks: op 0 sum of u64, 1 max of u64, 2 sum of f64, 3 max of f64
 * (max-over-ranks step time, total products). */
int b200_comm_allreduce(b200_comm *c, void *host_scalars, int n, int op);

/* ---- fixture generators on the device: large inputs without a host build ---------------------------
 * N-d Moore lattice / torus, CsrMatrix::lattice (src/graph_csr.rs:177-222; MagnusMatrix::lattice, src/graph_magnus.rs:146):
 * node ids row-major with the last dimension fastest, offsets enumerated with dimension 0 as the least-significant
 * digit, torus wrap via rem_euclid, the all-zero offset skipped, duplicate neighbours (side-2 torus) summed.
 * ndims <= 4, at most 2^32-1 nodes. */
int b200_lattice(b200_ctx *ctx, const uint64_t *dims, int ndims, int torus, int val_bits, b200_csr **out);
/* Symmetric Bernoulli thinning, CsrMatrix::thin (src/graph_csr.rs:225-247) with StdRng::from_seed(seed32) (rand 0.9:
 * ChaCha12): entries visited row-major, one draw per stored entry with r <= c, a kept (r,c) also keeps its stored mirror.
 * `skip_draws` = next_u64 outputs already taken from the same generator (bench_matmul_magnus shares one StdRng over its
 * grid, src/graph_magnus.rs:800-821); `draws_consumed` (may be NULL) receives the number this call took. */
int b200_thin(b200_ctx *ctx, const b200_csr *A, double density, const uint8_t *seed32, uint64_t skip_draws, b200_csr **out,
              uint64_t *draws_consumed);
/* COO triplets -> CSR on the device: CsrMatrix::from_coo (src/graph_csr.rs:83-129; MagnusMatrix::from_coo, src/graph_magnus.rs:34-76):
 * triplets ordered by (row, column) with a radix sort, duplicates summed -- plain `+=` (wrapping, as the reference's release
 * build) or, with `saturating`, the saturating sum of linalg's from_coo (linalg/src/csr.rs:158-195) -- zero sums dropped.
 * n triplets in host arrays (`_device`: arrays already on this context's device); B200_ERR_BADARG for an index out of range. */
int b200_csr_from_coo(b200_ctx *ctx, uint64_t rows, uint64_t cols, uint64_t n, const uint32_t *row_idx, const uint32_t *col_idx,
                      const void *values, int val_bits, int saturating, b200_csr **out);
int b200_csr_from_coo_device(b200_ctx *ctx, uint64_t rows, uint64_t cols, uint64_t n, const uint32_t *d_row_idx, const uint32_t *d_col_idx,
                             const void *d_values, int val_bits, int saturating, b200_csr **out);
/* R-MAT graph built on the device (BASELINE configs[3]; SURVEY.md App. C -- the reference has no such generator): 2^scale nodes,
 * edge_factor * 2^scale edges; level l of edge e takes draw e * scale + l of a counter-based splitmix64(seed): u < a -> quadrant
 * (0,0), < a+b -> (0,1), < a+b+c -> (1,0), else (1,1); duplicate edges are summed by from_coo (plain +=). */
int b200_rmat(b200_ctx *ctx, int scale, uint64_t edge_factor, double a, double b, double c, uint64_t seed, int val_bits, b200_csr **out);
/* ---- locality pre-pass: CsrMatrix::rcm / permute / unpermute / bandwidth_stats, src/graph_csr.rs:663-818 -------
 * (max |r - c|, mean |r - c|) over the stored entries: bandwidth_stats (:802-818); a device reduction. */
int b200_csr_bandwidth_stats(b200_ctx *ctx, const b200_csr *m, uint64_t *max_bw, double *avg_bw);
/* Rows and columns reordered by perm[new] = old (host array of n = rows entries): permute (:727-785); the entries are relabelled
 * and re-assembled on the device (radix sort), every row sorted by its new columns.  unpermute (:787-799) is a permute by the
 * inverse.  B200_ERR_BADARG unless perm is a permutation of 0..n-1, B200_ERR_SHAPE unless the matrix is square. */
int b200_csr_permute(b200_ctx *ctx, const b200_csr *A, const uint32_t *perm, b200_csr **out);
/* Reverse Cuthill-McKee order of the pattern, perm_out[new] = old (host array of n entries): the ordering loop of rcm (:663-723).
 * The queue is sequential by definition, so it runs on the host inside the library; apply it with b200_csr_permute. */
int b200_csr_rcm_order(b200_ctx *ctx, const b200_csr *A, uint32_t *perm_out);
/* Host twin of the generator the device kernels use (runs without a GPU): StdRng::from_seed(seed32).next_u64() outputs
 * number first .. first+n-1. */
int b200_stdrng_u64(const uint8_t *seed32, uint64_t first, uint64_t n, uint64_t *out);

#ifdef __cplusplus
}
#endif
#endif /* B200_SPGEMM_H */
