/*
 * oracle/spgemm_oracle.c -- TEST INFRASTRUCTURE ONLY.  NOT PART OF THE PRODUCT.
 *
 * A plain-C CPU restatement of the reference's CSR x CSR SpGEMM path
 * (imlvts/sparse-linear-algebra-tests).  It is the checker the CUDA engine is
 * compared against; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product library
 * (sparse_linear_algebra_tests_b200/csrc) never links or calls anything here.
 *
 * The reference is Rust (nightly) with two git dependencies and cannot be
 * built in this image (no cargo/rustc, no network), so there is no
 * oracle/_ref.  Parity of this restatement is pinned by
 *   (1) every known-answer test the reference's own test modules hold for the
 *       path (tests/test_oracle_kat.py cites each one), and
 *   (2) the reference-faithful benchmark instance (ChaCha12 StdRng seed
 *       [42;32], rand 0.9 f64 sampling) whose per-power nnz must round to the
 *       values published in the reference README.md:42-47.
 *
 * Each function cites the reference file:line it follows (paths relative to
 * /root/reference).  The reference multiplies square n x n matrices only
 * (src/graph_csr.rs:307,351); the restatement accepts a rows_A x n left
 * operand so that row-sharded blocks can be checked too -- the per-row
 * arithmetic is unchanged.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    uint64_t rows, cols, nnz;
    uint64_t *row_ptr;   /* rows+1, usize in the reference (graph_csr.rs:46) */
    uint32_t *col_idx;   /* NodeId = u32 (graph_csr.rs:14,48)                */
    void     *values;    /* u32 (graph_csr.rs:17) or u64 (graph_sprs.rs:16)  */
    int       val_bits;  /* 32 or 64                                         */
} ocsr_t;

void oracle_free(ocsr_t *m) {
    if (!m) return;
    free(m->row_ptr); free(m->col_idx); free(m->values); free(m);
}

static ocsr_t *ocsr_alloc(uint64_t rows, uint64_t cols, uint64_t nnz, int val_bits) {
    ocsr_t *m = (ocsr_t *)calloc(1, sizeof(ocsr_t));
    m->rows = rows; m->cols = cols; m->nnz = nnz; m->val_bits = val_bits;
    m->row_ptr = (uint64_t *)calloc(rows + 1, sizeof(uint64_t));
    m->col_idx = (uint32_t *)malloc((nnz ? nnz : 1) * sizeof(uint32_t));
    m->values  = malloc((nnz ? nnz : 1) * (size_t)(val_bits / 8));
    return m;
}

/* Wrap caller arrays (no copy) so ctypes can hand numpy buffers in. */
ocsr_t *oracle_wrap(uint64_t rows, uint64_t cols, uint64_t nnz, uint64_t *row_ptr,
                    uint32_t *col_idx, void *values, int val_bits) {
    ocsr_t *m = (ocsr_t *)calloc(1, sizeof(ocsr_t));
    m->rows = rows; m->cols = cols; m->nnz = nnz; m->val_bits = val_bits;
    m->row_ptr = row_ptr; m->col_idx = col_idx; m->values = values;
    return m;
}
void oracle_unwrap(ocsr_t *m) { free(m); }

/* ---------------------------------------------------------------- sorting */
/* Stand-in for Rust's slice::sort_unstable (graph_csr.rs:331,449): an inlined
 * introsort on u32 so the timed CPU baseline is not handicapped by qsort's
 * indirect comparator calls. */
static inline void isort_u32(uint32_t *a, size_t n) {
    for (size_t i = 1; i < n; i++) {
        uint32_t x = a[i]; size_t j = i;
        while (j > 0 && a[j - 1] > x) { a[j] = a[j - 1]; j--; }
        a[j] = x;
    }
}
static void heap_sift(uint32_t *a, size_t i, size_t n) {
    for (;;) {
        size_t l = 2 * i + 1, r = l + 1, m = i;
        if (l < n && a[l] > a[m]) m = l;
        if (r < n && a[r] > a[m]) m = r;
        if (m == i) return;
        uint32_t t = a[i]; a[i] = a[m]; a[m] = t; i = m;
    }
}
static void hsort_u32(uint32_t *a, size_t n) {
    for (size_t i = n / 2; i-- > 0;) heap_sift(a, i, n);
    for (size_t e = n; e-- > 1;) { uint32_t t = a[0]; a[0] = a[e]; a[e] = t; heap_sift(a, 0, e); }
}
static void qsort_u32_rec(uint32_t *a, size_t n, int depth) {
    while (n > 24) {
        if (depth-- == 0) { hsort_u32(a, n); return; }
        uint32_t x = a[0], y = a[n / 2], z = a[n - 1];
        uint32_t p = (x < y) ? ((y < z) ? y : (x < z ? z : x)) : ((x < z) ? x : (y < z ? z : y));
        size_t i = 0, j = n - 1;
        for (;;) {
            while (a[i] < p) i++;
            while (a[j] > p) j--;
            if (i >= j) break;
            uint32_t t = a[i]; a[i] = a[j]; a[j] = t; i++; j--;
        }
        size_t left = j + 1;
        if (left < n - left) { qsort_u32_rec(a, left, depth); a += left; n -= left; }
        else { qsort_u32_rec(a + left, n - left, depth); n = left; }
    }
    isort_u32(a, n);
}
static inline void sort_u32(uint32_t *a, size_t n) {
    int depth = 2; for (size_t t = n; t > 1; t >>= 1) depth += 2;
    qsort_u32_rec(a, n, depth);
}

/* ------------------------------------------------ saturating arithmetic */
/* graph_csr.rs:30-37 (u32), graph_sprs.rs:29-51 Sat64 (u64),
 * linalg/src/csr.rs:63-73 (generic). */
static inline uint32_t sadd_u32(uint32_t a, uint32_t b) { uint32_t s = a + b; return s < a ? UINT32_MAX : s; }
static inline uint32_t smul_u32(uint32_t a, uint32_t b) { uint64_t p = (uint64_t)a * b; return p > UINT32_MAX ? UINT32_MAX : (uint32_t)p; }
static inline uint64_t sadd_u64(uint64_t a, uint64_t b) { uint64_t s = a + b; return s < a ? UINT64_MAX : s; }
static inline uint64_t smul_u64(uint64_t a, uint64_t b) { unsigned __int128 p = (unsigned __int128)a * b; return p > UINT64_MAX ? UINT64_MAX : (uint64_t)p; }

uint32_t oracle_sadd_u32(uint32_t a, uint32_t b) { return sadd_u32(a, b); }
uint32_t oracle_smul_u32(uint32_t a, uint32_t b) { return smul_u32(a, b); }
uint64_t oracle_sadd_u64(uint64_t a, uint64_t b) { return sadd_u64(a, b); }
uint64_t oracle_smul_u64(uint64_t a, uint64_t b) { return smul_u64(a, b); }

/* -------------------------------------------------- growable output vecs */
typedef struct { uint32_t *c; void *v; size_t len, cap; int vb; } ovec_t;
static void ovec_reserve(ovec_t *o, size_t need) {
    if (need <= o->cap) return;
    size_t nc = o->cap ? o->cap : 1024;
    while (nc < need) nc *= 2;
    o->c = (uint32_t *)realloc(o->c, nc * sizeof(uint32_t));
    o->v = realloc(o->v, nc * (size_t)(o->vb / 8));
    o->cap = nc;
}

/* ------------------------------------------------------- typed kernels */
#define DEFINE_TYPED(VT, SUF, BITS)                                                              \
/* CsrMatrix::matmul -- graph_csr.rs:306-346 (u64 twin: linalg/src/csr.rs:308-356).            \
 * Gustavson row-wise product, dense accumulator + touched-column list, sort, emit non-zeros. */ \
ocsr_t *oracle_matmul_seq_##SUF(const ocsr_t *A, const ocsr_t *B) {                              \
    if (A->cols != B->rows || A->val_bits != BITS || B->val_bits != BITS) return NULL;          \
    const VT *av = (const VT *)A->values, *bv = (const VT *)B->values;                           \
    uint64_t m = A->rows, n = B->cols;                                                           \
    uint64_t *row_ptr = (uint64_t *)calloc(m + 1, sizeof(uint64_t));                             \
    ovec_t out = {0, 0, 0, 0, BITS};                                                             \
    VT *acc = (VT *)calloc(n ? n : 1, sizeof(VT));                                               \
    uint32_t *nz = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));                           \
    for (uint64_t i = 0; i < m; i++) {                                                           \
        size_t nnz_cols = 0;                                                                     \
        for (uint64_t ia = A->row_ptr[i]; ia < A->row_ptr[i + 1]; ia++) {                        \
            uint32_t k = A->col_idx[ia]; VT a = av[ia];                                          \
            for (uint64_t jb = B->row_ptr[k]; jb < B->row_ptr[k + 1]; jb++) {                    \
                uint32_t j = B->col_idx[jb];                                                     \
                if (acc[j] == 0) nz[nnz_cols++] = j;      /* :323 first-touch on acc==0 */      \
                acc[j] = sadd_##SUF(acc[j], smul_##SUF(a, bv[jb]));   /* :326 */                 \
            }                                                                                    \
        }                                                                                        \
        sort_u32(nz, nnz_cols);                               /* :331 */                         \
        ovec_reserve(&out, out.len + nnz_cols);                                                  \
        for (size_t t = 0; t < nnz_cols; t++) {                                                  \
            uint32_t j = nz[t]; VT v = acc[j];                                                   \
            if (v != 0) { out.c[out.len] = j; ((VT *)out.v)[out.len] = v; out.len++; } /* :334 */\
            acc[j] = 0;                                                                          \
        }                                                                                        \
        row_ptr[i + 1] = out.len;                             /* :342 */                         \
    }                                                                                            \
    free(acc); free(nz);                                                                         \
    ocsr_t *C = (ocsr_t *)calloc(1, sizeof(ocsr_t));                                             \
    C->rows = m; C->cols = n; C->nnz = out.len; C->val_bits = BITS; C->row_ptr = row_ptr;        \
    ovec_reserve(&out, 1); C->col_idx = out.c; C->values = out.v;                                \
    return C;                                                                                    \
}                                                                                                \
                                                                                                 \
/* CsrMatrix::matmul_par -- graph_csr.rs:350-484 (u64: linalg/src/csr.rs:361-466).             \
 * Pass 1 symbolic (bool mask, count, re-walk to clear) :362-403; serial prefix sum :412-417;   \
 * exact zero-filled allocation :420-421; pass 2 numeric (dense acc, sort, disjoint writes)     \
 * :430-476.  rayon's work-stealing row split becomes omp schedule(dynamic); per-thread         \
 * scratch is allocated once per thread (rayon's for_each_init allocates at least that often).*/\
ocsr_t *oracle_matmul_par_##SUF(const ocsr_t *A, const ocsr_t *B, int nthreads) {                \
    if (A->cols != B->rows || A->val_bits != BITS || B->val_bits != BITS) return NULL;          \
    const VT *av = (const VT *)A->values, *bv = (const VT *)B->values;                           \
    int64_t m = (int64_t)A->rows; uint64_t n = B->cols;                                          \
    if (nthreads <= 0) nthreads = 1;                                                             \
    uint64_t *nnz_per_row = (uint64_t *)calloc(m ? m : 1, sizeof(uint64_t));                     \
    _Pragma("omp parallel num_threads(nthreads)")                                                \
    {                                                                                            \
        uint8_t *mask = (uint8_t *)calloc(n ? n : 1, 1);                                         \
        _Pragma("omp for schedule(dynamic, 64)")                                                 \
        for (int64_t i = 0; i < m; i++) {                                                        \
            uint64_t count = 0;                                                                  \
            for (uint64_t ia = A->row_ptr[i]; ia < A->row_ptr[i + 1]; ia++) {                    \
                uint32_t k = A->col_idx[ia];                                                     \
                for (uint64_t jb = B->row_ptr[k]; jb < B->row_ptr[k + 1]; jb++) {                \
                    uint32_t j = B->col_idx[jb];                                                 \
                    if (!mask[j]) { mask[j] = 1; count++; }                                      \
                }                                                                                \
            }                                                                                    \
            nnz_per_row[i] = count;                                                              \
            for (uint64_t ia = A->row_ptr[i]; ia < A->row_ptr[i + 1]; ia++) {                    \
                uint32_t k = A->col_idx[ia];                                                     \
                for (uint64_t jb = B->row_ptr[k]; jb < B->row_ptr[k + 1]; jb++)                  \
                    mask[B->col_idx[jb]] = 0;                                                    \
            }                                                                                    \
        }                                                                                        \
        free(mask);                                                                              \
    }                                                                                            \
    uint64_t *row_ptr = (uint64_t *)calloc(m + 1, sizeof(uint64_t));                             \
    for (int64_t i = 0; i < m; i++) row_ptr[i + 1] = row_ptr[i] + nnz_per_row[i];                \
    uint64_t total = row_ptr[m];                                                                 \
    free(nnz_per_row);                                                                           \
    uint32_t *col_idx = (uint32_t *)calloc(total ? total : 1, sizeof(uint32_t));                 \
    VT *values = (VT *)calloc(total ? total : 1, sizeof(VT));                                    \
    _Pragma("omp parallel num_threads(nthreads)")                                                \
    {                                                                                            \
        VT *acc = (VT *)calloc(n ? n : 1, sizeof(VT));                                           \
        size_t nzcap = 1024; uint32_t *nz = (uint32_t *)malloc(nzcap * sizeof(uint32_t));        \
        _Pragma("omp for schedule(dynamic, 64)")                                                 \
        for (int64_t i = 0; i < m; i++) {                                                        \
            size_t cnt = 0;                                                                      \
            for (uint64_t ia = A->row_ptr[i]; ia < A->row_ptr[i + 1]; ia++) {                    \
                uint32_t k = A->col_idx[ia]; VT a = av[ia];                                      \
                for (uint64_t jb = B->row_ptr[k]; jb < B->row_ptr[k + 1]; jb++) {                \
                    uint32_t j = B->col_idx[jb];                                                 \
                    if (acc[j] == 0) {                                                           \
                        if (cnt == nzcap) { nzcap *= 2; nz = (uint32_t *)realloc(nz, nzcap * 4); }\
                        nz[cnt++] = j;                                                           \
                    }                                                                            \
                    acc[j] = sadd_##SUF(acc[j], smul_##SUF(a, bv[jb]));                          \
                }                                                                                \
            }                                                                                    \
            sort_u32(nz, cnt);                                                                   \
            uint64_t pos = row_ptr[i];                                                           \
            for (size_t t = 0; t < cnt; t++) {                                                   \
                uint32_t j = nz[t]; VT v = acc[j];                                               \
                if (v != 0) { col_idx[pos] = j; values[pos] = v; pos++; }                        \
                acc[j] = 0;                                                                      \
            }                                                                                    \
        }                                                                                        \
        free(acc); free(nz);                                                                     \
    }                                                                                            \
    ocsr_t *C = (ocsr_t *)calloc(1, sizeof(ocsr_t));                                             \
    C->rows = (uint64_t)m; C->cols = n; C->nnz = total; C->val_bits = BITS;                      \
    C->row_ptr = row_ptr; C->col_idx = col_idx; C->values = values;                              \
    return C;                                                                                    \
}                                                                                                \
                                                                                                 \
/* CsrMatrix::add -- graph_csr.rs:487-542: per-row sorted merge, saturating on equal cols. */    \
ocsr_t *oracle_add_##SUF(const ocsr_t *A, const ocsr_t *B) {                                     \
    if (A->rows != B->rows || A->cols != B->cols) return NULL;                                   \
    const VT *av = (const VT *)A->values, *bv = (const VT *)B->values;                           \
    ocsr_t *C = ocsr_alloc(A->rows, A->cols, A->nnz + B->nnz, BITS);                             \
    VT *cv = (VT *)C->values; uint64_t len = 0;                                                  \
    for (uint64_t r = 0; r < A->rows; r++) {                                                     \
        uint64_t ai = A->row_ptr[r], ae = A->row_ptr[r + 1], bi = B->row_ptr[r], be = B->row_ptr[r + 1]; \
        while (ai < ae && bi < be) {                                                             \
            uint32_t ac = A->col_idx[ai], bc = B->col_idx[bi];                                   \
            if (ac < bc) { C->col_idx[len] = ac; cv[len++] = av[ai++]; }                         \
            else if (ac > bc) { C->col_idx[len] = bc; cv[len++] = bv[bi++]; }                    \
            else { VT v = sadd_##SUF(av[ai], bv[bi]); if (v != 0) { C->col_idx[len] = ac; cv[len++] = v; } ai++; bi++; } \
        }                                                                                        \
        while (ai < ae) { C->col_idx[len] = A->col_idx[ai]; cv[len++] = av[ai++]; }              \
        while (bi < be) { C->col_idx[len] = B->col_idx[bi]; cv[len++] = bv[bi++]; }              \
        C->row_ptr[r + 1] = len;                                                                 \
    }                                                                                            \
    C->nnz = len;                                                                                \
    return C;                                                                                    \
}                                                                                                \
                                                                                                 \
/* CsrMatrix::from_coo -- graph_csr.rs:83-129: sort by (r,c), sum duplicates (plain `+=`,      \
 * wrapping in release: :93; saturating in linalg/src/csr.rs:167 -> `saturating` flag),         \
 * drop zeros (:108), close trailing rows (:117-120). */                                         \
ocsr_t *oracle_from_coo_##SUF(uint64_t rows, uint64_t cols, uint64_t nt, const uint32_t *tr,     \
                              const uint32_t *tc, const VT *tv, int saturating) {                \
    coo_ent_t *e = (coo_ent_t *)malloc((nt ? nt : 1) * sizeof(coo_ent_t));                       \
    for (uint64_t i = 0; i < nt; i++) { e[i].key = ((uint64_t)tr[i] << 32) | tc[i]; e[i].val = tv[i]; } \
    qsort(e, nt, sizeof(coo_ent_t), coo_cmp);                                                    \
    uint64_t nd = 0;                                                                             \
    for (uint64_t i = 0; i < nt; i++) {                                                          \
        if (nd > 0 && e[nd - 1].key == e[i].key) {                                               \
            VT s = (VT)e[nd - 1].val, v = (VT)e[i].val;                                          \
            e[nd - 1].val = saturating ? sadd_##SUF(s, v) : (VT)(s + v);                         \
        } else e[nd++] = e[i];                                                                   \
    }                                                                                            \
    ocsr_t *C = ocsr_alloc(rows, cols, nd, BITS);                                                \
    VT *cv = (VT *)C->values; uint64_t len = 0, cur = 0;                                         \
    for (uint64_t i = 0; i < nd; i++) {                                                          \
        if ((VT)e[i].val == 0) continue;                                                         \
        uint64_t r = e[i].key >> 32;                                                             \
        while (cur <= r) C->row_ptr[cur++] = len;                                                \
        C->col_idx[len] = (uint32_t)e[i].key; cv[len++] = (VT)e[i].val;                          \
    }                                                                                            \
    while (cur <= rows) C->row_ptr[cur++] = len;                                                 \
    C->nnz = len; free(e);                                                                       \
    return C;                                                                                    \
}                                                                                                \
                                                                                                 \
/* CsrMatrix::get -- graph_csr.rs:250-257 (binary search in the sorted row). */                  \
VT oracle_get_##SUF(const ocsr_t *A, uint64_t r, uint32_t c) {                                   \
    uint64_t lo = A->row_ptr[r], hi = A->row_ptr[r + 1];                                         \
    while (lo < hi) { uint64_t mid = (lo + hi) / 2;                                              \
        if (A->col_idx[mid] < c) lo = mid + 1; else hi = mid; }                                  \
    if (lo < A->row_ptr[r + 1] && A->col_idx[lo] == c) return ((const VT *)A->values)[lo];       \
    return 0;                                                                                    \
}

typedef struct { uint64_t key, val; } coo_ent_t;
static int coo_cmp(const void *a, const void *b) {
    uint64_t x = ((const coo_ent_t *)a)->key, y = ((const coo_ent_t *)b)->key;
    return x < y ? -1 : (x > y ? 1 : 0);
}

DEFINE_TYPED(uint32_t, u32, 32)
DEFINE_TYPED(uint64_t, u64, 64)

/* value-width conversion (the bench rebuilds the MAGNUS u64 matrix from the u32
 * CSR triplets: graph_magnus.rs:720-729) */
ocsr_t *oracle_widen_to_u64(const ocsr_t *A) {
    ocsr_t *C = ocsr_alloc(A->rows, A->cols, A->nnz, 64);
    memcpy(C->row_ptr, A->row_ptr, (A->rows + 1) * sizeof(uint64_t));
    memcpy(C->col_idx, A->col_idx, A->nnz * sizeof(uint32_t));
    for (uint64_t i = 0; i < A->nnz; i++)
        ((uint64_t *)C->values)[i] = (A->val_bits == 32) ? ((uint32_t *)A->values)[i] : ((uint64_t *)A->values)[i];
    return C;
}
ocsr_t *oracle_narrow_to_u32(const ocsr_t *A) {
    ocsr_t *C = ocsr_alloc(A->rows, A->cols, A->nnz, 32);
    memcpy(C->row_ptr, A->row_ptr, (A->rows + 1) * sizeof(uint64_t));
    memcpy(C->col_idx, A->col_idx, A->nnz * sizeof(uint32_t));
    for (uint64_t i = 0; i < A->nnz; i++)
        ((uint32_t *)C->values)[i] = (A->val_bits == 32) ? ((uint32_t *)A->values)[i] : (uint32_t)((uint64_t *)A->values)[i];
    return C;
}

/* per-row intermediate-product counts: sum over (i,k) in A of nnz(B[k,:]) */
void oracle_row_products(const ocsr_t *A, const ocsr_t *B, uint64_t *out) {
    for (uint64_t i = 0; i < A->rows; i++) {
        uint64_t p = 0;
        for (uint64_t ia = A->row_ptr[i]; ia < A->row_ptr[i + 1]; ia++) {
            uint32_t k = A->col_idx[ia];
            p += B->row_ptr[k + 1] - B->row_ptr[k];
        }
        out[i] = p;
    }
}

/* ------------------------------------------------------------ generators */
/* CsrMatrix::lattice -- graph_csr.rs:177-222 (doc: graph.rs:80-139).  Row-major node ids with
 * the last dim fastest (:181-184); offsets enumerated base-3 with dimension 0 as the least
 * significant digit (:192-198); torus wraps with rem_euclid (:201-202); self offset skipped
 * (:211); duplicates (side-2 torus dims) are summed by from_coo. */
ocsr_t *oracle_lattice(const uint64_t *dims, int ndim, int torus, int val_bits) {
    uint64_t total = 1; for (int d = 0; d < ndim; d++) total *= dims[d];
    uint64_t strides[16]; for (int d = 0; d < ndim; d++) strides[d] = 1;
    for (int i = ndim - 2; i >= 0; i--) strides[i] = strides[i + 1] * dims[i + 1];
    uint64_t nnb = 1; for (int d = 0; d < ndim; d++) nnb *= 3;
    size_t cap = (size_t)(total * nnb + 1);
    uint32_t *tr = (uint32_t *)malloc(cap * 4), *tc = (uint32_t *)malloc(cap * 4);
    uint64_t nt = 0; uint64_t coord[16] = {0};
    for (uint64_t node = 0; node < total; node++) {
        for (uint64_t off = 0; off < nnb; off++) {
            uint64_t tmp = off; int all_zero = 1, valid = 1; uint64_t nb = 0;
            for (int d = 0; d < ndim; d++) {
                int64_t delta = (int64_t)(tmp % 3) - 1; tmp /= 3;
                if (delta != 0) all_zero = 0;
                int64_t c = (int64_t)coord[d] + delta;
                if (torus) { int64_t md = (int64_t)dims[d]; c = ((c % md) + md) % md; }
                else if (c < 0 || c >= (int64_t)dims[d]) { valid = 0; break; }
                nb += (uint64_t)c * strides[d];
            }
            if (all_zero || !valid) continue;
            tr[nt] = (uint32_t)node; tc[nt] = (uint32_t)nb; nt++;
        }
        for (int d = ndim - 1; d >= 0; d--) { coord[d]++; if (coord[d] < dims[d]) break; coord[d] = 0; }
    }
    ocsr_t *C;
    if (val_bits == 32) {
        uint32_t *tv = (uint32_t *)malloc((nt ? nt : 1) * 4); for (uint64_t i = 0; i < nt; i++) tv[i] = 1;
        C = oracle_from_coo_u32(total, total, nt, tr, tc, tv, 0); free(tv);
    } else {
        uint64_t *tv = (uint64_t *)malloc((nt ? nt : 1) * 8); for (uint64_t i = 0; i < nt; i++) tv[i] = 1;
        C = oracle_from_coo_u64(total, total, nt, tr, tc, tv, 0); free(tv);
    }
    free(tr); free(tc);
    return C;
}

/* ChaCha12 block RNG == rand 0.9.2 StdRng (rand_chacha 0.9.0, Cargo.lock:851-878):
 * key = seed bytes as 8 LE words, 64-bit block counter in words 12-13 starting at 0,
 * stream id 0 in words 14-15, 12 rounds, 4 blocks (64 words) buffered per refill;
 * next_u64 = lo word | hi word << 32 of two consecutive buffer words. */
typedef struct { uint32_t key[8]; uint64_t counter; uint32_t buf[64]; int idx; } chacha12_t;
#define ROTL32(x, n) (((x) << (n)) | ((x) >> (32 - (n))))
#define QR(a, b, c, d) a += b; d ^= a; d = ROTL32(d, 16); c += d; b ^= c; b = ROTL32(b, 12); \
                       a += b; d ^= a; d = ROTL32(d, 8);  c += d; b ^= c; b = ROTL32(b, 7);
static void chacha12_block(const uint32_t key[8], uint64_t counter, uint32_t out[16]) {
    uint32_t s[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u,
                      key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                      (uint32_t)counter, (uint32_t)(counter >> 32), 0, 0};
    uint32_t x[16]; memcpy(x, s, sizeof(x));
    for (int r = 0; r < 6; r++) {
        QR(x[0], x[4], x[8], x[12]) QR(x[1], x[5], x[9], x[13]) QR(x[2], x[6], x[10], x[14]) QR(x[3], x[7], x[11], x[15])
        QR(x[0], x[5], x[10], x[15]) QR(x[1], x[6], x[11], x[12]) QR(x[2], x[7], x[8], x[13]) QR(x[3], x[4], x[9], x[14])
    }
    for (int i = 0; i < 16; i++) out[i] = x[i] + s[i];
}
static void chacha12_init(chacha12_t *g, const uint8_t seed[32]) {
    for (int i = 0; i < 8; i++)
        g->key[i] = (uint32_t)seed[4 * i] | ((uint32_t)seed[4 * i + 1] << 8) | ((uint32_t)seed[4 * i + 2] << 16) | ((uint32_t)seed[4 * i + 3] << 24);
    g->counter = 0; g->idx = 64;
}
static uint64_t chacha12_next_u64(chacha12_t *g) {
    if (g->idx >= 64) {
        for (int b = 0; b < 4; b++) chacha12_block(g->key, g->counter + (uint64_t)b, g->buf + 16 * b);
        g->counter += 4; g->idx = 0;
    }
    uint64_t v = (uint64_t)g->buf[g->idx] | ((uint64_t)g->buf[g->idx + 1] << 32);
    g->idx += 2;
    return v;
}
/* rand 0.9 `rng.random_range(0.0..1.0)` for f64 (UniformFloat::sample_single): take the
 * top 52 bits of one u64 as the mantissa of a float in [1,2), subtract 1, scale by
 * (high-low)=1 and add low=0. */
static double stdrng_f64_01(chacha12_t *g) {
    return (double)(chacha12_next_u64(g) >> 12) * (1.0 / 4503599627370496.0);
}
void oracle_chacha12_u64(const uint8_t seed[32], uint64_t n, uint64_t *out) {
    chacha12_t g; chacha12_init(&g, seed);
    for (uint64_t i = 0; i < n; i++) out[i] = chacha12_next_u64(&g);
}

/* CsrMatrix::thin -- graph_csr.rs:225-247: rows ascending, entries in column order; a draw is
 * consumed only when r <= c (`&&` short-circuit, :235); keep (r,c,v) and mirror (c,r,get(c,r))
 * when present (:236-241); from_coo.  RNG = StdRng::from_seed(seed) (graph_magnus.rs:707). */
ocsr_t *oracle_thin_stdrng(const ocsr_t *A, const uint8_t seed[32], double density) {
    chacha12_t g; chacha12_init(&g, seed);
    size_t cap = A->nnz + 1; uint64_t nt = 0;
    uint32_t *tr = (uint32_t *)malloc(cap * 4), *tc = (uint32_t *)malloc(cap * 4);
    uint64_t *tv = (uint64_t *)malloc(cap * 8);
    for (uint64_t r = 0; r < A->rows; r++) {
        for (uint64_t idx = A->row_ptr[r]; idx < A->row_ptr[r + 1]; idx++) {
            uint32_t c = A->col_idx[idx];
            uint64_t v = A->val_bits == 32 ? ((uint32_t *)A->values)[idx] : ((uint64_t *)A->values)[idx];
            if (r <= c && stdrng_f64_01(&g) < density) {
                tr[nt] = (uint32_t)r; tc[nt] = c; tv[nt] = v; nt++;
                if (r != c) {
                    uint64_t rev = A->val_bits == 32 ? oracle_get_u32(A, c, (uint32_t)r) : oracle_get_u64(A, c, (uint32_t)r);
                    if (rev > 0) { tr[nt] = c; tc[nt] = (uint32_t)r; tv[nt] = rev; nt++; }
                }
            }
        }
    }
    ocsr_t *C;
    if (A->val_bits == 32) {
        uint32_t *tv32 = (uint32_t *)malloc((nt ? nt : 1) * 4); for (uint64_t i = 0; i < nt; i++) tv32[i] = (uint32_t)tv[i];
        C = oracle_from_coo_u32(A->rows, A->cols, nt, tr, tc, tv32, 0); free(tv32);
    } else C = oracle_from_coo_u64(A->rows, A->cols, nt, tr, tc, tv, 0);
    free(tr); free(tc); free(tv);
    return C;
}

/* linalg/benches/perf.rs:43-52 xorshift64 (13/7/17) and :62-95 lattice_csr: the portable
 * 3-D Moore torus, directed Bernoulli thinning with keep_p = clamp(epn/26,0,1), then
 * from_coo (linalg flavour: saturating duplicate sum, linalg/src/csr.rs:158-195). */
ocsr_t *oracle_lattice_csr_xorshift(uint64_t s, double target_epn, uint64_t seed, int val_bits) {
    uint64_t nu = s * s * s; size_t cap = (size_t)(nu * 26 + 1);
    uint32_t *tr = (uint32_t *)malloc(cap * 4), *tc = (uint32_t *)malloc(cap * 4);
    double keep_p = target_epn / 26.0; if (keep_p < 0) keep_p = 0; if (keep_p > 1) keep_p = 1;
    uint64_t x = seed ? seed : 1, nt = 0; int64_t S = (int64_t)s;
    for (uint64_t node = 0; node < nu; node++) {
        int64_t cx = (int64_t)(node / (s * s)), cy = (int64_t)((node / s) % s), cz = (int64_t)(node % s);
        for (int dx = -1; dx <= 1; dx++) for (int dy = -1; dy <= 1; dy++) for (int dz = -1; dz <= 1; dz++) {
            if (!dx && !dy && !dz) continue;
            uint64_t nx = (uint64_t)(((cx + dx) % S + S) % S), ny = (uint64_t)(((cy + dy) % S + S) % S), nz = (uint64_t)(((cz + dz) % S + S) % S);
            x ^= x << 13; x ^= x >> 7; x ^= x << 17;
            double p = (double)x / (double)UINT64_MAX;
            if (p < keep_p) { tr[nt] = (uint32_t)node; tc[nt] = (uint32_t)(nx * s * s + ny * s + nz); nt++; }
        }
    }
    ocsr_t *C;
    if (val_bits == 32) {
        uint32_t *tv = (uint32_t *)malloc((nt ? nt : 1) * 4); for (uint64_t i = 0; i < nt; i++) tv[i] = 1;
        C = oracle_from_coo_u32(nu, nu, nt, tr, tc, tv, 1); free(tv);
    } else {
        uint64_t *tv = (uint64_t *)malloc((nt ? nt : 1) * 8); for (uint64_t i = 0; i < nt; i++) tv[i] = 1;
        C = oracle_from_coo_u64(nu, nu, nt, tr, tc, tv, 1); free(tv);
    }
    free(tr); free(tc);
    return C;
}

/* R-MAT (not in the reference): scale levels, edge_factor * 2^scale
 * edges, per level one splitmix64 draw u in [0,1): u<a -> (0,0), <a+b -> (0,1), <a+b+c -> (1,0),
 * else (1,1); bit `level` of row/col set accordingly; duplicates summed by from_coo. */
static inline uint64_t splitmix64_at(uint64_t seed, uint64_t i) {
    uint64_t z = seed + (i + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
ocsr_t *oracle_rmat(int scale, uint64_t edge_factor, double a, double b, double c, uint64_t seed, int val_bits) {
    uint64_t n = 1ull << scale, m = edge_factor << scale;
    uint32_t *tr = (uint32_t *)malloc((m ? m : 1) * 4), *tc = (uint32_t *)malloc((m ? m : 1) * 4);
    for (uint64_t e = 0; e < m; e++) {
        uint32_t r = 0, cc = 0;
        for (int l = 0; l < scale; l++) {
            double u = (double)(splitmix64_at(seed, e * (uint64_t)scale + (uint64_t)l) >> 11) * (1.0 / 9007199254740992.0);
            if (u < a) {} else if (u < a + b) cc |= 1u << l; else if (u < a + b + c) r |= 1u << l; else { r |= 1u << l; cc |= 1u << l; }
        }
        tr[e] = r; tc[e] = cc;
    }
    ocsr_t *C;
    if (val_bits == 32) {
        uint32_t *tv = (uint32_t *)malloc((m ? m : 1) * 4); for (uint64_t i = 0; i < m; i++) tv[i] = 1;
        C = oracle_from_coo_u32(n, n, m, tr, tc, tv, 0); free(tv);
    } else {
        uint64_t *tv = (uint64_t *)malloc((m ? m : 1) * 8); for (uint64_t i = 0; i < m; i++) tv[i] = 1;
        C = oracle_from_coo_u64(n, n, m, tr, tc, tv, 0); free(tv);
    }
    free(tr); free(tc);
    return C;
}

/* ------------------------------------------------------- timing helpers */
/* ---------------------------------------------------------------------------------------------
 * Locality pre-pass: CsrMatrix::rcm / permute / bandwidth_stats, graph_csr.rs:663-818.
 * --------------------------------------------------------------------------------------------- */
/* CsrMatrix::rcm's ordering -- graph_csr.rs:663-723: for every unvisited seed in index order, a plain BFS (own
 * visited set) whose LAST dequeued node is the start (:674-694); then a BFS from the start where each node's unvisited
 * neighbours are appended by ascending degree (:697-717); the whole order reversed (:721).  perm[new] = old.
 * `sort_unstable_by_key` (:711) leaves the order of equal degrees to the implementation; Rust's is an insertion sort
 * for slices of <= 20 elements, i.e. equal degrees keep their adjacency order -- restated here for every length. */
void oracle_rcm_order(const ocsr_t *A, uint32_t *perm) {
    const uint64_t n = A->rows;
    uint8_t *visited = (uint8_t *)calloc(n ? n : 1, 1), *vis2 = (uint8_t *)malloc(n ? n : 1);
    uint32_t *queue = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t)), *order = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
    uint32_t *nbrs = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
    uint64_t no = 0;
    for (uint64_t seed = 0; seed < n; seed++) {
        if (visited[seed]) continue;
        memset(vis2, 0, n);
        uint64_t qh = 0, qt = 0; uint32_t last = (uint32_t)seed;
        queue[qt++] = (uint32_t)seed; vis2[seed] = 1;
        while (qh < qt) {
            const uint32_t u = queue[qh++]; last = u;
            for (uint64_t i = A->row_ptr[u]; i < A->row_ptr[u + 1]; i++) { const uint32_t v = A->col_idx[i]; if (!vis2[v]) { vis2[v] = 1; queue[qt++] = v; } }
        }
        qh = qt = 0; queue[qt++] = last; visited[last] = 1;
        while (qh < qt) {
            const uint32_t u = queue[qh++]; order[no++] = u;
            uint64_t k = 0;
            for (uint64_t i = A->row_ptr[u]; i < A->row_ptr[u + 1]; i++) { const uint32_t v = A->col_idx[i]; if (!visited[v]) nbrs[k++] = v; }
            for (uint64_t a = 1; a < k; a++) {                      /* insertion sort by degree: stable */
                const uint32_t v = nbrs[a]; const uint64_t dv = A->row_ptr[v + 1] - A->row_ptr[v];
                uint64_t b = a;
                while (b > 0 && A->row_ptr[nbrs[b - 1] + 1] - A->row_ptr[nbrs[b - 1]] > dv) { nbrs[b] = nbrs[b - 1]; b--; }
                nbrs[b] = v;
            }
            for (uint64_t a = 0; a < k; a++) { const uint32_t v = nbrs[a]; if (!visited[v]) { visited[v] = 1; queue[qt++] = v; } }
        }
    }
    for (uint64_t i = 0; i < n; i++) perm[i] = order[n - 1 - i];
    free(visited); free(vis2); free(queue); free(order); free(nbrs);
}

/* CsrMatrix::permute -- graph_csr.rs:727-785: perm[new] = old; rows moved, columns relabelled through the inverse,
 * every row re-sorted by column. */
static int cmp_pair_u64(const void *a, const void *b) { const uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b; return x < y ? -1 : x > y; }
ocsr_t *oracle_permute(const ocsr_t *A, const uint32_t *perm) {
    const uint64_t n = A->rows;
    uint32_t *inv = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
    for (uint64_t i = 0; i < n; i++) inv[perm[i]] = (uint32_t)i;
    ocsr_t *C = ocsr_alloc(n, A->cols, A->nnz, A->val_bits);
    for (uint64_t nr = 0; nr < n; nr++) C->row_ptr[nr + 1] = C->row_ptr[nr] + (A->row_ptr[perm[nr] + 1] - A->row_ptr[perm[nr]]);
    const size_t vb = (size_t)A->val_bits / 8;
    uint64_t maxlen = 1;
    for (uint64_t r = 0; r < n; r++) if (A->row_ptr[r + 1] - A->row_ptr[r] > maxlen) maxlen = A->row_ptr[r + 1] - A->row_ptr[r];
    uint64_t *pairs = (uint64_t *)malloc(maxlen * sizeof(uint64_t));      /* (new column << 32 | position in the old row) */
    for (uint64_t nr = 0; nr < n; nr++) {
        const uint64_t os = A->row_ptr[perm[nr]], len = A->row_ptr[perm[nr] + 1] - os, ns = C->row_ptr[nr];
        for (uint64_t j = 0; j < len; j++) pairs[j] = ((uint64_t)inv[A->col_idx[os + j]] << 32) | j;
        qsort(pairs, len, sizeof(uint64_t), cmp_pair_u64);
        for (uint64_t j = 0; j < len; j++) {
            C->col_idx[ns + j] = (uint32_t)(pairs[j] >> 32);
            memcpy((char *)C->values + (ns + j) * vb, (const char *)A->values + (os + (pairs[j] & 0xFFFFFFFFu)) * vb, vb);
        }
    }
    free(pairs); free(inv);
    return C;
}

/* CsrMatrix::bandwidth_stats -- graph_csr.rs:802-818: (max |r-c|, sum |r-c| / max(count, 1)). */
void oracle_bandwidth_stats(const ocsr_t *A, uint64_t *max_bw, double *avg_bw) {
    uint64_t mx = 0, sum = 0, cnt = 0;
    for (uint64_t r = 0; r < A->rows; r++)
        for (uint64_t i = A->row_ptr[r]; i < A->row_ptr[r + 1]; i++) {
            const uint64_t c = A->col_idx[i], d = r > c ? r - c : c - r;
            if (d > mx) mx = d;
            sum += d; cnt++;
        }
    *max_bw = mx; *avg_bw = (double)sum / (double)(cnt ? cnt : 1);
}

/* The reference's protocol (graph_magnus.rs:758-772): wall clock around ITERS multiplies,
 * results dropped (allocation + free inside the timed region). Returns seconds per iter. */
static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
double oracle_time_matmul(const ocsr_t *A, const ocsr_t *B, int par, int nthreads, int iters) {
    double t0 = now_s();
    for (int it = 0; it < iters; it++) {
        ocsr_t *C;
        if (A->val_bits == 32) C = par ? oracle_matmul_par_u32(A, B, nthreads) : oracle_matmul_seq_u32(A, B);
        else C = par ? oracle_matmul_par_u64(A, B, nthreads) : oracle_matmul_seq_u64(A, B);
        oracle_free(C);
    }
    return (now_s() - t0) / (iters > 0 ? iters : 1);
}
int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
