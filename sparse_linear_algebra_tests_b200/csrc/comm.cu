// comm.cu -- multi-GPU below the C ABI: one communicator per (process, GPU) over NCCL / NVLink.
//
// The path shards by rows of the left operand (row i of C needs row i of A and all of B, /root/reference/src/graph_csr.rs:433-446),
// so the data path has exactly two collectives, both outside the multiply:
//   b200_comm_broadcast_csr   replicate the right operand once (header, row_ptr, col_idx, values: four ncclBroadcast);
//   b200_comm_allgather_csr   assemble C (or the next right operand of a squaring chain, power_until_stable at
//                             src/graph_csr.rs:561-575) from the per-rank row blocks: sizes by ncclAllGather, the three
//                             arrays by one grouped ncclBroadcast per rank straight into their place in the result (blocks
//                             differ in size), row_ptr segments re-based by a kernel.
// plus b200_comm_allreduce for the scalars a report needs (max time, total products).
// NCCL is bound at run time (dlopen of libnccl.so.2): single-GPU users of the library need no NCCL at all, and inside a
// torchrun process the copy torch already loaded is the one that gets used.  Without it every entry point here returns
// B200_ERR_NCCL.  Two ways to form the communicators: b200_comm_init_rank (one process per GPU; the 128-byte id from
// b200_comm_unique_id travels through the launcher's own channel -- torch.distributed store, MPI, a file) and
// b200_comm_init_all (one process driving N GPUs, e.g. tools/b200_bench.cpp with a thread per GPU).
#include <dlfcn.h>
#include <mutex>
#include "engine.cuh"
#include "devutil.cuh"

// ---- the slice of nccl.h this file needs (NCCL 2.x ABI: /usr/include/nccl.h)
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclUint8 = 1, ncclUint32 = 3, ncclUint64 = 5, ncclFloat64 = 8 };
enum { ncclSum = 0, ncclMax = 2, ncclMin = 3 };
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    const char *(*GetErrorString)(ncclResult_t);
    bool ok;
};
static NcclApi g_nccl;
static std::once_flag g_nccl_once;
static const NcclApi *nccl() {
    std::call_once(g_nccl_once, [] {
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
#define BIND(field, sym) *(void **)(&g_nccl.field) = dlsym(h, sym)
        BIND(GetUniqueId, "ncclGetUniqueId"); BIND(CommInitRank, "ncclCommInitRank"); BIND(CommInitAll, "ncclCommInitAll");
        BIND(CommDestroy, "ncclCommDestroy"); BIND(Broadcast, "ncclBroadcast"); BIND(AllGather, "ncclAllGather");
        BIND(AllReduce, "ncclAllReduce"); BIND(GroupStart, "ncclGroupStart"); BIND(GroupEnd, "ncclGroupEnd"); BIND(GetErrorString, "ncclGetErrorString");
#undef BIND
        g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommInitAll && g_nccl.CommDestroy && g_nccl.Broadcast && g_nccl.AllGather &&
                    g_nccl.AllReduce && g_nccl.GroupStart && g_nccl.GroupEnd && g_nccl.GetErrorString;
    });
    return g_nccl.ok ? &g_nccl : nullptr;
}
#define NCCL_API(var) const NcclApi *var = nccl(); if (!var) return set_err(B200_ERR_NCCL, "NCCL is not available (libnccl.so.2 could not be loaded)")
#define NCCL_TRY(api, expr) do { ncclResult_t _r = (expr); if (_r != 0) return set_err(B200_ERR_NCCL, "%s failed: %s (%s:%d)", #expr, (api)->GetErrorString(_r), __FILE__, __LINE__); } while (0)

struct b200_comm { b200_ctx *ctx; ncclComm_t comm; int rank, size; u64 *d_small; };   // d_small: 64 device words for headers / scalars

__global__ void __launch_bounds__(256) k_offset_rowptr(u64 n, u64 *rp, u64 add) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) rp[i] += add;
}

extern "C" int b200_comm_unique_id(uint8_t id128[128]) {
    if (!id128) return set_err(B200_ERR_BADARG, "NULL argument");
    NCCL_API(api);
    ncclUniqueId id;
    NCCL_TRY(api, api->GetUniqueId(&id));
    memcpy(id128, id.internal, 128);
    return B200_OK;
}

static int comm_finish(b200_ctx *ctx, ncclComm_t c, int rank, int size, b200_comm **out) {
    b200_comm *m = new b200_comm();
    m->ctx = ctx; m->comm = c; m->rank = rank; m->size = size; m->d_small = nullptr;
    cudaSetDevice(ctx->device);
    if (cudaMalloc((void **)&m->d_small, 64 * 8 + (size_t)size * 64) != cudaSuccess) { delete m; return set_err(B200_ERR_ALLOC, "communicator scratch allocation failed"); }
    *out = m;
    return B200_OK;
}

extern "C" int b200_comm_init_rank(b200_ctx *ctx, int nranks, int rank, const uint8_t id128[128], b200_comm **out) {
    if (!ctx || !id128 || !out || nranks < 1 || rank < 0 || rank >= nranks) return set_err(B200_ERR_BADARG, "bad communicator arguments");
    NCCL_API(api);
    CUDA_TRY(cudaSetDevice(ctx->device));
    ncclUniqueId id; memcpy(id.internal, id128, 128);
    ncclComm_t c;
    NCCL_TRY(api, api->CommInitRank(&c, nranks, id, rank));
    return comm_finish(ctx, c, rank, nranks, out);
}

extern "C" int b200_comm_init_all(b200_ctx **ctxs, int ngpus, b200_comm **out) {
    if (!ctxs || !out || ngpus < 1) return set_err(B200_ERR_BADARG, "bad communicator arguments");
    NCCL_API(api);
    std::vector<int> devs(ngpus);
    for (int i = 0; i < ngpus; i++) { if (!ctxs[i]) return set_err(B200_ERR_BADARG, "NULL context"); devs[i] = ctxs[i]->device; }
    std::vector<ncclComm_t> cs(ngpus);
    NCCL_TRY(api, api->CommInitAll(cs.data(), ngpus, devs.data()));
    for (int i = 0; i < ngpus; i++) TRY(comm_finish(ctxs[i], cs[i], i, ngpus, &out[i]));
    return B200_OK;
}

extern "C" int b200_comm_destroy(b200_comm *c) {
    if (!c) return B200_OK;
    const NcclApi *api = nccl();
    cudaSetDevice(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream);
    if (api) api->CommDestroy(c->comm);
    cudaFree(c->d_small);
    delete c;
    return B200_OK;
}
extern "C" int b200_comm_rank(const b200_comm *c, int *rank, int *size) {
    if (!c) return set_err(B200_ERR_BADARG, "NULL argument");
    if (rank) *rank = c->rank;
    if (size) *size = c->size;
    return B200_OK;
}

// op: 0 sum of u64, 1 max of u64, 2 sum of f64, 3 max of f64; in place on a host array of n <= 32 scalars
extern "C" int b200_comm_allreduce(b200_comm *c, void *host_scalars, int n, int op) {
    if (!c || !host_scalars || n < 1 || n > 32 || op < 0 || op > 3) return set_err(B200_ERR_BADARG, "bad allreduce arguments");
    NCCL_API(api);
    CUDA_TRY(cudaSetDevice(c->ctx->device));
    cudaStream_t s = c->ctx->stream;
    CUDA_TRY(cudaMemcpyAsync(c->d_small, host_scalars, (size_t)n * 8, cudaMemcpyHostToDevice, s));
    NCCL_TRY(api, api->AllReduce(c->d_small, c->d_small, (size_t)n, op < 2 ? ncclUint64 : ncclFloat64, (op & 1) ? ncclMax : ncclSum, c->comm, s));
    CUDA_TRY(cudaMemcpyAsync(host_scalars, c->d_small, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return B200_OK;
}

// Replicate `src` (given on rank `root`, NULL elsewhere) on every rank of the communicator.
extern "C" int b200_comm_broadcast_csr(b200_comm *c, const b200_csr *src, int root, b200_csr **out) {
    if (!c || !out || root < 0 || root >= c->size || (c->rank == root && !src)) return set_err(B200_ERR_BADARG, "bad broadcast arguments");
    NCCL_API(api);
    b200_ctx *ctx = c->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    u64 hdr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (c->rank == root) {
        RESOLVE(ctx, src);
        if (src->ctx != ctx) return set_err(B200_ERR_BADARG, "the matrix belongs to another context");
        hdr[0] = src->rows; hdr[1] = src->cols; hdr[2] = src->nnz; hdr[3] = (u64)src->val_bits; hdr[4] = src->max_row_len;
        CUDA_TRY(cudaMemcpyAsync(c->d_small, hdr, sizeof(hdr), cudaMemcpyHostToDevice, s));
    }
    NCCL_TRY(api, api->Broadcast(c->d_small, c->d_small, 8, ncclUint64, root, c->comm, s));
    CUDA_TRY(cudaMemcpyAsync(hdr, c->d_small, sizeof(hdr), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    const u64 rows = hdr[0], cols = hdr[1], nnz = hdr[2];
    const int vb = (int)hdr[3];
    if (vb != 32 && vb != 64) return set_err(B200_ERR_NCCL, "broadcast header is corrupt (value width %d)", vb);
    b200_csr *m = nullptr;
    TRY(csr_alloc(ctx, rows, cols, nnz, vb, true, &m));
    const b200_csr *from = c->rank == root ? src : m;
    int r = B200_OK;
    ncclResult_t nr = api->Broadcast(from->d_rp, m->d_rp, rows + 1, ncclUint64, root, c->comm, s);
    if (nr == 0 && nnz) nr = api->Broadcast(from->d_col, m->d_col, nnz, ncclUint32, root, c->comm, s);
    if (nr == 0 && nnz) nr = api->Broadcast(from->d_val, m->d_val, nnz * (size_t)(vb / 8), ncclUint8, root, c->comm, s);
    if (nr != 0) r = set_err(B200_ERR_NCCL, "ncclBroadcast failed: %s", api->GetErrorString(nr));
    if (r == B200_OK) { m->max_row_len = hdr[4]; r = finish_new_csr(ctx, m, true, true); }
    if (r != B200_OK) { b200_csr_free(ctx, m); return r; }
    *out = m;
    return B200_OK;
}

// Row blocks in rank order -> the whole matrix on every rank.  Blocks may differ in rows and entries (and may be empty);
// all must have the same column count and value width.
extern "C" int b200_comm_allgather_csr(b200_comm *c, const b200_csr *block, b200_csr **out) {
    if (!c || !block || !out) return set_err(B200_ERR_BADARG, "NULL argument");
    NCCL_API(api);
    b200_ctx *ctx = c->ctx;
    if (block->ctx != ctx) return set_err(B200_ERR_BADARG, "the block belongs to another context");
    CUDA_TRY(cudaSetDevice(ctx->device));
    RESOLVE(ctx, block);
    cudaStream_t s = ctx->stream;
    const int P = c->size;
    // ---- sizes of every block
    u64 mine[4] = {block->rows, block->nnz, block->cols, (u64)block->val_bits};
    u64 *d_mine = c->d_small, *d_all = c->d_small + 64;
    CUDA_TRY(cudaMemcpyAsync(d_mine, mine, sizeof(mine), cudaMemcpyHostToDevice, s));
    NCCL_TRY(api, api->AllGather(d_mine, d_all, 4, ncclUint64, c->comm, s));
    std::vector<u64> all((size_t)P * 4);
    CUDA_TRY(cudaMemcpyAsync(all.data(), d_all, (size_t)P * 32, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    u64 rows = 0, nnz = 0;
    std::vector<u64> row_off(P + 1, 0), nnz_off(P + 1, 0);
    for (int r = 0; r < P; r++) {
        if (all[4 * r + 2] != block->cols || all[4 * r + 3] != (u64)block->val_bits)
            return set_err(B200_ERR_SHAPE, "allgather: rank %d holds a %llu-column u%llu block, this rank %llu-column u%d", r, (ull)all[4 * r + 2], (ull)all[4 * r + 3], (ull)block->cols, block->val_bits);
        row_off[r + 1] = row_off[r] + all[4 * r]; nnz_off[r + 1] = nnz_off[r] + all[4 * r + 1];
    }
    rows = row_off[P]; nnz = nnz_off[P];
    b200_csr *m = nullptr;
    TRY(csr_alloc(ctx, rows, block->cols, nnz, block->val_bits, true, &m));
    const size_t vb = (size_t)block->val_bits / 8;
    // ---- every rank's three arrays straight into place (one group: the transfers of all ranks overlap)
    ncclResult_t nr = api->GroupStart();
    for (int r = 0; r < P && nr == 0; r++) {
        const u64 br = all[4 * r], bn = all[4 * r + 1];
        const bool me = r == c->rank;
        if (br) nr = api->Broadcast(me ? (const void *)block->d_rp : (const void *)(m->d_rp + row_off[r]), m->d_rp + row_off[r], br, ncclUint64, r, c->comm, s);
        if (nr == 0 && bn) nr = api->Broadcast(me ? (const void *)block->d_col : (const void *)(m->d_col + nnz_off[r]), m->d_col + nnz_off[r], bn, ncclUint32, r, c->comm, s);
        if (nr == 0 && bn) nr = api->Broadcast(me ? (const void *)block->d_val : (const void *)((char *)m->d_val + nnz_off[r] * vb), (char *)m->d_val + nnz_off[r] * vb, bn * vb, ncclUint8, r, c->comm, s);
    }
    if (nr == 0) nr = api->GroupEnd(); else api->GroupEnd();
    if (nr != 0) { b200_csr_free(ctx, m); return set_err(B200_ERR_NCCL, "allgather: NCCL failed: %s", api->GetErrorString(nr)); }
    // ---- row_ptr segments start at their block's first entry; the last word closes the matrix
    for (int r = 0; r < P; r++) {
        const u64 br = all[4 * r];
        if (br && nnz_off[r]) { k_offset_rowptr<<<(unsigned)std::min<u64>((br + 255) / 256, (u64)ctx->num_sms * 8), 256, 0, s>>>(br, m->d_rp + row_off[r], nnz_off[r]); ctx->launches++; }
    }
    CUDA_TRY(cudaMemcpyAsync(m->d_rp + rows, &nnz, 8, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaStreamSynchronize(s));                                    // (`nnz` is a stack word)
    int rr = finish_new_csr(ctx, m, true, true);
    if (rr != B200_OK) { b200_csr_free(ctx, m); return rr; }
    *out = m;
    return B200_OK;
}
