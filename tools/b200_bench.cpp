// b200_bench -- compiled host above the C ABI (include/b200_spgemm.h): the role the reference's Rust benchmark functions play
// upstream, for a maintainer without Python.  One process drives N GPUs (a host thread and a b200_ctx per GPU, NCCL
// communicators from b200_comm_init_all); nothing here touches CUDA or NCCL directly.
//
//   --config torus30   bench_repeated_exponentiation (/root/reference/src/graph_magnus.rs:699-788): A^2..A^7 on the 30^3 Moore
//                      torus thinned to ~3 e/n with StdRng([42;32]); checks nnz per power against the README column
//   --config torus200  BASELINE configs[4]: A^2..A^5 on the 200^3 torus, row blocks resident per GPU
//   --config rmat      BASELINE configs[3]: A^2 of an R-MAT graph (--scale, --ef, --abc), rows sharded by product count
//   --config sweep     bench_matmul_magnus (:790-929): side x e_per_n grid from one shared StdRng, A x A, 1 warm-up + 10 timed
// Multi-GPU protocol: GPU 0 builds the operand on the device, one broadcast replicates it, every GPU cuts the
// same product-balanced row blocks, keeps its block of every power resident and multiplies it by the replicated operand;
// time = max over GPUs of the host clock around a synchronised region.  Output: CSV on stdout, one summary line per config.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "b200_spgemm.h"

#define CHECK(expr) do { int _r = (expr); if (_r != B200_OK) { fprintf(stderr, "b200_bench: %s failed (%d): %s\n", #expr, _r, b200_last_error()); exit(1); } } while (0)

struct Barrier {                                 // C++17: no std::barrier
    std::mutex m; std::condition_variable cv; int n, waiting = 0; long gen = 0;
    explicit Barrier(int n_) : n(n_) {}
    void wait() {
        std::unique_lock<std::mutex> l(m);
        const long g = gen;
        if (++waiting == n) { waiting = 0; gen++; cv.notify_all(); }
        else cv.wait(l, [&] { return gen != g; });
    }
};
static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct Args {
    std::string config = "torus30";
    int gpus = 1, side = 0, max_power = 0, scale = 20, bits = 64, iters = 5, warmup = 2;
    uint64_t ef = 16;
    double a = 0.45, b = 0.15, c = 0.15, epn = 3.0;
};

struct Shared {
    Args args; int n;
    std::vector<b200_ctx *> ctx; std::vector<b200_comm *> comm;
    Barrier bar;
    std::vector<std::vector<double>> ms;         // [rank][multiply]: best time
    std::vector<std::vector<uint64_t>> nnz, prod; // [rank][multiply]
    explicit Shared(const Args &a) : args(a), n(a.gpus), ctx(a.gpus), comm(a.gpus, nullptr), bar(a.gpus), ms(a.gpus), nnz(a.gpus), prod(a.gpus) {}
};

static b200_csr *build_torus(b200_ctx *ctx, int side, double epn, int bits) {
    const uint64_t dims[3] = {(uint64_t)side, (uint64_t)side, (uint64_t)side};
    b200_csr *full = nullptr, *a = nullptr;
    CHECK(b200_lattice(ctx, dims, 3, 1, bits, &full));
    uint8_t seed[32]; memset(seed, 42, 32);
    CHECK(b200_thin(ctx, full, epn / 26.0, seed, 0, &a, nullptr));
    CHECK(b200_csr_free(ctx, full));
    return a;
}

// chain of `steps` multiplies on this GPU's row block; multiply k is p <- p x a
static void rank_main(Shared *S, int rank) {
    const Args &A = S->args;
    b200_ctx *ctx = S->ctx[rank];
    b200_csr *a = nullptr;
    if (rank == 0) {
        const double t0 = now_ms();
        if (A.config == "rmat") CHECK(b200_rmat(ctx, A.scale, A.ef, A.a, A.b, A.c, 42, A.bits, &a));
        else a = build_torus(ctx, A.side, A.epn, A.bits);
        uint64_t rows, nnz; CHECK(b200_csr_info(a, &rows, nullptr, &nnz, nullptr));
        fprintf(stderr, "operand: %llu nodes, %llu entries, built on GPU 0 in %.1f ms\n", (unsigned long long)rows, (unsigned long long)nnz, now_ms() - t0);
    }
    if (S->n > 1) {
        b200_csr *rep = nullptr;
        CHECK(b200_comm_broadcast_csr(S->comm[rank], a, 0, &rep));
        if (rank == 0) CHECK(b200_csr_free(ctx, a));
        a = rep;
    }
    uint64_t rows = 0; CHECK(b200_csr_info(a, &rows, nullptr, nullptr, nullptr));
    b200_csr *blk = a;
    if (S->n > 1) {
        std::vector<uint64_t> cuts(S->n + 1);
        CHECK(b200_shard_rows_by_products(ctx, a, a, S->n, cuts.data()));
        CHECK(b200_csr_row_block(ctx, a, cuts[rank], cuts[rank + 1], &blk));
    }
    const int steps = A.config == "rmat" ? 1 : A.max_power - 1;
    S->ms[rank].assign(steps, 1e30); S->nnz[rank].assign(steps, 0); S->prod[rank].assign(steps, 0);
    CHECK(b200_ctx_set_timing(ctx, 0));
    for (int it = 0; it < A.warmup + A.iters; it++) {
        std::vector<b200_csr *> keep;                // every power stays alive, as in the reference loop (:737-786)
        b200_csr *p = blk;
        for (int k = 0; k < steps; k++) {
            CHECK(b200_ctx_synchronize(ctx));
            S->bar.wait();
            const double t0 = now_ms();
            b200_csr *c = nullptr;
            CHECK(b200_spgemm(ctx, p, a, &c, nullptr));
            CHECK(b200_ctx_synchronize(ctx));
            const double dt = now_ms() - t0;
            if (it >= A.warmup) S->ms[rank][k] = std::min(S->ms[rank][k], dt);
            if (it == A.warmup + A.iters - 1) {
                b200_stats st; CHECK(b200_csr_product_stats(ctx, c, &st));
                S->nnz[rank][k] = st.nnz_c; S->prod[rank][k] = st.products;
            }
            keep.push_back(c); p = c;
        }
        for (b200_csr *c : keep) CHECK(b200_csr_free(ctx, c));
    }
    if (blk != a) CHECK(b200_csr_free(ctx, blk));
    CHECK(b200_csr_free(ctx, a));
}

static int run_chain(const Args &args) {
    Shared S(args);
    for (int g = 0; g < S.n; g++) CHECK(b200_ctx_create(g, nullptr, &S.ctx[g]));
    if (S.n > 1) CHECK(b200_comm_init_all(S.ctx.data(), S.n, S.comm.data()));
    std::vector<std::thread> th;
    for (int g = 0; g < S.n; g++) th.emplace_back(rank_main, &S, g);
    for (auto &t : th) t.join();
    const int steps = (int)S.ms[0].size();
    // the reference's chain table starts `step,nnz,...` (src/graph_magnus.rs:734); `multiply` is its step, times are the max over GPUs
    printf("config,gpus,multiply,nnz,products,ms_max_over_gpus,products_per_s\n");
    double tot_ms = 0; uint64_t tot_prod = 0;
    std::vector<uint64_t> nnz_k(steps, 0);
    for (int k = 0; k < steps; k++) {
        double ms = 0; uint64_t nnz = 0, prod = 0;
        for (int g = 0; g < S.n; g++) { ms = std::max(ms, S.ms[g][k]); nnz += S.nnz[g][k]; prod += S.prod[g][k]; }
        nnz_k[k] = nnz; tot_ms += ms; tot_prod += prod;
        printf("%s,%d,%s%d,%llu,%llu,%.4f,%.4g\n", args.config.c_str(), S.n, args.config == "rmat" ? "A^" : "A^", k + 2, (unsigned long long)nnz,
               (unsigned long long)prod, ms, prod / (ms * 1e-3));
    }
    printf("# %s on %d GPU(s): %llu intermediate products in %.3f ms = %.3f G products/s (best of %d, host clock around a synchronised multiply, max over GPUs)\n",
           args.config.c_str(), S.n, (unsigned long long)tot_prod, tot_ms, tot_prod / (tot_ms * 1e-3) / 1e9, args.iters);
    int rc = 0;
    if (args.config == "torus30" && args.side == 30 && args.max_power == 7 && args.epn == 3.0) {
        const uint64_t want[6] = {251590, 655391, 1574848, 3383207, 6590100, 11736555};   // README.md:42-47 (252 k .. 11.7 M) with the exact StdRng stream
        bool ok = true;
        for (int k = 0; k < 6; k++) ok &= nnz_k[k] == want[k];
        printf("# nnz check %s\n", ok ? "OK" : "FAILED");
        rc = ok ? 0 : 2;
    }
    for (int g = 0; g < S.n; g++) { if (S.comm[g]) b200_comm_destroy(S.comm[g]); b200_ctx_destroy(S.ctx[g]); }
    return rc;
}

// bench_matmul_magnus: side x e_per_n grid, every instance thinned from the full lattice with ONE generator shared over the grid
static int run_sweep(const Args &args) {
    b200_ctx *ctx = nullptr;
    CHECK(b200_ctx_create(0, nullptr, &ctx));
    const int sides[4] = {5, 10, 20, 30};
    const double epns[5] = {2, 3, 4, 8, 26};
    uint8_t seed[32]; memset(seed, 42, 32);
    uint64_t skip = 0;
    // the reference's columns (src/graph_magnus.rs:805) up to `components`, then this engine's: mean and best of the 10 timed
    // multiplies in microseconds (the reference prints elapsed / ITERS), the product's nnz and the intermediate products
    printf("side,nodes,e_per_n,nnz,components,b200_us,b200_best_us,nnz_c,products\n");
    for (int si = 0; si < 4; si++) {
        const uint64_t dims[3] = {(uint64_t)sides[si], (uint64_t)sides[si], (uint64_t)sides[si]};
        b200_csr *full = nullptr;
        CHECK(b200_lattice(ctx, dims, 3, 1, args.bits, &full));
        for (int ei = 0; ei < 5; ei++) {
            b200_csr *a = nullptr; uint64_t draws = 0;
            CHECK(b200_thin(ctx, full, std::min(1.0, epns[ei] / 26.0), seed, skip, &a, &draws));
            skip += draws;
            uint64_t rows, nnz_a; CHECK(b200_csr_info(a, &rows, nullptr, &nnz_a, nullptr));
            double best = 1e30, sum = 0; b200_stats st; memset(&st, 0, sizeof(st));
            for (int it = 0; it < 11; it++) {                        // 1 warm-up + 10 timed (:856)
                b200_csr *c = nullptr;
                CHECK(b200_ctx_synchronize(ctx));
                const double t0 = now_ms();
                CHECK(b200_spgemm(ctx, a, a, &c, nullptr));
                CHECK(b200_ctx_synchronize(ctx));
                const double dt = now_ms() - t0;
                if (it) { best = std::min(best, dt); sum += dt; }
                if (it == 10) CHECK(b200_csr_product_stats(ctx, c, &st));
                CHECK(b200_csr_free(ctx, c));
            }
            // num_components (src/graph_csr.rs:654-657) of the instance: union-find over the downloaded pattern
            std::vector<uint64_t> rp(rows + 1); std::vector<uint32_t> ci(std::max<uint64_t>(nnz_a, 1)); std::vector<uint64_t> vv(std::max<uint64_t>(nnz_a, 1));
            CHECK(b200_csr_download(ctx, a, rp.data(), ci.data(), args.bits == 64 ? (void *)vv.data() : (void *)vv.data()));
            std::vector<uint32_t> parent(rows);
            for (uint64_t i = 0; i < rows; i++) parent[i] = (uint32_t)i;
            auto find = [&](uint32_t x) { while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; } return x; };
            for (uint64_t r = 0; r < rows; r++) for (uint64_t i = rp[r]; i < rp[r + 1]; i++) { const uint32_t x = find((uint32_t)r), y = find(ci[i]); if (x != y) parent[x] = y; }
            uint64_t components = 0;
            for (uint64_t i = 0; i < rows; i++) components += find((uint32_t)i) == i;
            printf("%d,%llu,%.0f,%llu,%llu,%.1f,%.1f,%llu,%llu\n", sides[si], (unsigned long long)rows, epns[ei], (unsigned long long)nnz_a, (unsigned long long)components,
                   sum / 10 * 1e3, best * 1e3, (unsigned long long)st.nnz_c, (unsigned long long)st.products);
            CHECK(b200_csr_free(ctx, a));
        }
        CHECK(b200_csr_free(ctx, full));
    }
    b200_ctx_destroy(ctx);
    return 0;
}

int main(int argc, char **argv) {
    Args a;
    for (int i = 1; i < argc; i++) {
        const std::string k = argv[i];
        auto val = [&]() -> const char * { if (i + 1 >= argc) { fprintf(stderr, "b200_bench: %s needs a value\n", k.c_str()); exit(1); } return argv[++i]; };
        if (k == "--config") a.config = val();
        else if (k == "--gpus") a.gpus = atoi(val());
        else if (k == "--side") a.side = atoi(val());
        else if (k == "--max-power") a.max_power = atoi(val());
        else if (k == "--scale") a.scale = atoi(val());
        else if (k == "--ef") a.ef = strtoull(val(), nullptr, 10);
        else if (k == "--abc") { a.a = atof(val()); a.b = atof(val()); a.c = atof(val()); }
        else if (k == "--epn") a.epn = atof(val());
        else if (k == "--bits") a.bits = atoi(val());
        else if (k == "--iters") a.iters = atoi(val());
        else if (k == "--warmup") a.warmup = atoi(val());
        else { fprintf(stderr, "usage: b200_bench --config torus30|torus200|rmat|sweep [--gpus N] [--side S] [--max-power K] [--scale S] [--ef E] [--abc a b c] [--epn E] [--bits 32|64] [--iters I] [--warmup W]\n"); return 1; }
    }
    if (a.config == "torus30") { if (!a.side) a.side = 30; if (!a.max_power) a.max_power = 7; }
    else if (a.config == "torus200") { if (!a.side) a.side = 200; if (!a.max_power) a.max_power = 5; }
    else if (a.config != "rmat" && a.config != "sweep") { fprintf(stderr, "b200_bench: unknown config %s\n", a.config.c_str()); return 1; }
    const int have = b200_device_count();
    if (have < 1) { fprintf(stderr, "b200_bench: no CUDA device -- the engine has no CPU fallback\n"); return 1; }
    if (a.gpus < 1 || a.gpus > have) { fprintf(stderr, "b200_bench: --gpus %d but %d device(s) visible\n", a.gpus, have); return 1; }
    if (a.config == "sweep") return run_sweep(a);
    return run_chain(a);
}
