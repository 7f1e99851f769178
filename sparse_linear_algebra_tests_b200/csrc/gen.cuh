// gen.cuh -- fixture generators on the device: the reference's Moore lattice / torus and its symmetric
// Bernoulli thinning, bit-identical to the host builders in hostgen.py / oracle, so that 8 M-node inputs are not host-bound.
//
//   k_lattice       CsrMatrix::lattice (src/graph_csr.rs:177-222): N-d Moore neighbourhood, node ids row-major with the last
//                   dimension fastest, the 3^N offsets enumerated with dimension 0 as the least-significant digit, torus wrap
//                   via rem_euclid, the all-zero offset skipped, duplicate neighbours (side-2 torus) summed as from_coo does
//   k_thin_*        CsrMatrix::thin (src/graph_csr.rs:225-247) with StdRng::from_seed(seed): entries are visited row-major, one
//                   draw per stored entry with r <= c, a kept (r,c) also keeps its stored mirror (c,r).  StdRng is ChaCha12
//                   (rand 0.9 / rand_chacha 0.9) and counter based, so draw k is computed where it is needed: next_u64 number
//                   k is keystream words 2k, 2k+1 (low word first), random_range(0.0..1.0) is (u64 >> 12) * 2^-52.
#pragma once
#include "common.cuh"

struct ChaChaKey { u32 k[8]; };

__host__ __device__ __forceinline__ u32 b200_rotl32(u32 x, int n) { return (x << n) | (x >> (32 - n)); }

// one ChaCha12 block (stream id 0, 64-bit block counter): 16 little-endian keystream words
__host__ __device__ inline void chacha12_block(const ChaChaKey &key, u64 counter, u32 out[16]) {
    u32 s[16] = {0x61707865u, 0x3320646Eu, 0x79622D32u, 0x6B206574u,
                 key.k[0], key.k[1], key.k[2], key.k[3], key.k[4], key.k[5], key.k[6], key.k[7],
                 (u32)counter, (u32)(counter >> 32), 0u, 0u};
    u32 x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = s[i];
#define B200_QR(a, b, c, d)                                                    \
    x[a] += x[b]; x[d] = b200_rotl32(x[d] ^ x[a], 16); x[c] += x[d]; x[b] = b200_rotl32(x[b] ^ x[c], 12); \
    x[a] += x[b]; x[d] = b200_rotl32(x[d] ^ x[a], 8);  x[c] += x[d]; x[b] = b200_rotl32(x[b] ^ x[c], 7);
#pragma unroll
    for (int r = 0; r < 6; r++) {
        B200_QR(0, 4, 8, 12) B200_QR(1, 5, 9, 13) B200_QR(2, 6, 10, 14) B200_QR(3, 7, 11, 15)
        B200_QR(0, 5, 10, 15) B200_QR(1, 6, 11, 12) B200_QR(2, 7, 8, 13) B200_QR(3, 4, 9, 14)
    }
#undef B200_QR
#pragma unroll
    for (int i = 0; i < 16; i++) out[i] = x[i] + s[i];
}

// StdRng's k-th next_u64 (k counted from the seeding)
__host__ __device__ inline u64 stdrng_u64_at(const ChaChaKey &key, u64 k) {
    u32 w[16];
    chacha12_block(key, k >> 3, w);
    const int o = (int)(k & 7) * 2;
    return (u64)w[o] | ((u64)w[o + 1] << 32);
}
// the reference's `rng.random_range(0.0..1.0) < density` for draw k
__host__ __device__ inline bool stdrng_keep(const ChaChaKey &key, u64 k, double density) {
    return (double)(stdrng_u64_at(key, k) >> 12) * (1.0 / 4503599627370496.0) < density;
}

#ifdef __CUDACC__
#define B200_LATTICE_MAXD 4
struct LatticeDims { u64 dim[B200_LATTICE_MAXD]; u64 stride[B200_LATTICE_MAXD]; int nd; int torus; };

// thread per node: its (sorted, duplicate-summed) neighbour list; FILL = false counts, FILL = true writes
template <typename VT, bool FILL>
__global__ void __launch_bounds__(128) k_lattice(u64 total, LatticeDims L, u32 *__restrict__ nnz_row, const u64 *__restrict__ rp,
                                                 u32 *__restrict__ col, VT *__restrict__ val) {
    const u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= total) return;
    long long c[B200_LATTICE_MAXD];
    u64 rem = r;
    for (int d = L.nd - 1; d >= 0; d--) { c[d] = (long long)(rem % L.dim[d]); rem /= L.dim[d]; }
    int noff = 1;
    for (int d = 0; d < L.nd; d++) noff *= 3;
    u32 nb[81];
    int n = 0;
    for (int off = 0; off < noff; off++) {
        int tmp = off; bool self = true, valid = true; u64 id = 0;
        for (int d = 0; d < L.nd; d++) {
            const int delta = tmp % 3 - 1; tmp /= 3;
            if (delta) self = false;
            long long cd = c[d] + delta;
            const long long dim = (long long)L.dim[d];
            if (L.torus) { cd %= dim; if (cd < 0) cd += dim; }
            else if (cd < 0 || cd >= dim) valid = false;
            id += (u64)(valid ? cd : 0) * L.stride[d];
        }
        if (self || !valid) continue;
        // insertion into the sorted list
        int j = n++;
        const u32 v = (u32)id;
        while (j > 0 && nb[j - 1] > v) { nb[j] = nb[j - 1]; j--; }
        nb[j] = v;
    }
    if (!FILL) {
        u32 runs = 0;
        for (int j = 0; j < n; j++) runs += (j == 0 || nb[j] != nb[j - 1]);
        nnz_row[r] = runs;
    } else {
        u64 o = rp[r];
        for (int j = 0; j < n;) {
            int e = j + 1;
            while (e < n && nb[e] == nb[j]) e++;
            col[o] = nb[j]; val[o] = (VT)(e - j); o++;
            j = e;
        }
    }
}

__device__ __forceinline__ u64 row_lower_bound(const u32 *__restrict__ col, u64 s, u64 e, u32 key) {
    while (s < e) { const u64 m = (s + e) >> 1; if (col[m] < key) s = m + 1; else e = m; }
    return s;
}

// entries with column >= row (the ones that draw), per row
__global__ void __launch_bounds__(256) k_thin_upper_count(u64 rows, const u64 *__restrict__ rp, const u32 *__restrict__ col, u32 *__restrict__ nnz_row) {
    const u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const u64 s = rp[r], e = rp[r + 1];
    const u64 f = r > 0xFFFFFFFFull ? e : row_lower_bound(col, s, e, (u32)r);
    nnz_row[r] = (u32)(e - f);
}

// thread per row: keep decision of every entry (own draw above the diagonal, the mirror's draw below it)
template <typename VT, bool FILL>
__global__ void __launch_bounds__(256) k_thin_rows(u64 rows, const u64 *__restrict__ rp, const u32 *__restrict__ col, const VT *__restrict__ val,
                                                   const u64 *__restrict__ base_u, ChaChaKey key, u64 skip, double density,
                                                   u32 *__restrict__ nnz_row, const u64 *__restrict__ rpC, u32 *__restrict__ colC, VT *__restrict__ valC) {
    const u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const u64 s = rp[r], e = rp[r + 1];
    const u64 fu = r > 0xFFFFFFFFull ? e : row_lower_bound(col, s, e, (u32)r);
    u32 kept = 0;
    u64 o = FILL ? rpC[r] : 0;
    for (u64 j = s; j < e; j++) {
        const u32 c = col[j];
        bool keep;
        if (j >= fu) keep = stdrng_keep(key, skip + base_u[r] + (j - fu), density);
        else {
            // entry (r, c) below the diagonal: kept iff its mirror (c, r) is stored and was kept by its own draw
            const u64 ms = rp[c], me = rp[(u64)c + 1];
            const u64 mj = row_lower_bound(col, ms, me, (u32)r);
            keep = false;
            if (mj < me && col[mj] == (u32)r) {
                const u64 mfu = row_lower_bound(col, ms, me, c);
                keep = stdrng_keep(key, skip + base_u[c] + (mj - mfu), density);
            }
        }
        if (keep) {
            if (FILL) { colC[o] = c; valC[o] = val[j]; o++; }
            kept++;
        }
    }
    if (!FILL) nnz_row[r] = kept;
}
#endif  // __CUDACC__
