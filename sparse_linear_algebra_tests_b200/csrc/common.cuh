// common.cuh -- shared types, row-bin table and small device helpers of the SpGEMM engine.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

typedef uint32_t u32;
typedef uint64_t u64;
typedef unsigned long long ull;

#define B200_EMPTY_KEY 0xFFFFFFFFu
#define B200_NBINS 24
#define B200_WARP 32

// ---------------------------------------------------------------------------------------
// Row bins (MAGNUS-style categorisation, re-cut for a GPU).  A row of the LEFT operand is
// classified twice: for the symbolic pass by its intermediate-product count P (an upper
// bound of its output nnz) and for the numeric pass by its exact output nnz.
//   bin 0      "tiny"  : P <= 32 and deg_A <= 32 -> one warp, products in registers,
//                        shuffle bitonic sort + segmented saturating reduce.
//   bins 1..8  "hash"  : shared-memory open-addressing table of kHashSlots[bin] slots owned
//                        by kHashThreads[bin] threads (a warp for the small ones, a CTA above).
//   bin 9      "heavy" : does not fit a CTA's shared memory -> global-memory table + bitmap.
// ---------------------------------------------------------------------------------------
#define B200_BIN_TINY 0
#define B200_BIN_HASH0 1
#define B200_NUM_HASH_BINS 8
#define B200_BIN_HEAVY 9
#define B200_BIN_WIDE0 10   // one-pass: hash bin hb's rows whose column window is too wide for the bin's bitmap -> list 10 + hb
#define B200_BIN_NONE 0xFF
#define B200_STAT_BINS 16   // bins reported in b200_stats

// slots per hash bin; a row goes to the first bin whose capacity covers it.
__host__ __device__ __forceinline__ u32 b200_hash_slots(int hb) { return 128u << hb; }          // 128 .. 16384
// symbolic: key-only tables, P <= slots/2 ... except the last bin which takes P <= slots*3/4
// numeric : nnz <= slots/2 (load factor <= 0.5)
__host__ __device__ __forceinline__ u32 b200_hash_cap(int hb) { return 64u << hb; }             // 64 .. 8192
__host__ __device__ __forceinline__ int b200_hash_threads(int hb) {
    // 32,32,64,128,256,256,512,1024
    const int t[B200_NUM_HASH_BINS] = {32, 32, 64, 128, 256, 256, 512, 1024};
    return t[hb];
}

__host__ __device__ __forceinline__ int b200_bin_by_size(u64 sz) {
    // first hash bin with cap >= sz, or heavy
#pragma unroll
    for (int hb = 0; hb < B200_NUM_HASH_BINS; hb++)
        if (sz <= (u64)b200_hash_cap(hb)) return B200_BIN_HASH0 + hb;
    return B200_BIN_HEAVY;
}

// Device control block: everything the host reads back in its single per-multiply sync.
struct B200Ctrl {
    ull total_products;
    ull max_row_products;
    ull total_nnz;
    ull max_row_nnz;
    ull max_val_out;          // running max of emitted C values
    u32 sym_bin_count[B200_NBINS];
    ull total_bound;          // one-pass mode: sum of per-row bounds min(P_i, cols)
    u32 scan_ticket[2];
    u32 error_flag;           // set by kernels on impossible states (table overflow)
    u32 scan_done;            // CTAs of the final row_ptr scan that have finished (the last one reports to the host)
    // fused path (fused.cu)
    u32 num_units;            // work units the pre-pass cut the rows into
    u32 unit_ticket;          // next unit to hand out (persistent numeric kernel)
    u32 fused_done, pre_done; // CTAs that have finished (the last one reports / cleans up)
    u32 class_count[4];       // rows per class: empty, tiny, dense, other
    u32 pad_[2];
};

// Read-only view of a device CSR.
template <typename VT>
struct CsrView {
    u64 rows, cols, nnz;
    const u64 *rp;
    const u32 *col;
    const VT *val;
};

__device__ __forceinline__ u32 b200_hash(u32 c, int shift) { return (c * 0x9E3779B1u) >> shift; }

__device__ __forceinline__ u32 ld_volatile_u32(const u32 *p) { return *reinterpret_cast<const volatile u32 *>(p); }
__device__ __forceinline__ u64 ld_volatile_u64(const u64 *p) { return *reinterpret_cast<const volatile u64 *>(p); }
__device__ __forceinline__ void st_volatile_u64(u64 *p, u64 v) { *reinterpret_cast<volatile u64 *>(p) = v; }

// saturating arithmetic (reference: src/graph_csr.rs:30-37, src/graph_sprs.rs:29-51)
__device__ __forceinline__ u32 sat_add(u32 a, u32 b) { u32 s = a + b; return s < a ? 0xFFFFFFFFu : s; }
__device__ __forceinline__ u64 sat_add(u64 a, u64 b) { u64 s = a + b; return s < a ? ~0ull : s; }
__device__ __forceinline__ u32 sat_mul(u32 a, u32 b) { u64 p = (u64)a * b; return p > 0xFFFFFFFFull ? 0xFFFFFFFFu : (u32)p; }
__device__ __forceinline__ u64 sat_mul(u64 a, u64 b) { return __umul64hi(a, b) ? ~0ull : a * b; }

__device__ __forceinline__ u64 shfl_u64(u64 v, int src) {
    u32 lo = __shfl_sync(0xFFFFFFFFu, (u32)v, src), hi = __shfl_sync(0xFFFFFFFFu, (u32)(v >> 32), src);
    return ((u64)hi << 32) | lo;
}
__device__ __forceinline__ u64 shfl_xor_u64(u64 v, int m) {
    u32 lo = __shfl_xor_sync(0xFFFFFFFFu, (u32)v, m), hi = __shfl_xor_sync(0xFFFFFFFFu, (u32)(v >> 32), m);
    return ((u64)hi << 32) | lo;
}
__device__ __forceinline__ u64 shfl_up_u64(u64 v, int d) {
    u32 lo = __shfl_up_sync(0xFFFFFFFFu, (u32)v, d), hi = __shfl_up_sync(0xFFFFFFFFu, (u32)(v >> 32), d);
    return ((u64)hi << 32) | lo;
}
__device__ __forceinline__ u32 shfl_any(u32 v, int src) { return __shfl_sync(0xFFFFFFFFu, v, src); }
__device__ __forceinline__ u64 shfl_any(u64 v, int src) { return shfl_u64(v, src); }
__device__ __forceinline__ u32 shfl_xor_any(u32 v, int m) { return __shfl_xor_sync(0xFFFFFFFFu, v, m); }
__device__ __forceinline__ u64 shfl_xor_any(u64 v, int m) { return shfl_xor_u64(v, m); }
__device__ __forceinline__ u32 shfl_up_any(u32 v, int d) { return __shfl_up_sync(0xFFFFFFFFu, v, d); }
__device__ __forceinline__ u64 shfl_up_any(u64 v, int d) { return shfl_up_u64(v, d); }

__device__ __forceinline__ u64 warp_max_u64(u64 v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) { u64 o = shfl_xor_u64(v, m); v = o > v ? o : v; }
    return v;
}
__device__ __forceinline__ u32 warp_sum_u32(u32 v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, m);
    return v;
}
__device__ __forceinline__ u64 warp_sum_u64(u64 v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += shfl_xor_u64(v, m);
    return v;
}
