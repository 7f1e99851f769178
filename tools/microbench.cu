// Shared-memory atomic throughput probe (design input for the numeric accumulator choice).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/microbench.cu -o tools/microbench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long ull;

template <int OP>
__global__ void __launch_bounds__(256) k_atom(int iters, int slots, unsigned *sink) {
    extern __shared__ unsigned sm[];
    for (int i = threadIdx.x; i < slots * 2; i += blockDim.x) sm[i] = (OP == 2) ? 0xFFFFFFFFu : 0u;
    __syncthreads();
    unsigned x = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 12345u;
    unsigned acc = 0;
    for (int it = 0; it < iters; it++) {
        x = x * 1664525u + 1013904223u;
        unsigned h = (x >> 8) & (slots - 1);
        if (OP == 0) atomicAdd(&sm[h], x | 1u);                      // u32 add, no return
        else if (OP == 1) acc += atomicOr(&sm[h], 1u << (x & 31)); // u32 or with return
        else if (OP == 2) acc += atomicCAS(&sm[h], 0xFFFFFFFFu, x); // u32 CAS with return
        else if (OP == 3) atomicAdd((ull *)&sm[2 * h], (ull)x);   // u64 add, no return
        else if (OP == 4) acc += sm[h];                            // plain load
        else if (OP == 5) sm[h] = x;                               // plain store
        else if (OP == 6) atomicAdd(&sm[(threadIdx.x * 33 + it) & (slots - 1)], x | 1u); // conflict-free add
        else if (OP == 7) { unsigned k = sm[h]; if (k != x) { if (k == 0xFFFFFFFFu) acc += atomicCAS(&sm[h], 0xFFFFFFFFu, x); } atomicAdd(&sm[slots + h], x | 1u); } // hash-like
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

template <int OP>
void run(const char *name, int slots) {
    int iters = 4096, grid = 148 * 8;
    unsigned *sink; cudaMalloc(&sink, 4);
    cudaFuncSetAttribute(k_atom<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    size_t smem = (size_t)slots * 8;
    k_atom<OP><<<grid, 256, smem>>>(16, slots, sink);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k_atom<OP><<<grid, 256, smem>>>(iters, slots, sink);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double ops = (double)grid * 256 * iters;
    int per_sm; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_atom<OP>, 256, smem);
    printf("%-28s slots=%5d ctas/sm=%d  %.3f ms  %.1f Gop/s  %.2f lane-ops/clk/SM @1.9GHz  err=%s\n", name, slots, per_sm, ms, ops / ms / 1e6,
           ops / (ms * 1e-3) / 148 / 1.9e9, cudaGetErrorString(cudaGetLastError()));
    cudaFree(sink);
}

int main() {
    for (int slots : {256, 2048, 8192}) {
        run<0>("atomicAdd u32 (RED)", slots);
        run<1>("atomicOr u32 (ret)", slots);
        run<2>("atomicCAS u32 (ret)", slots);
        run<3>("atomicAdd u64 (RED)", slots);
        run<4>("plain LDS random", slots);
        run<5>("plain STS random", slots);
        run<6>("atomicAdd u32 conflict-free", slots);
        run<7>("hash-like ld+cas+add", slots);
    }
    return 0;
}
