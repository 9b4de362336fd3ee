// Does cp.async.bulk.prefetch.L2 actually make a later read faster on B200?  (standalone microbenchmark)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void k_prefetch_bulk(const uint8_t *p, size_t bytes, unsigned chunk) {
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * chunk;
    if (i < bytes) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p + i), "r"(chunk) : "memory");
}
__global__ void k_prefetch_line(const uint8_t *p, size_t bytes) {
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 128;
    if (i < bytes) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + i) : "memory");
}
__global__ void k_read(const int4 *p, size_t n, int *sink) {
    int acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        int4 v;
        asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p + i));
        acc += v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345678) *sink = acc;
}
int main() {
    const size_t big = 1ull << 30;
    uint8_t *buf, *flush;
    int *sink;
    CK(cudaMalloc(&buf, big)); CK(cudaMalloc(&flush, big)); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(buf, 1, big)); CK(cudaMemset(flush, 2, big));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (size_t mb : {8, 32, 64, 96}) {
        size_t bytes = mb << 20;
        for (int mode = 0; mode < 4; mode++) { // 0 cold, 1 bulk prefetch 32KB chunks, 2 line prefetch, 3 warm (read twice)
            k_read<<<1184, 256>>>((const int4 *)flush, big / 16, sink); // evict
            CK(cudaDeviceSynchronize());
            if (mode == 1) k_prefetch_bulk<<<(unsigned)((bytes / 32768 + 255) / 256), 256>>>(buf, bytes, 32768);
            if (mode == 2) k_prefetch_line<<<(unsigned)((bytes / 128 + 255) / 256), 256>>>(buf, bytes);
            if (mode == 3) k_read<<<1184, 256>>>((const int4 *)buf, bytes / 16, sink);
            CK(cudaDeviceSynchronize());
            cudaEventRecord(e0);
            k_read<<<1184, 256>>>((const int4 *)buf, bytes / 16, sink);
            cudaEventRecord(e1);
            CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const char *names[4] = {"cold", "bulk-prefetch", "line-prefetch", "warm(read before)"};
            printf("%3zu MB %-18s %8.1f us  %7.0f GB/s\n", mb, names[mode], ms * 1e3, bytes / ms / 1e6);
        }
    }
    return 0;
}
