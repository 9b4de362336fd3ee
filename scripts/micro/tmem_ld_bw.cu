// Micro-benchmark: how fast can the CUDA cores read tensor memory?  (tcgen05.ld bandwidth per SM on B200)
//
// The group-scaled prefill GEMM (q3_prefill.cuh) has to pull EVERY int32 group accumulator out of TMEM -- 128 x 128 x 4 B =
// 64 KB per tile and quantisation group, against 2 x 128x128x32 int8 MMAs (~130 clk of tensor pipe at gs 64).  If the TMEM
// read path moves B bytes per clock and SM, the kernel can never spend less than 65536 / B clocks per group-tile, whatever
// the epilogue warps do with the numbers.  This program measures B: one CTA per SM, W epilogue-style warps (warp w reads the
// TMEM lane quadrant w % 4, as the hardware requires), back-to-back tcgen05.ld.32x32b.xN with L loads in flight per warp.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tmem_ld_bw tmem_ld_bw.cu && ./tmem_ld_bw
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

#define LD16(d, taddr)                                                                                                     \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
                 : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]), "=r"(d[8]), \
                   "=r"(d[9]), "=r"(d[10]), "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15])                  \
                 : "r"(taddr)                                                                                              \
                 : "memory")
#define LD32(d, taddr)                                                                                                                    \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                                \
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
                 : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]), "=r"(d[8]), "=r"(d[9]),   \
                   "=r"(d[10]), "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15]), "=r"(d[16]), "=r"(d[17]), "=r"(d[18]),     \
                   "=r"(d[19]), "=r"(d[20]), "=r"(d[21]), "=r"(d[22]), "=r"(d[23]), "=r"(d[24]), "=r"(d[25]), "=r"(d[26]), "=r"(d[27]),     \
                   "=r"(d[28]), "=r"(d[29]), "=r"(d[30]), "=r"(d[31])                                                                      \
                 : "r"(taddr)                                                                                                             \
                 : "memory")

// SHAPE 16: .x16 loads, two in flight; SHAPE 32: .x32 loads, one in flight (+ the next issued before the previous is consumed)
template <int SHAPE>
__global__ void __launch_bounds__(1024, 1) k_tmem_ld(int iters, unsigned *out, long long *cycles) {
    __shared__ unsigned tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned base = tmem_slot + ((unsigned)((warp & 3) * 32) << 16);
    unsigned acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    if (SHAPE == 16) {
        unsigned a[16], b[16];
        for (int it = 0; it < iters; it++) {
            const unsigned col = (unsigned)((it * 32 + (warp >> 2) * 64) & 511) & ~31u;
            LD16(a, base + col);
            LD16(b, base + col + 16);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 16; j++) acc ^= a[j] ^ b[j];
        }
    } else {
        unsigned a[32];
        for (int it = 0; it < iters; it++) {
            const unsigned col = (unsigned)((it * 32 + (warp >> 2) * 64) & 511) & ~31u;
            LD32(a, base + col);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; j++) acc ^= a[j];
        }
    }
    const long long t1 = clock64();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + lane;
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_slot), "n"(512) : "memory");
    }
}

int main() {
    int sms = 0;
    CK(cudaSetDevice(0));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    unsigned *out;
    long long *cyc;
    CK(cudaMalloc(&out, (size_t)sms * 1024 * 4));
    CK(cudaMalloc(&cyc, (size_t)sms * 8));
    long long *h = (long long *)malloc(sms * 8);
    const int iters = 4096;
    printf("tcgen05.ld.32x32b bandwidth, %d SMs, %d iterations of 32 columns (4 KB per warp-iteration)\n", sms, iters);
    printf("%-8s %-6s %14s %16s %22s\n", "shape", "warps", "clk/iter/CTA", "B/clk/SM", "clk per 64 KB (1 group-tile)");
    for (int shape : {16, 32})
        for (int warps : {4, 8, 16, 32}) {
            for (int rep = 0; rep < 2; rep++) {
                if (shape == 16) k_tmem_ld<16><<<sms, warps * 32>>>(iters, out, cyc);
                else k_tmem_ld<32><<<sms, warps * 32>>>(iters, out, cyc);
                CK(cudaDeviceSynchronize());
            }
            CK(cudaMemcpy(h, cyc, sms * 8, cudaMemcpyDeviceToHost));
            double mean = 0;
            for (int i = 0; i < sms; i++) mean += (double)h[i];
            mean /= sms;
            const double bytes = (double)iters * warps * 32 * 32 * 4; // per CTA
            printf(".x%-6d %-6d %14.1f %16.1f %22.0f\n", shape, warps, mean / iters, bytes / mean, 65536.0 / (bytes / mean));
        }
    return 0;
}
