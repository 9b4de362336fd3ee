// Micro-benchmark: latency of a 148-CTA grid barrier on B200, alone and followed by the
// "every CTA re-reads a 16 KB vector that all CTAs just wrote" pattern of the decode kernel's prologues.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o grid_barrier grid_barrier.cu && ./grid_barrier
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int NT = 512;
constexpr int DIM = 4096;

__device__ __forceinline__ void bar_atomic(unsigned long long *bar, unsigned long long target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(bar), "l"(1ULL) : "memory");
        unsigned long long v;
        do {
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(bar) : "memory");
        } while (v < target);
    }
    __syncthreads();
}

// flags: CTA c publishes flag[c] = epoch; lanes of the first ceil(grid/32) warps each watch one flag
__device__ __forceinline__ void bar_flags(unsigned *flags, unsigned epoch) {
    __syncthreads();
    if (threadIdx.x == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + blockIdx.x), "r"(epoch) : "memory");
    if (threadIdx.x < gridDim.x) {
        unsigned v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + threadIdx.x) : "memory");
        } while (v < epoch);
    }
    __syncthreads();
}

// mode 0: atomic counter; 1: flags; data: 0 none, 1 = write slice + read whole vector + block reduce
__global__ void __launch_bounds__(NT, 1) k_bar(unsigned long long *bar, unsigned *flags, float *vec, int iters, int mode, int data, float *out,
                                                 unsigned long long base) {
    __shared__ float sred[32];
    float acc = 0.f;
    const int per = (DIM + gridDim.x - 1) / gridDim.x;
    for (int it = 0; it < iters; it++) {
        float *v = vec + (it & 1) * DIM;
        if (data) {
            int i = blockIdx.x * per + threadIdx.x;
            if (threadIdx.x < per && i < DIM) v[i] = acc * 1e-9f + (float)it;
        }
        if (mode == 0) bar_atomic(bar, base + (unsigned long long)(it + 1) * gridDim.x);
        else bar_flags(flags, (unsigned)base + it + 1);
        if (data) {
            float4 x[2];
#pragma unroll
            for (int k = 0; k < 2; k++) x[k] = __ldcg(reinterpret_cast<const float4 *>(v) + threadIdx.x + k * NT);
            float ss = 0.f;
#pragma unroll
            for (int k = 0; k < 2; k++) ss += x[k].x * x[k].x + x[k].y * x[k].y + x[k].z * x[k].z + x[k].w * x[k].w;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = ss;
            __syncthreads();
            float t = 0.f;
#pragma unroll
            for (int i = 0; i < NT / 32; i++) t += sred[i];
            acc += t;
            if (data == 2) __syncthreads(); // second CTA barrier as after the quantise step
        }
    }
    if (threadIdx.x == 0) out[blockIdx.x] = acc;
}

int main() {
    int dev = 0, sms = 0;
    CK(cudaSetDevice(dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    unsigned long long *bar;
    unsigned *flags;
    float *vec, *out;
    CK(cudaMalloc(&bar, 8));
    CK(cudaMalloc(&flags, 4 * 1024));
    CK(cudaMalloc(&vec, 2 * DIM * 4));
    CK(cudaMalloc(&out, 4 * 1024));
    CK(cudaMemset(bar, 0, 8));
    CK(cudaMemset(flags, 0, 4 * 1024));
    CK(cudaMemset(vec, 0, 2 * DIM * 4));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int iters = 2000;
    unsigned long long abase = 0, fbase = 0;
    for (int mode = 0; mode < 2; mode++)
        for (int data = 0; data < 3; data++) {
            float best = 1e30f;
            for (int rep = 0; rep < 4; rep++) {
                unsigned long long base = mode == 0 ? abase : fbase;
                void *args[] = {&bar, &flags, &vec, (void *)&iters, &mode, &data, &out, &base};
                CK(cudaEventRecord(e0));
                CK(cudaLaunchCooperativeKernel((void *)k_bar, dim3(sms), dim3(NT), args, 0, 0));
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                float ms;
                CK(cudaEventElapsedTime(&ms, e0, e1));
                if (ms < best) best = ms;
                if (mode == 0) abase += (unsigned long long)iters * sms;
                else fbase += iters;
            }
            printf("grid %d x %d threads, %s barrier, %s: %.3f us per iteration\n", sms, NT, mode == 0 ? "atomic-counter" : "flag-array",
                   data == 0 ? "barrier only" : data == 1 ? "+ write slice / read 16 KB / reduce" : "+ write / read / reduce / 2nd bar.sync", best * 1000.f / iters);
        }
    return 0;
}
