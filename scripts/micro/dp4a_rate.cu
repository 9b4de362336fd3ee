// Micro-benchmark: dp4a (IDP.4A) issue rate per SM on B200, and the rate of the decode kernel's
// "row tile out of shared memory x register-resident activation" inner loop with nothing else going on.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o dp4a_rate dp4a_rate.cu && ./dp4a_rate
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int CHAINS>
__global__ void __launch_bounds__(512, 1) k_dp4a(int *out, int iters, int seed) {
    int acc[CHAINS];
    int a = seed + threadIdx.x, b = seed * 3 + 1;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) acc[c] = c;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) acc[c] = __dp4a(a, b + c, acc[c]);
    }
    int s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) s += acc[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CHAINS>
__global__ void __launch_bounds__(512, 1) k_imad(int *out, int iters, int seed) {
    int acc[CHAINS];
    int a = seed + threadIdx.x, b = seed * 3 + 1;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) acc[c] = c;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) acc[c] = acc[c] * a + (b + c);
    }
    int s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) s += acc[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ int4 lds128(uint32_t saddr) {
    int4 r;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(saddr));
    return r;
}
__device__ __forceinline__ int dot16(int4 w, int4 x, int acc) {
    acc = __dp4a(w.x, x.x, acc);
    acc = __dp4a(w.y, x.y, acc);
    acc = __dp4a(w.z, x.z, acc);
    return __dp4a(w.w, x.w, acc);
}

// 16 warps, each sweeps `tiles` row tiles of 4096 int8 + 64 f32 scales laid out as in the decode kernel
// (lane l owns groups l and l+32; chunk p of block b at b*2048 + p*512 + l*16)
__global__ void __launch_bounds__(512, 1) k_tile(float *out, int tiles, int nwarps) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 8 * 4352 / 4; i += blockDim.x) reinterpret_cast<int *>(smem)[i] = i * 2654435761u;
    __syncthreads();
    if (warp >= nwarps) return;
    int4 x[2][4];
    float xs[2];
#pragma unroll
    for (int b = 0; b < 2; b++) {
        xs[b] = 1.0f + lane;
#pragma unroll
        for (int p = 0; p < 4; p++) x[b][p] = make_int4(lane + p, b, 3, 4);
    }
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem) + (warp & 7) * 4352;
    float total = 0.f;
    for (int t = 0; t < tiles; t++) {
        int4 w[2][4];
        float ws[2];
#pragma unroll
        for (int b = 0; b < 2; b++) {
#pragma unroll
            for (int p = 0; p < 4; p++) w[b][p] = lds128(base + b * 2048 + p * 512 + lane * 16);
            float f;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(f) : "r"(base + 4096 + (b * 32 + lane) * 4));
            ws[b] = f;
        }
        float acc = 0.f;
#pragma unroll
        for (int b = 0; b < 2; b++) {
            int d0 = 0, d1 = 0;
#pragma unroll
            for (int p = 0; p < 4; p += 2) {
                d0 = dot16(w[b][p], x[b][p], d0);
                d1 = dot16(w[b][p + 1], x[b][p + 1], d1);
            }
            acc += ((float)(d0 + d1) * ws[b]) * xs[b];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        total += acc;
        x[0][0].x += 1; // keep the loop from being hoisted
    }
    if (lane == 0) out[blockIdx.x * 16 + warp] = total;
}

int main() {
    int sms = 0;
    CK(cudaSetDevice(0));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    int *out;
    CK(cudaMalloc(&out, sms * 512 * 4));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    int clk_khz = 0;
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    const double ghz = clk_khz / 1e6;
    const int iters = 20000;
    auto time = [&](auto launch) {
        float best = 1e30f;
        for (int r = 0; r < 3; r++) {
            CK(cudaEventRecord(e0));
            launch();
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            best = ms < best ? ms : best;
        }
        return best;
    };
    {
        float ms = time([&] { k_dp4a<8><<<sms, 512>>>(out, iters, 1); });
        double per_sm_clk = 512.0 * 8 * iters / (ms * 1e-3 * ghz * 1e9);
        printf("dp4a : %.1f lane-ops/clk/SM (at the nominal %.3f GHz; %d SMs, 16 warps x 8 chains)\n", per_sm_clk, ghz, sms);
    }
    {
        float ms = time([&] { k_imad<8><<<sms, 512>>>(out, iters, 1); });
        double per_sm_clk = 512.0 * 8 * iters / (ms * 1e-3 * ghz * 1e9);
        printf("imad : %.1f lane-ops/clk/SM\n", per_sm_clk);
    }
    CK(cudaFuncSetAttribute(k_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 4352));
    for (int nw : {16, 8, 4}) {
        const int tiles = 20000;
        float ms = time([&] { k_tile<<<sms, 512, 8 * 4352>>>((float *)out, tiles, nw); });
        double clk_per_tile_round = ms * 1e-3 * ghz * 1e9 / tiles;
        double tbps = (double)sms * nw * tiles * 4352 / (ms * 1e-3) / 1e12;
        printf("tile loop, %2d warps/SM: %.0f clk per round of %d tiles (%.1f TB/s of weight bytes chip-wide)\n", nw, clk_per_tile_round, nw, tbps);
    }
    return 0;
}
