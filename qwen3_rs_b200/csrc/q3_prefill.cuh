// q3_prefill.cuh -- batched (T-token) group-scaled int8 GEMM on the 5th-gen tensor cores.
//
//   out[t, r] = sum_g ( (f32)(sum_{k in g} xq[t,k] * wq[r,k]) * ws[r,g] ) * xs[t,g]        (tensor.rs:41-61, T rows at once)
//
// The int32 accumulator is only meaningful within one quantisation group, so the K loop is cut at
// every group: tcgen05.mma.kind::i8 (A, B int8 from shared memory via TMA, 128B swizzle; D int32
// in TMEM) accumulates GS/32 K-steps into one of four TMEM accumulator buffers, commits, and
// moves on to the next buffer while the epilogue warps drain the previous one
// (tcgen05.ld -> cvt -> (dot*ws)*xs -> f32 add).  Each epilogue thread owns one token row and
// adds its groups in order g = 0..ng-1 with unfused multiplies, i.e. exactly the reference's
// left fold: the GEMM result is bit-identical to `matmul` applied token by token.
//
// Warp roles (576 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..17 = epilogue (TMEM lane quadrant = warp % 4, column quarter = (warp - 2) / 4): 32 accumulator
// columns per thread keep the register count low enough for 4 epilogue warps per scheduler, which the
// latency-bound drain (tcgen05.ld -> cvt -> 2 mul -> add per element) needs to fill the issue slots.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "q3_mega.cuh" // mbarrier helpers

namespace q3 {

constexpr int PF_BM = 128, PF_BN = 128, PF_BK = 128; // tile: tokens x weight rows x K bytes per stage
constexpr int PF_STAGES = 4;
constexpr int PF_NACC = 4; // TMEM accumulator buffers (128 columns each)
constexpr int PF_EPI_WARPS = 16;                   // 4 TMEM lane quadrants x 4 column quarters
constexpr int PF_THREADS = (2 + PF_EPI_WARPS) * 32;
constexpr int PF_COLS = PF_BN / (PF_EPI_WARPS / 4); // accumulator columns per epilogue thread (32)
constexpr int PF_SMEM = PF_STAGES * (PF_BM * PF_BK + PF_BN * PF_BK) + 8 * 1024 /* scale-row ring */ + 1024 + 512;

enum { PF_EPI_STORE = 0, PF_EPI_QKV = 1, PF_EPI_RESID = 2, PF_EPI_SWIGLU = 3 };

struct PrefillGemmArgs {
    const float *wsT;  // [K/GS][N]  weight scales, group-major (transposed at load)
    const float *xsT;  // [K/GS][Tpad] activation scales, group-major
    int T, Tpad, N, K;
    float *out;        // STORE: [T][N]; RESID: x[T][N] += ; SWIGLU: hb[T][N/2]
    int ld_out;
    // QKV epilogue
    float *q;          // [T][AH]
    float *kc, *vc;    // layer base of the caches [seq][KV]
    int AH, KV, pos0;
};

__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t saddr) {
    // K-major, SWIZZLE_128B: 8-row atoms of 128 B, atoms 1024 B apart (SBO), LBO unused (=1), version 1
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ULL << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ULL << 46) | (2ULL << 61);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait_spin(uint64_t *bar, uint32_t parity) {
    long long t0 = clock64();
    unsigned spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 1023u) == 0 && clock64() - t0 > 4000000000LL) __trap(); // never hang the GPU
    }
}

// EXACT: every output is the reference's left fold of unfused (dot * ws) * xs terms -- bit-identical to `matmul` (the
// operator-level proof).  !EXACT (what q3_prefill runs): the same exact int32 group dots, drained with
// acc = fma(f32(dot) * ws, xs, acc) in packed f32x2 form, the scale rows of a group (128 weight scales + 128 token scales)
// brought into a small shared-memory ring by the TMA warp instead of 32 global loads per thread and group, and the TMEM
// read of the second half of a thread's columns in flight while the first half is being scaled.
constexpr int PF_SRING = 8; // scale-row slots (one quantisation group each: 512 B of weight scales + 512 B of token scales)

// MODE: 0 = fast drain, 1 = EXACT, 2 = dense ceiling (timing experiment only: the group structure is ignored, all of K is
// accumulated in ONE TMEM buffer and drained once -- what this tiling / pipeline reaches as a plain int8 GEMM; the output
// is the raw integer dot converted to f32)
template <int GS, int EPI, int MODE>
__global__ void __launch_bounds__(PF_THREADS, 1)
    k_gemm_q8(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const PrefillGemmArgs a) {
    extern __shared__ __align__(1024) uint8_t pf_smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(pf_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;                                  // [STAGES][128 rows][128 B]
    uint8_t *sB = smem + PF_STAGES * PF_BM * PF_BK;      // [STAGES][128 rows][128 B]
    float *sscale = reinterpret_cast<float *>(smem + PF_STAGES * (PF_BM + PF_BN) * PF_BK); // [PF_SRING][ws 128 | xs 128]
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + PF_STAGES * (PF_BM + PF_BN) * PF_BK + PF_SRING * 1024);
    uint64_t *empty = full + PF_STAGES;
    uint64_t *tfull = empty + PF_STAGES;
    uint64_t *tempty = tfull + PF_NACC;
    uint64_t *sfull = tempty + PF_NACC;
    uint64_t *sempty = sfull + PF_SRING;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(sempty + PF_SRING);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * PF_BN, m0 = blockIdx.y * PF_BM;
    const int nkb = a.K / PF_BK;
    constexpr int GPS = PF_BK / GS; // groups per stage
    constexpr int KPG = GS / 32;    // MMA K-steps per group
    const int ng = a.K / GS;
    constexpr bool EXACT = MODE == 1, DENSE = MODE == 2;

    if (threadIdx.x == 0) {
        for (int s = 0; s < PF_STAGES; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < PF_NACC; b++) {
            mbar_init(&tfull[b], 1);
            mbar_init(&tempty[b], PF_EPI_WARPS); // one arrive per epilogue warp
        }
        for (int b = 0; b < PF_SRING; b++) {
            mbar_init(&sfull[b], 1);
            mbar_init(&sempty[b], PF_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    }
    if (warp == 1) { // TMEM: all 512 columns = 4 accumulator buffers of 128 x 128 int32
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------- TMA producer -------------------------------
        if (lane == 0) {
            uint64_t policy;
            asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
            int gi = 0;
            for (int kb = 0; kb < nkb; kb++) {
                const int s = kb % PF_STAGES;
                mbar_wait_spin(&empty[s], ((kb / PF_STAGES) & 1) ^ 1);
                mbar_expect_tx(&full[s], (PF_BM + PF_BN) * PF_BK);
                tma_load_2d(sA + (size_t)s * PF_BM * PF_BK, &map_x, kb * PF_BK, m0, &full[s]);
                tma_load_2d(sB + (size_t)s * PF_BN * PF_BK, &map_w, kb * PF_BK, n0, &full[s]);
                if (MODE == 0) { // the scale rows of this stage's groups
                    for (int gg = 0; gg < GPS; gg++, gi++) {
                        const int sl = gi % PF_SRING;
                        mbar_wait_spin(&sempty[sl], ((gi / PF_SRING) & 1) ^ 1);
                        mbar_expect_tx(&sfull[sl], 1024);
                        bulk_g2s(sscale + sl * 256, a.wsT + (size_t)gi * a.N + n0, 512, &sfull[sl], policy);
                        bulk_g2s(sscale + sl * 256 + 128, a.xsT + (size_t)gi * a.Tpad + m0, 512, &sfull[sl], policy);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------- MMA issuer -------------------------------
        if (lane == 0) {
            // instruction descriptor: D = S32, A = B = signed int8, both K-major, N = 128, M = 128
            constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((PF_BN >> 3) << 17) | ((PF_BM >> 4) << 24);
            int gi = 0;
            for (int kb = 0; kb < nkb; kb++) {
                const int s = kb % PF_STAGES;
                mbar_wait_spin(&full[s], (kb / PF_STAGES) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_s = smem_u32(sA + (size_t)s * PF_BM * PF_BK);
                const uint32_t b_s = smem_u32(sB + (size_t)s * PF_BN * PF_BK);
#pragma unroll
                for (int gg = 0; gg < GPS; gg++, gi++) {
                    const int buf = DENSE ? 0 : gi % PF_NACC;
                    if (!DENSE) {
                        mbar_wait_spin(&tempty[buf], ((gi / PF_NACC) & 1) ^ 1);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    }
#pragma unroll
                    for (int kk = 0; kk < KPG; kk++) {
                        const uint32_t koff = gg * GS + kk * 32; // bytes along K inside the 128 B swizzle span
                        umma_i8(tmem_base + buf * PF_BN, umma_desc_k_sw128(a_s + koff), umma_desc_k_sw128(b_s + koff), IDESC, kk > 0 || (DENSE && gi > 0));
                    }
                    if (!DENSE || gi == ng - 1) umma_commit(&tfull[buf]); // accumulator of group gi complete -> epilogue
                }
                umma_commit(&empty[s]); // all MMAs reading this stage retired -> TMA may refill it
            }
        }
    } else {
        // ------------------------------- epilogue -------------------------------
        const int quad = warp & 3;          // TMEM lanes [32*quad, 32*quad+32) are the only ones this warp may read
        const int part = (warp - 2) >> 2;   // which PF_COLS accumulator columns
        const int m = quad * 32 + lane;     // row of the tile = token
        const int t = m0 + m;
        float acc[PF_COLS];
#pragma unroll
        for (int j = 0; j < PF_COLS; j++) acc[j] = 0.0f;
        const float *ws_col = a.wsT + n0 + part * PF_COLS;
        for (int gi = DENSE ? ng - 1 : 0; gi < ng; gi++) {
            const int buf = DENSE ? 0 : gi % PF_NACC;
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * PF_BN + part * PF_COLS;
            if (EXACT || DENSE) {
                const float xs = DENSE ? 1.0f : a.xsT[(size_t)gi * a.Tpad + t];
                mbar_wait_spin(&tfull[buf], DENSE ? 0 : (gi / PF_NACC) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint32_t d[PF_COLS];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                    : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]), "=r"(d[8]), "=r"(d[9]),
                      "=r"(d[10]), "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15]), "=r"(d[16]), "=r"(d[17]), "=r"(d[18]),
                      "=r"(d[19]), "=r"(d[20]), "=r"(d[21]), "=r"(d[22]), "=r"(d[23]), "=r"(d[24]), "=r"(d[25]), "=r"(d[26]), "=r"(d[27]),
                      "=r"(d[28]), "=r"(d[29]), "=r"(d[30]), "=r"(d[31])
                    : "r"(taddr)
                    : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[buf]); // buffer drained into registers: MMA may reuse it
                const float4 *wsg = reinterpret_cast<const float4 *>(ws_col + (size_t)gi * a.N);
#pragma unroll
                for (int j4 = 0; j4 < PF_COLS / 4; j4++) {
                    const float4 w = DENSE ? make_float4(1.f, 1.f, 1.f, 1.f) : __ldg(wsg + j4);
                    // (dot as f32 * weight_scale) * input_scale, then the left-fold add (tensor.rs:59-61)
                    acc[4 * j4 + 0] = __fadd_rn(acc[4 * j4 + 0], __fmul_rn(__fmul_rn((float)(int)d[4 * j4 + 0], w.x), xs));
                    acc[4 * j4 + 1] = __fadd_rn(acc[4 * j4 + 1], __fmul_rn(__fmul_rn((float)(int)d[4 * j4 + 1], w.y), xs));
                    acc[4 * j4 + 2] = __fadd_rn(acc[4 * j4 + 2], __fmul_rn(__fmul_rn((float)(int)d[4 * j4 + 2], w.z), xs));
                    acc[4 * j4 + 3] = __fadd_rn(acc[4 * j4 + 3], __fmul_rn(__fmul_rn((float)(int)d[4 * j4 + 3], w.w), xs));
                }
            } else {
                const int sl = gi % PF_SRING;
                const float *sw = sscale + sl * 256 + part * PF_COLS;
                mbar_wait_spin(&tfull[buf], (gi / PF_NACC) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint32_t d[PF_COLS];
                // first half of this thread's columns, then the second half in flight while the first is being scaled
#define PF_LD16(off, base)                                                                                                             \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"              \
                 : "=r"(d[base + 0]), "=r"(d[base + 1]), "=r"(d[base + 2]), "=r"(d[base + 3]), "=r"(d[base + 4]), "=r"(d[base + 5]),     \
                   "=r"(d[base + 6]), "=r"(d[base + 7]), "=r"(d[base + 8]), "=r"(d[base + 9]), "=r"(d[base + 10]), "=r"(d[base + 11]),   \
                   "=r"(d[base + 12]), "=r"(d[base + 13]), "=r"(d[base + 14]), "=r"(d[base + 15])                                       \
                 : "r"(taddr + off)                                                                                                    \
                 : "memory")
                PF_LD16(0, 0);
                mbar_wait_spin(&sfull[sl], (gi / PF_SRING) & 1); // scale rows of this group landed
                const float xs1 = sscale[sl * 256 + 128 + m];
                const float2 xs = make_float2(xs1, xs1);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                PF_LD16(16, 16);
#pragma unroll
                for (int j4 = 0; j4 < 4; j4++) {
                    const float4 w = *reinterpret_cast<const float4 *>(sw + 4 * j4);
                    float2 *ac = reinterpret_cast<float2 *>(acc + 4 * j4);
                    ac[0] = __ffma2_rn(__fmul2_rn(make_float2((float)(int)d[4 * j4 + 0], (float)(int)d[4 * j4 + 1]), make_float2(w.x, w.y)), xs, ac[0]);
                    ac[1] = __ffma2_rn(__fmul2_rn(make_float2((float)(int)d[4 * j4 + 2], (float)(int)d[4 * j4 + 3]), make_float2(w.z, w.w)), xs, ac[1]);
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[buf]); // buffer drained into registers: MMA may reuse it
#pragma unroll
                for (int j4 = 4; j4 < 8; j4++) {
                    const float4 w = *reinterpret_cast<const float4 *>(sw + 4 * j4);
                    float2 *ac = reinterpret_cast<float2 *>(acc + 4 * j4);
                    ac[0] = __ffma2_rn(__fmul2_rn(make_float2((float)(int)d[4 * j4 + 0], (float)(int)d[4 * j4 + 1]), make_float2(w.x, w.y)), xs, ac[0]);
                    ac[1] = __ffma2_rn(__fmul2_rn(make_float2((float)(int)d[4 * j4 + 2], (float)(int)d[4 * j4 + 3]), make_float2(w.z, w.w)), xs, ac[1]);
                }
#undef PF_LD16
                __syncwarp();
                if (lane == 0) mbar_arrive(&sempty[sl]); // scale rows consumed
            }
        }
        // ---- write the PF_COLS outputs of this token row ----
        if (t < a.T) {
            const int c0 = n0 + part * PF_COLS;
            if (EPI == PF_EPI_STORE) {
                float4 *dst = reinterpret_cast<float4 *>(a.out + (size_t)t * a.ld_out + c0);
#pragma unroll
                for (int j4 = 0; j4 < PF_COLS / 4; j4++) dst[j4] = make_float4(acc[4 * j4], acc[4 * j4 + 1], acc[4 * j4 + 2], acc[4 * j4 + 3]);
            } else if (EPI == PF_EPI_RESID) { // x += (layers.rs:249-259)
                float4 *dst = reinterpret_cast<float4 *>(a.out + (size_t)t * a.ld_out + c0);
#pragma unroll
                for (int j4 = 0; j4 < PF_COLS / 4; j4++) {
                    float4 o = dst[j4];
                    o.x = __fadd_rn(o.x, acc[4 * j4]);
                    o.y = __fadd_rn(o.y, acc[4 * j4 + 1]);
                    o.z = __fadd_rn(o.z, acc[4 * j4 + 2]);
                    o.w = __fadd_rn(o.w, acc[4 * j4 + 3]);
                    dst[j4] = o;
                }
            } else if (EPI == PF_EPI_SWIGLU) { // rows interleaved (gate_j, up_j): layers.rs:472-475
                float2 *dst = reinterpret_cast<float2 *>(a.out + (size_t)t * a.ld_out + c0 / 2);
#pragma unroll
                for (int j = 0; j < PF_COLS / 2; j += 2) {
                    float g0 = acc[2 * j], u0 = acc[2 * j + 1], g1 = acc[2 * j + 2], u1 = acc[2 * j + 3];
                    float s0 = __fmul_rn(g0, __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-g0))));
                    float s1 = __fmul_rn(g1, __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-g1))));
                    dst[j / 2] = make_float2(__fmul_rn(s0, u0), __fmul_rn(s1, u1));
                }
            } else { // PF_EPI_QKV: rows [0,AH) -> q[t], [AH,AH+KV) -> K cache row pos0+t, then V (layers.rs:334-336)
                float *dst;
                if (c0 < a.AH) dst = a.q + (size_t)t * a.AH + c0;
                else if (c0 < a.AH + a.KV) dst = a.kc + (size_t)(a.pos0 + t) * a.KV + (c0 - a.AH);
                else dst = a.vc + (size_t)(a.pos0 + t) * a.KV + (c0 - a.AH - a.KV);
                float4 *d4 = reinterpret_cast<float4 *>(dst);
#pragma unroll
                for (int j4 = 0; j4 < PF_COLS / 4; j4++) d4[j4] = make_float4(acc[4 * j4], acc[4 * j4 + 1], acc[4 * j4 + 2], acc[4 * j4 + 3]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
}

// ------------------------------------------------------------------------------------------
// batched helpers around the GEMM
// ------------------------------------------------------------------------------------------
// per-token RMSNorm (+ embedding gather) + group quantise; scales written group-major [ng][Tpad].
// grid = T, block = 256.  (layers.rs:72-76, 109-130; tensor.rs:91-119)
template <int GS>
__global__ void __launch_bounds__(256) k_pf_norm_quant(float *x, const float *w, int8_t *q, float *sT, int n, int Tpad,
                                                       const int8_t *embed_q, const float *embed_s, const int *tokens, int write_normed) {
    __shared__ float red[8];
    __shared__ float s_f;
    const int t = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float *xr = x + (size_t)t * n;
    const int n4 = n >> 2;
    constexpr int MAXV = 16; // n <= 16384
    float4 v[MAXV];
    float ss = 0.0f;
#pragma unroll
    for (int k = 0; k < MAXV; k++) {
        int i4 = tid + k * 256;
        v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i4 < n4) {
            if (embed_q) {
                size_t base = (size_t)tokens[t] * n + (size_t)i4 * 4;
                char4 e = *reinterpret_cast<const char4 *>(embed_q + base);
                float sc = embed_s[base / GS];
                v[k] = make_float4((float)e.x * sc, (float)e.y * sc, (float)e.z * sc, (float)e.w * sc);
                reinterpret_cast<float4 *>(xr)[i4] = v[k];
            } else {
                v[k] = reinterpret_cast<const float4 *>(xr)[i4];
            }
            ss += __fmul_rn(v[k].x, v[k].x);
            ss += __fmul_rn(v[k].y, v[k].y);
            ss += __fmul_rn(v[k].z, v[k].z);
            ss += __fmul_rn(v[k].w, v[k].w);
        }
    }
    ss = warp_sum(ss);
    if (lane == 0) red[warp] = ss;
    __syncthreads();
    if (tid == 0) {
        float tsum = 0.0f;
        for (int i = 0; i < 8; i++) tsum += red[i];
        s_f = __fdiv_rn(1.0f, sqrtf(__fadd_rn(__fdiv_rn(tsum, (float)n), NORM_EPS)));
    }
    __syncthreads();
    const float f = s_f;
#pragma unroll
    for (int k = 0; k < MAXV; k++) {
        int i4 = tid + k * 256;
        if (i4 - lane < n4) {
            float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i4 < n4) {
                float4 wv = __ldg(reinterpret_cast<const float4 *>(w) + i4);
                y.x = __fmul_rn(wv.x, __fmul_rn(f, v[k].x));
                y.y = __fmul_rn(wv.y, __fmul_rn(f, v[k].y));
                y.z = __fmul_rn(wv.z, __fmul_rn(f, v[k].z));
                y.w = __fmul_rn(wv.w, __fmul_rn(f, v[k].w));
            }
            uint32_t packed;
            float scale;
            quantize_group4<GS>(y, packed, scale);
            if (i4 < n4) {
                reinterpret_cast<uint32_t *>(q + (size_t)t * n)[i4] = packed;
                if ((i4 % (GS / 4)) == 0) sT[(size_t)(i4 / (GS / 4)) * Tpad + t] = scale;
                if (write_normed) reinterpret_cast<float4 *>(xr)[i4] = y;
            }
        }
    }
}

// per-token group quantise of [T][n] f32; scales group-major.  grid = T.
template <int GS>
__global__ void __launch_bounds__(256) k_pf_quantize(const float *__restrict__ x, int8_t *q, float *sT, int n, int Tpad) {
    const int t = blockIdx.x, lane = threadIdx.x & 31;
    const int n4 = n >> 2;
    const float *xr = x + (size_t)t * n;
    for (int base = threadIdx.x - lane; base < n4; base += 256) {
        int i4 = base + lane;
        float4 y = (i4 < n4) ? reinterpret_cast<const float4 *>(xr)[i4] : make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t packed;
        float scale;
        quantize_group4<GS>(y, packed, scale);
        if (i4 < n4) {
            reinterpret_cast<uint32_t *>(q + (size_t)t * n)[i4] = packed;
            if ((i4 % (GS / 4)) == 0) sT[(size_t)(i4 / (GS / 4)) * Tpad + t] = scale;
        }
    }
}

// QK-norm + RoPE for T tokens: grid = (ceil((n_heads + n_kv)/4), T), block 128 (one warp per head).
__global__ void __launch_bounds__(128) k_pf_qknorm_rope(float *q, float *kc_layer, const float *q_ln, const float *k_ln,
                                                        const float *rope, int pos0, int n_heads, int n_kv, int AH, int KV) {
    const int lane = threadIdx.x & 31;
    const int head = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int t = blockIdx.y;
    if (head >= n_heads + n_kv) return;
    const int pos = pos0 + t;
    float *p = head < n_heads ? q + (size_t)t * AH + (size_t)head * HEAD_DIM
                              : kc_layer + (size_t)pos * KV + (size_t)(head - n_heads) * HEAD_DIM;
    float4 v = reinterpret_cast<float4 *>(p)[lane];
    float4 r = qk_norm_rope(v, head < n_heads ? q_ln : k_ln, rope + (size_t)pos * HEAD_DIM, lane);
    reinterpret_cast<float4 *>(p)[lane] = r;
}

// Causal attention for T query tokens over the cache (key positions 0..pos0+t for query t), f32 on the
// CUDA cores, flash-attention tiling (layers.rs:374-419 with an online softmax).  One CTA = one kv head
// x a tile of BQ = 64/KVMUL query tokens, i.e. 64 query rows (token, head-in-group) that all share the
// same K/V stream; key tiles of 32 positions go through shared memory.  256 threads: thread (ty, tx)
// owns a 4-row x 2-key micro-tile of the score tile and a 4-row x 8-dim micro-tile of the output.
constexpr int PFA_R = 64, PFA_BK = 32, PFA_LD = HEAD_DIM + 4, PFA_LDP = PFA_BK + 4;
constexpr int PFA_SMEM = (PFA_R * PFA_LD + 2 * PFA_BK * PFA_LD + PFA_R * PFA_LDP) * 4;

template <int KVMUL>
__global__ void __launch_bounds__(256, 2) k_pf_attention(const float *__restrict__ q, const float *__restrict__ kc,
                                                         const float *__restrict__ vc, float *out, int T, int pos0, int AH, int KV) {
    extern __shared__ __align__(16) float pfa_smem[];
    float *Qs = pfa_smem;                 // [64][132]
    float *Ks = Qs + PFA_R * PFA_LD;      // [32][132]
    float *Vs = Ks + PFA_BK * PFA_LD;     // [32][132]
    float *Ps = Vs + PFA_BK * PFA_LD;     // [64][36]
    constexpr int BQ = PFA_R / KVMUL;
    const int kvh = blockIdx.x, q0 = blockIdx.y * BQ;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const float scale = __fdiv_rn(1.0f, sqrtf((float)HEAD_DIM));
    // Q tile: row r <-> (token q0 + r / KVMUL, head kvh*KVMUL + r % KVMUL)
    for (int i = tid; i < PFA_R * 32; i += 256) {
        const int r = i >> 5, c4 = i & 31;
        const int tq = q0 + r / KVMUL;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tq < T) v = reinterpret_cast<const float4 *>(q + (size_t)tq * AH + (size_t)(kvh * KVMUL + r % KVMUL) * HEAD_DIM)[c4];
        *reinterpret_cast<float4 *>(Qs + r * PFA_LD + c4 * 4) = v;
    }
    float m[4], l[4], acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        m[i] = -INFINITY;
        l[i] = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = 0.0f;
    }
    int last_q = q0 + BQ - 1;
    if (last_q > T - 1) last_q = T - 1;
    const int nkeys = pos0 + last_q + 1; // causal horizon of the last query in the tile
    const float *kb = kc + (size_t)kvh * HEAD_DIM, *vb = vc + (size_t)kvh * HEAD_DIM;
    for (int k0 = 0; k0 < nkeys; k0 += PFA_BK) {
        __syncthreads(); // previous tile fully consumed (also orders the Q tile stores on the first pass)
        for (int i = tid; i < PFA_BK * 32; i += 256) {
            const int p = i >> 5, c4 = i & 31;
            float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
            if (k0 + p < nkeys) {
                kv = reinterpret_cast<const float4 *>(kb + (size_t)(k0 + p) * KV)[c4];
                vv = reinterpret_cast<const float4 *>(vb + (size_t)(k0 + p) * KV)[c4];
            }
            *reinterpret_cast<float4 *>(Ks + p * PFA_LD + c4 * 4) = kv;
            *reinterpret_cast<float4 *>(Vs + p * PFA_LD + c4 * 4) = vv;
        }
        __syncthreads();
        // scores: rows 4*ty..+3, keys tx and tx+16 (adjacent lanes read adjacent rows: with the 132-float
        // row pitch that is the conflict-free pattern for 128-bit shared loads)
        float sc[4][2];
#pragma unroll
        for (int i = 0; i < 4; i++) sc[i][0] = sc[i][1] = 0.0f;
#pragma unroll 8
        for (int d4 = 0; d4 < 32; d4++) {
            float4 kq[2], qq[4];
#pragma unroll
            for (int j = 0; j < 2; j++) kq[j] = *reinterpret_cast<const float4 *>(Ks + (tx + 16 * j) * PFA_LD + d4 * 4);
#pragma unroll
            for (int i = 0; i < 4; i++) qq[i] = *reinterpret_cast<const float4 *>(Qs + (4 * ty + i) * PFA_LD + d4 * 4);
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 2; j++)
                    sc[i][j] += qq[i].x * kq[j].x + qq[i].y * kq[j].y + qq[i].z * kq[j].z + qq[i].w * kq[j].w;
        }
        // online softmax per row (a row's 32 scores live in the 16 lanes that share ty)
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int r = 4 * ty + i;
            const int qpos = pos0 + q0 + r / KVMUL;
            float s0 = (k0 + tx <= qpos) ? __fmul_rn(sc[i][0], scale) : -INFINITY;
            float s1 = (k0 + tx + 16 <= qpos) ? __fmul_rn(sc[i][1], scale) : -INFINITY;
            float mx = fmaxf(s0, s1);
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            const float mn = fmaxf(m[i], mx);
            const float corr = (mn == -INFINITY) ? 1.0f : expf(m[i] - mn);
            const float p0 = (s0 == -INFINITY) ? 0.0f : expf(s0 - mn);
            const float p1 = (s1 == -INFINITY) ? 0.0f : expf(s1 - mn);
            float ps = p0 + p1;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, o);
            l[i] = l[i] * corr + ps;
            m[i] = mn;
#pragma unroll
            for (int j = 0; j < 8; j++) acc[i][j] *= corr;
            Ps[r * PFA_LDP + tx] = p0;
            Ps[r * PFA_LDP + tx + 16] = p1;
        }
        __syncthreads();
        // out[r][8*tx .. +7] += sum_p P[r][p] * V[p][8*tx .. +7]
#pragma unroll 4
        for (int p = 0; p < PFA_BK; p++) {
            const float4 v0 = *reinterpret_cast<const float4 *>(Vs + p * PFA_LD + 8 * tx);
            const float4 v1 = *reinterpret_cast<const float4 *>(Vs + p * PFA_LD + 8 * tx + 4);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const float pr = Ps[(4 * ty + i) * PFA_LDP + p];
                acc[i][0] += pr * v0.x; acc[i][1] += pr * v0.y; acc[i][2] += pr * v0.z; acc[i][3] += pr * v0.w;
                acc[i][4] += pr * v1.x; acc[i][5] += pr * v1.y; acc[i][6] += pr * v1.z; acc[i][7] += pr * v1.w;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int r = 4 * ty + i;
        const int tq = q0 + r / KVMUL;
        if (tq < T) {
            const float inv = __fdiv_rn(1.0f, l[i]);
            float *dst = out + (size_t)tq * AH + (size_t)(kvh * KVMUL + r % KVMUL) * HEAD_DIM + 8 * tx;
            reinterpret_cast<float4 *>(dst)[0] = make_float4(__fmul_rn(acc[i][0], inv), __fmul_rn(acc[i][1], inv), __fmul_rn(acc[i][2], inv), __fmul_rn(acc[i][3], inv));
            reinterpret_cast<float4 *>(dst)[1] = make_float4(__fmul_rn(acc[i][4], inv), __fmul_rn(acc[i][5], inv), __fmul_rn(acc[i][6], inv), __fmul_rn(acc[i][7], inv));
        }
    }
}

// ------------------------------------------------------------------------------------------
// Causal attention on the tensor cores: mma.sync m16n8k8 TF32 with the 3xTF32 split (x = hi + lo, hi = tf32(x),
// lo = tf32(x - hi); a*b ~ hi*hi + hi*lo + lo*hi), which keeps the products at f32-class accuracy (relative error ~2^-21
// instead of TF32's 2^-11: the attention output is re-quantised to int8 right after, and a coarser score would flip far
// more int8 values than the f32 path does).  Same tiling idea as k_pf_attention: one CTA = one kv head x 64 query rows
// (64 / KVMUL tokens x the KVMUL heads that share the K / V stream), 4 warps x 16 rows, key tiles of 32 positions whose
// K and V rows are split into hi / lo once when they are staged in shared memory.
//   S = Q K^T : A = Q fragment (split on the fly), B = K tile (key-major = "col" operand), 4 n-tiles x 16 k-steps
//   O += P V  : the accumulator fragment of S doubles as the A fragment of P when the 8 keys of a k-step are taken in the
//               order (0,2,4,6,1,3,5,7) -- the same permutation is applied to the rows of V in the B fragment, no shuffles.
// ------------------------------------------------------------------------------------------
constexpr int PFT_R = 64, PFT_BK = 32, PFT_LD = HEAD_DIM + 4;
constexpr int PFT_SMEM = (PFT_R * PFT_LD + 4 * PFT_BK * PFT_LD) * 4;

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split_tf32(float x, uint32_t &hi, uint32_t &lo) {
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
    const float r = __fsub_rn(x, __uint_as_float(hi));
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}

template <int KVMUL>
__global__ void __launch_bounds__(128, 2) k_pf_attention_tc(const float *__restrict__ q, const float *__restrict__ kc,
                                                            const float *__restrict__ vc, float *out, int T, int pos0, int AH, int KV) {
    extern __shared__ __align__(16) float pft_smem[];
    float *Qs = pft_smem;                   // [64][132] f32
    float *Kh = Qs + PFT_R * PFT_LD;        // [32][132] tf32 hi
    float *Kl = Kh + PFT_BK * PFT_LD;       // [32][132] tf32 lo
    float *Vh = Kl + PFT_BK * PFT_LD;
    float *Vl = Vh + PFT_BK * PFT_LD;
    constexpr int BQ = PFT_R / KVMUL;
    const int kvh = blockIdx.x, q0 = blockIdx.y * BQ;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const float scale = __fdiv_rn(1.0f, sqrtf((float)HEAD_DIM));
    // Q tile: row r <-> (token q0 + r / KVMUL, head kvh*KVMUL + r % KVMUL)
    for (int i = tid; i < PFT_R * 32; i += 128) {
        const int r = i >> 5, c4 = i & 31;
        const int tq = q0 + r / KVMUL;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tq < T) v = reinterpret_cast<const float4 *>(q + (size_t)tq * AH + (size_t)(kvh * KVMUL + r % KVMUL) * HEAD_DIM)[c4];
        *reinterpret_cast<float4 *>(Qs + r * PFT_LD + c4 * 4) = v;
    }
    const int r0 = warp * 16 + g, r1 = r0 + 8;             // this thread's two rows of the tile
    const int qp0 = pos0 + q0 + r0 / KVMUL, qp1 = pos0 + q0 + r1 / KVMUL; // their absolute positions
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.0f, l1 = 0.0f;
    float o[16][4];
#pragma unroll
    for (int j = 0; j < 16; j++) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.0f;
    int last_q = q0 + BQ - 1;
    if (last_q > T - 1) last_q = T - 1;
    const int nkeys = pos0 + last_q + 1; // causal horizon of the last query in the tile
    const float *kb = kc + (size_t)kvh * HEAD_DIM, *vb = vc + (size_t)kvh * HEAD_DIM;
    for (int k0 = 0; k0 < nkeys; k0 += PFT_BK) {
        __syncthreads(); // previous tile fully consumed (also orders the Q tile stores on the first pass)
        for (int i = tid; i < PFT_BK * 32; i += 128) {
            const int p = i >> 5, c4 = i & 31;
            float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
            if (k0 + p < nkeys) {
                kv = reinterpret_cast<const float4 *>(kb + (size_t)(k0 + p) * KV)[c4];
                vv = reinterpret_cast<const float4 *>(vb + (size_t)(k0 + p) * KV)[c4];
            }
            uint32_t h[4], lo[4];
            split_tf32(kv.x, h[0], lo[0]); split_tf32(kv.y, h[1], lo[1]); split_tf32(kv.z, h[2], lo[2]); split_tf32(kv.w, h[3], lo[3]);
            *reinterpret_cast<uint4 *>(Kh + p * PFT_LD + c4 * 4) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4 *>(Kl + p * PFT_LD + c4 * 4) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            split_tf32(vv.x, h[0], lo[0]); split_tf32(vv.y, h[1], lo[1]); split_tf32(vv.z, h[2], lo[2]); split_tf32(vv.w, h[3], lo[3]);
            *reinterpret_cast<uint4 *>(Vh + p * PFT_LD + c4 * 4) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4 *>(Vl + p * PFT_LD + c4 * 4) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        __syncthreads();
        // ---- S = Q K^T (16 rows x 32 keys per warp) ----
        float s[4][4];
#pragma unroll
        for (int j = 0; j < 4; j++) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.0f;
#pragma unroll 4
        for (int ks = 0; ks < 16; ks++) {
            const int d0 = ks * 8 + t;
            uint32_t ah[4], al[4];
            split_tf32(Qs[r0 * PFT_LD + d0], ah[0], al[0]);
            split_tf32(Qs[r1 * PFT_LD + d0], ah[1], al[1]);
            split_tf32(Qs[r0 * PFT_LD + d0 + 4], ah[2], al[2]);
            split_tf32(Qs[r1 * PFT_LD + d0 + 4], ah[3], al[3]);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int krow = (8 * j + g) * PFT_LD + d0;
                const uint32_t bh0 = __float_as_uint(Kh[krow]), bh1 = __float_as_uint(Kh[krow + 4]);
                const uint32_t bl0 = __float_as_uint(Kl[krow]), bl1 = __float_as_uint(Kl[krow + 4]);
                mma_tf32(s[j], al, bh0, bh1);
                mma_tf32(s[j], ah, bl0, bl1);
                mma_tf32(s[j], ah, bh0, bh1);
            }
        }
        // ---- online softmax: thread holds keys k0 + 8j + {2t, 2t+1} of rows r0 (s[j][0..1]) and r1 (s[j][2..3]) ----
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int key = k0 + 8 * j + 2 * t;
            s[j][0] = (key <= qp0) ? __fmul_rn(s[j][0], scale) : -INFINITY;
            s[j][1] = (key + 1 <= qp0) ? __fmul_rn(s[j][1], scale) : -INFINITY;
            s[j][2] = (key <= qp1) ? __fmul_rn(s[j][2], scale) : -INFINITY;
            s[j][3] = (key + 1 <= qp1) ? __fmul_rn(s[j][3], scale) : -INFINITY;
            mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
            mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
        const float c0 = (mn0 == -INFINITY) ? 1.0f : expf(m0 - mn0), c1 = (mn1 == -INFINITY) ? 1.0f : expf(m1 - mn1);
        m0 = mn0;
        m1 = mn1;
        l0 *= c0;
        l1 *= c1;
#pragma unroll
        for (int j = 0; j < 16; j++) {
            o[j][0] *= c0;
            o[j][1] *= c0;
            o[j][2] *= c1;
            o[j][3] *= c1;
        }
        // ---- O += P V: k-step j = keys 8j .. 8j+7 in the order (0,2,4,6,1,3,5,7) ----
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float p00 = (s[j][0] == -INFINITY) ? 0.0f : expf(s[j][0] - mn0), p01 = (s[j][1] == -INFINITY) ? 0.0f : expf(s[j][1] - mn0);
            const float p10 = (s[j][2] == -INFINITY) ? 0.0f : expf(s[j][2] - mn1), p11 = (s[j][3] == -INFINITY) ? 0.0f : expf(s[j][3] - mn1);
            l0 += p00 + p01;
            l1 += p10 + p11;
            uint32_t ah[4], al[4]; // a0 (r0, k = t) = P[r0][2t], a1 (r1, t) = P[r1][2t], a2 (r0, t+4) = P[r0][2t+1], a3 (r1, t+4) = P[r1][2t+1]
            split_tf32(p00, ah[0], al[0]);
            split_tf32(p10, ah[1], al[1]);
            split_tf32(p01, ah[2], al[2]);
            split_tf32(p11, ah[3], al[3]);
            const int vrow0 = (8 * j + 2 * t) * PFT_LD + g, vrow1 = vrow0 + PFT_LD;
#pragma unroll
            for (int n = 0; n < 16; n++) {
                const uint32_t bh0 = __float_as_uint(Vh[vrow0 + 8 * n]), bh1 = __float_as_uint(Vh[vrow1 + 8 * n]);
                const uint32_t bl0 = __float_as_uint(Vl[vrow0 + 8 * n]), bl1 = __float_as_uint(Vl[vrow1 + 8 * n]);
                mma_tf32(o[n], al, bh0, bh1);
                mma_tf32(o[n], ah, bl0, bl1);
                mma_tf32(o[n], ah, bh0, bh1);
            }
        }
    }
    // row sums live spread over the quad
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = __fdiv_rn(1.0f, l0), i1 = __fdiv_rn(1.0f, l1);
    const int tq0 = q0 + r0 / KVMUL, tq1 = q0 + r1 / KVMUL;
    float *d0 = out + (size_t)tq0 * AH + (size_t)(kvh * KVMUL + r0 % KVMUL) * HEAD_DIM + 2 * t;
    float *d1 = out + (size_t)tq1 * AH + (size_t)(kvh * KVMUL + r1 % KVMUL) * HEAD_DIM + 2 * t;
#pragma unroll
    for (int n = 0; n < 16; n++) {
        if (tq0 < T) *reinterpret_cast<float2 *>(d0 + 8 * n) = make_float2(__fmul_rn(o[n][0], i0), __fmul_rn(o[n][1], i0));
        if (tq1 < T) *reinterpret_cast<float2 *>(d1 + 8 * n) = make_float2(__fmul_rn(o[n][2], i1), __fmul_rn(o[n][3], i1));
    }
}

// [rows][ng] -> [ng][rows] (weight scales, once at load)
__global__ void k_transpose_f32(const float *__restrict__ in, float *out, int rows, int cols) {
    __shared__ float tile[32][33];
    int c = blockIdx.x * 32 + threadIdx.x, r = blockIdx.y * 32 + threadIdx.y;
    for (int j = 0; j < 32; j += 8)
        if (r + j < rows && c < cols) tile[threadIdx.y + j][threadIdx.x] = in[(size_t)(r + j) * cols + c];
    __syncthreads();
    int oc = blockIdx.y * 32 + threadIdx.x, orow = blockIdx.x * 32 + threadIdx.y;
    for (int j = 0; j < 32; j += 8)
        if (orow + j < cols && oc < rows) out[(size_t)(orow + j) * rows + oc] = tile[threadIdx.x][threadIdx.y + j];
}

} // namespace q3
