// q3_sampler.cuh -- the reference's token sampler on the device (qwen3-inference/src/sampler.rs), SURVEY 8f-1.
//
// Sampler::sample (:116-136): temperature == 0 -> argmax (k_argmax / the persistent kernel's fused argmax); otherwise
//   logits /= temperature; softmax (layers.rs:495-506); coin = random_f32() (xorshift64*, :44-54); multinomial (:62-71)
//   or top-p (:74-110).
// The point of doing it on the device is that only 4 bytes (the token) cross PCIe per step instead of vocab x f32, and that
// the RNG stream and every decision are those of the reference: the float operations whose ORDER decides the outcome are
// kept in the reference's order -- the softmax denominator and the two cumulative walks are left folds evaluated by one
// thread (152 k dependent adds = ~0.3 ms for the Qwen3 vocabulary; everything element-wise is parallel), exp() is the glibc
// restatement (expf_ref), division is IEEE.  Candidates of equal probability are ordered by index (the reference uses an
// unstable sort there, i.e. leaves that order unspecified).
//
// One CTA of 1024 threads; scratch in global memory: p[n] f32, keys[npad] u64.
#pragma once
#include "q3_kernels.cuh"

namespace q3 {

constexpr int SAMPLE_THREADS = 1024;
constexpr int SAMPLE_SMEM_SORT = 4096; // candidate lists up to this size are sorted in shared memory

struct SampleArgs {
    const float *logits;
    int n;
    float temperature, topp;
    unsigned long long *rng_state; // device-resident xorshift64* state (sampler.rs:20)
    float *p;                      // [n] scratch: probabilities
    unsigned long long *keys;      // [next_pow2(n)] scratch: (descending-prob key << 32 | index)
    int *token_out;
    int *token_feedback, *pos_advance, *history, *history_idx; // optional, like k_argmax
};

__device__ __forceinline__ unsigned long long xorshift_next(unsigned long long &s) { // sampler.rs:44-49
    s ^= s >> 12;
    s ^= s << 25;
    s ^= s >> 27;
    return s;
}
__device__ __forceinline__ float xorshift_f32(unsigned long long &s) { // :52-54
    const unsigned r = (unsigned)((xorshift_next(s) * 0x2545F4914F6CDD1DULL) >> 32);
    return __fdiv_rn((float)(r >> 8), 16777216.0f);
}
// ascending sort key: larger probability first (f32::total_cmp order reversed), then smaller index
__device__ __forceinline__ unsigned long long sample_key(float prob, int idx) {
    const unsigned k = (unsigned)total_key(prob) ^ 0x80000000u; // monotone unsigned image
    return ((unsigned long long)(~k) << 32) | (unsigned)idx;
}
__device__ __forceinline__ float key_prob(unsigned long long key) {
    const unsigned k = ~(unsigned)(key >> 32) ^ 0x80000000u;
    int b = (int)k;
    b ^= (int)(((unsigned)(b >> 31)) >> 1); // total_key is an involution
    return __int_as_float(b);
}

__global__ void __launch_bounds__(SAMPLE_THREADS) k_sample(SampleArgs a) {
    __shared__ float red[32];
    __shared__ int scan[SAMPLE_THREADS];
    __shared__ float s_bcast;
    __shared__ int s_n0, s_tok;
    __shared__ unsigned long long skeys[SAMPLE_SMEM_SORT];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = a.n;
    // ---- logits /= temperature (:122-124), max (layers.rs:496) ----
    float mx = -INFINITY;
    for (int i = tid; i < n; i += SAMPLE_THREADS) {
        const float x = __fdiv_rn(a.logits[i], a.temperature);
        a.p[i] = x;
        mx = fmaxf(mx, x);
    }
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = warp_max(red[lane]);
    // ---- exp (layers.rs:499-502) ----
    for (int i = tid; i < n; i += SAMPLE_THREADS) a.p[i] = expf_ref(__fsub_rn(a.p[i], mx));
    __syncthreads();
    // ---- sum: the reference's left fold, one thread; loads run 32 elements ahead of the dependent adds ----
    if (tid == 0) {
        float sum = 0.0f;
        int i = 0;
        const float4 *p4 = reinterpret_cast<const float4 *>(a.p);
        for (; i + 32 <= n; i += 32) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) v[u] = p4[(i >> 2) + u];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                sum = __fadd_rn(sum, v[u].x);
                sum = __fadd_rn(sum, v[u].y);
                sum = __fadd_rn(sum, v[u].z);
                sum = __fadd_rn(sum, v[u].w);
            }
        }
        for (; i < n; i++) sum = __fadd_rn(sum, a.p[i]);
        s_bcast = __fdiv_rn(1.0f, sum); // layers.rs:504
    }
    __syncthreads();
    const float inv = s_bcast;
    for (int i = tid; i < n; i += SAMPLE_THREADS) a.p[i] = __fmul_rn(a.p[i], inv);
    __syncthreads();
    // ---- coin (:129) ----
    unsigned long long rng = *a.rng_state;
    const float coin = xorshift_f32(rng); // every thread computes the same value; thread 0 stores the new state at the end
    int token = 0;
    if (a.topp <= 0.0f || a.topp >= 1.0f) {
        // ---- sample_mult (:62-71): cumulative walk in index order ----
        if (tid == 0) {
            float cdf = 0.0f;
            int hit = n > 0 ? n - 1 : 0;
            for (int i = 0; i < n; i++) {
                cdf = __fadd_rn(cdf, a.p[i]);
                if (coin < cdf) {
                    hit = i;
                    break;
                }
            }
            s_tok = hit;
        }
    } else {
        // ---- sample_topp (:74-110) ----
        const int denom = n - 1 > 1 ? n - 1 : 1;
        const float cutoff = __fdiv_rn(__fsub_rn(1.0f, a.topp), (float)denom);
        // candidates (prob >= cutoff) compacted in index order: contiguous chunk per thread + block scan of the counts
        const int chunk = (n + SAMPLE_THREADS - 1) / SAMPLE_THREADS;
        const int i0 = tid * chunk, i1 = min(n, i0 + chunk);
        int cnt = 0;
        for (int i = i0; i < i1; i++) cnt += a.p[i] >= cutoff;
        scan[tid] = cnt;
        __syncthreads();
        for (int off = 1; off < SAMPLE_THREADS; off <<= 1) { // Hillis-Steele inclusive scan
            const int v = tid >= off ? scan[tid - off] : 0;
            __syncthreads();
            scan[tid] += v;
            __syncthreads();
        }
        const int n0 = scan[SAMPLE_THREADS - 1];
        int npad = 1;
        while (npad < n0) npad <<= 1;
        unsigned long long *keys = npad <= SAMPLE_SMEM_SORT ? skeys : a.keys;
        int w = scan[tid] - cnt;
        for (int i = i0; i < i1; i++) {
            const float pr = a.p[i];
            if (pr >= cutoff) keys[w++] = sample_key(pr, i);
        }
        for (int i = n0 + tid; i < npad; i += SAMPLE_THREADS) keys[i] = ~0ULL;
        __syncthreads();
        // descending by probability (:95): bitonic sort of the composite keys, ascending
        for (int k = 2; k <= npad; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = tid; i < npad; i += SAMPLE_THREADS) {
                    const int ixj = i ^ j;
                    if (ixj > i) {
                        const unsigned long long x = keys[i], y = keys[ixj];
                        const bool up = (i & k) == 0;
                        if ((x > y) == up) {
                            keys[i] = y;
                            keys[ixj] = x;
                        }
                    }
                }
                __syncthreads();
            }
        }
        if (tid == 0) {
            // truncation point (:98-106) and the draw from the truncated list (:108-116): left folds over the sorted list
            float cum = 0.0f;
            int last = n0 > 0 ? n0 - 1 : 0;
            for (int i = 0; i < n0; i++) {
                cum = __fadd_rn(cum, key_prob(keys[i]));
                if (cum > a.topp) {
                    last = i;
                    break;
                }
            }
            const float r = __fmul_rn(coin, cum);
            float cdf = 0.0f;
            int hit = n0 > 0 ? (int)(keys[last] & 0xffffffffu) : 0;
            for (int i = 0; i <= last && i < n0; i++) {
                cdf = __fadd_rn(cdf, key_prob(keys[i]));
                if (r < cdf) {
                    hit = (int)(keys[i] & 0xffffffffu);
                    break;
                }
            }
            s_tok = hit;
        }
    }
    if (tid == 0) {
        token = s_tok;
        *a.rng_state = rng;
        *a.token_out = token;
        if (a.token_feedback) *a.token_feedback = token;
        if (a.pos_advance) *a.pos_advance += 1;
        if (a.history) {
            a.history[*a.history_idx] = token;
            *a.history_idx += 1;
        }
    }
}

// advance the RNG by n draws without sampling (the reference's chat loop samples -- and discards -- once per prompt
// token, generation.rs:116-122; a host loop that prefills in one call keeps the stream aligned with this)
__global__ void k_rng_skip(unsigned long long *state, int n) {
    unsigned long long s = *state;
    for (int i = 0; i < n; i++) xorshift_next(s);
    *state = s;
}

} // namespace q3
