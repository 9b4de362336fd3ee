/*
 * q3_oracle.c -- CPU restatement of qwen3-rs's quantized Qwen3 forward pass.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under qwen3_rs_b200/ (the product) may
 * import, link or execute this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may.  The product path is the
 * CUDA library and fails loudly without it.
 *
 * PARITY STATUS
 *   - forward path (quantize, matmul, RMSNorm, RoPE, attention, SwiGLU, argmax):
 *     PARITY UNPINNED by the reference -- qwen3-inference ships zero tests and
 *     no golden vectors for it, and the Rust toolchain is absent here so the
 *     reference itself cannot be run.  This restatement is cross-checked against
 *     an independent numpy restatement (oracle/np_forward.py), hand-computed
 *     known answers (tests/test_oracle_*.py) and, for the architecture as a whole,
 *     Hugging Face's fp32 Qwen3 on the same weights (tests/test_oracle_vs_hf.py:
 *     agreement to the int8 quantisation noise, ~3 %; one wrong convention: 40-50 %).
 *   - exporter quantizer (quantize_q80, round_half_to_even,
 *     find_optimal_group_size, header constants): PINNED against the reference's
 *     own known-answer tests, qwen3-export/tests/unit/model_exporter_test.rs.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the reference root).  Rust semantics that matter for bit fidelity and how
 * they are kept here (build: gcc -O2 -ffp-contract=off, no fast-math):
 *   - `.sum::<f32>()` / `.sum::<i32>()` are left folds      -> plain for loops
 *   - Rust never contracts a*b+c into an FMA                -> -ffp-contract=off
 *   - f32::round = half away from zero                      -> roundf
 *   - `as i8` saturates, NaN -> 0                           -> sat_i8()
 *   - f32::max ignores a NaN operand                        -> fmaxf
 *   - exp/cos/sin/powf come from the platform libm          -> glibc expf/cosf/sinf/powf
 *   - rayon only splits rows/heads, never a reduction       -> OpenMP parallel for over rows/heads
 */
#define _GNU_SOURCE
#include <fcntl.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* layers.rs:6,9 */
static const float EPSILON = 1e-6f;
static const float ROPE_BASE_FREQ = 1e6f;

/* configuration.rs:8-12 */
#define CHECKPOINT_MAGIC 0x616a6331
#define CHECKPOINT_VERSION 1
#define HEADER_SIZE 256

static __thread char g_err[512];
ORC_API const char *orc_last_error(void) { return g_err; }

ORC_API void orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
ORC_API int orc_get_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Sensitivity knob (tests only): 0 = reference order (default).  1 = same arithmetic but
 * re-associated float sums (8 interleaved partial sums in RMSNorm / softmax / matmul group
 * fold), i.e. the kind of last-ulp difference any parallel implementation has.  Used to
 * measure how far logits move when nothing but summation order changes. */
static int g_perturb = 0;
ORC_API void orc_set_perturb(int mode) { g_perturb = mode; }

static inline float sum8(const float *t, int n) {
    float p[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < n; i++) p[i & 7] += t[i];
    return ((p[0] + p[1]) + (p[2] + p[3])) + ((p[4] + p[5]) + (p[6] + p[7]));
}

/* Rust `f32 as i8`: saturating, NaN -> 0. */
static inline int8_t sat_i8(float v) {
    if (v != v) return 0;
    if (v >= 127.0f) return 127;
    if (v <= -128.0f) return -128;
    return (int8_t)v; /* truncation toward zero, same as Rust */
}

/* ------------------------------------------------------------------------- */
/* tensor.rs                                                                  */
/* ------------------------------------------------------------------------- */

/* tensor.rs:91-119 quantize(): per group wmax = fold(0, max(|x|)), scale = wmax/127,
 * q = round_half_away(x/scale) as i8, q = 0 when scale == 0; s[g] = scale. */
ORC_API void orc_quantize(int8_t *q, float *s, const float *x, int size, int gs) {
    const float Q_MAX = 127.0f;
    int num_groups = size / gs;
    for (int g = 0; g < num_groups; g++) {
        const float *xg = x + (size_t)g * gs;
        float wmax = 0.0f;
        for (int i = 0; i < gs; i++) wmax = fmaxf(wmax, fabsf(xg[i]));
        float scale = wmax / Q_MAX;
        s[g] = scale;
        for (int i = 0; i < gs; i++) {
            float qv = (scale != 0.0f) ? xg[i] / scale : 0.0f;
            q[(size_t)g * gs + i] = sat_i8(roundf(qv));
        }
    }
}

/* tensor.rs:72-80 dequantize(): x[i] = q[i] as f32 * s[i / gs]. */
ORC_API void orc_dequantize(float *x, const int8_t *q, const float *s, size_t size, int gs) {
    for (size_t i = 0; i < size; i++) x[i] = (float)q[i] * s[i / (size_t)gs];
}

/* tensor.rs:32-62 compute_matmul_row(): per group an i32 dot, then
 * (dot as f32 * weight_scale) * input_scale, groups summed left to right. */
static inline float matmul_row(const int8_t *xq, const float *xs, const int8_t *wq, const float *ws,
                               size_t row, int n, int gs) {
    size_t row_off = row * (size_t)n;
    int num_groups = n / gs;
    float acc = 0.0f;
    float part[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int g = 0; g < num_groups; g++) {
        size_t woff = row_off + (size_t)g * gs;
        const int8_t *xp = xq + (size_t)g * gs;
        const int8_t *wp = wq + woff;
        int32_t dot = 0;
        for (int k = 0; k < gs; k++) dot += (int32_t)xp[k] * (int32_t)wp[k];
        float wscale = ws[woff / (size_t)gs];
        float xscale = xs[g];
        float term = (float)dot * wscale * xscale;
        if (g_perturb) part[g & 7] += term;
        else acc += term;
    }
    if (g_perturb) acc = ((part[0] + part[1]) + (part[2] + part[3])) + ((part[4] + part[5]) + (part[6] + part[7]));
    return acc;
}

/* tensor.rs:23-30 matmul(): rows in parallel (rayon par_iter_mut -> omp for), first d outputs. */
ORC_API void orc_matmul(float *xout, const int8_t *xq, const float *xs, const int8_t *wq, const float *ws,
                        int n, int d, int gs) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < d; i++) xout[i] = matmul_row(xq, xs, wq, ws, (size_t)i, n, gs);
}

/* The per-group int32 dot products of tensor.rs:47-51, exposed on their own so the
 * CUDA kernels' integer part can be compared bit-for-bit: dots[row*(n/gs)+g]. */
ORC_API void orc_group_dots(int32_t *dots, const int8_t *xq, const int8_t *wq, int n, int d, int gs) {
    int ng = n / gs;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < d; i++) {
        for (int g = 0; g < ng; g++) {
            const int8_t *xp = xq + (size_t)g * gs;
            const int8_t *wp = wq + (size_t)i * n + (size_t)g * gs;
            int32_t dot = 0;
            for (int k = 0; k < gs; k++) dot += (int32_t)xp[k] * (int32_t)wp[k];
            dots[(size_t)i * ng + g] = dot;
        }
    }
}

/* ------------------------------------------------------------------------- */
/* layers.rs                                                                  */
/* ------------------------------------------------------------------------- */

/* layers.rs:109-119 RMSNorm::forward (and :121-130 forward_inplace, same arithmetic;
 * out may alias in). */
ORC_API void orc_rmsnorm(float *out, const float *in, const float *w, int n) {
    float ss = 0.0f;
    if (g_perturb) {
        float p[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int i = 0; i < n; i++) p[i & 7] += in[i] * in[i];
        ss = ((p[0] + p[1]) + (p[2] + p[3])) + ((p[4] + p[5]) + (p[6] + p[7]));
    } else {
        for (int i = 0; i < n; i++) ss += in[i] * in[i];
    }
    float f = 1.0f / sqrtf((ss / (float)n) + EPSILON);
    for (int i = 0; i < n; i++) out[i] = w[i] * (f * in[i]);
}

/* layers.rs:161-171 RoPE::compute_freqs: freq = 1e6^(-(i as f32)/half), angle = pos as f32 * freq. */
ORC_API void orc_rope_freqs(float *cos_sin /* [half][2] */, int pos, int head_dim) {
    int half = head_dim / 2;
    for (int i = 0; i < half; i++) {
        float freq = powf(ROPE_BASE_FREQ, -((float)i) / (float)half);
        float angle = (float)pos * freq;
        cos_sin[2 * i + 0] = cosf(angle);
        cos_sin[2 * i + 1] = sinf(angle);
    }
}

/* layers.rs:173-185 RoPE::apply: half-split pairs (x_i, x_{i+half}). */
ORC_API void orc_rope_apply(float *v, const float *cos_sin, int head_dim) {
    int half = head_dim / 2;
    for (int i = 0; i < half; i++) {
        float c = cos_sin[2 * i], s = cos_sin[2 * i + 1];
        float x = v[i], y = v[i + half];
        v[i] = x * c - y * s;
        v[i + half] = x * s + y * c;
    }
}

/* The platform expf (what Rust's f32::exp resolves to), exposed for testing the device restatement. */
ORC_API void orc_expf_array(float *out, const float *x, size_t n) {
    for (size_t i = 0; i < n; i++) out[i] = expf(x[i]);
}

/* layers.rs:495-506 softmax(): max fold from -inf, exp(x-max), left-fold sum, multiply by 1/sum. */
ORC_API void orc_softmax(float *x, int n) {
    float m = -INFINITY;
    for (int i = 0; i < n; i++) m = fmaxf(m, x[i]);
    float sum = 0.0f;
    for (int i = 0; i < n; i++) {
        x[i] = expf(x[i] - m);
        sum += x[i];
    }
    if (g_perturb) sum = sum8(x, n);
    float inv = 1.0f / sum;
    for (int i = 0; i < n; i++) x[i] *= inv;
}

/* ------------------------------------------------------------------------- */
/* sampler.rs                                                                 */
/* ------------------------------------------------------------------------- */

/* total_cmp key: sign-magnitude bits -> monotone signed integer (f32::total_cmp). */
static inline int32_t total_key(float f) {
    int32_t b;
    memcpy(&b, &f, 4);
    b ^= (int32_t)(((uint32_t)(b >> 31)) >> 1);
    return b;
}

/* sampler.rs:57-59 sample_argmax: Iterator::max_by(total_cmp) keeps the LAST of equal maxima. */
ORC_API int orc_argmax(const float *logits, int n) {
    int best = 0;
    for (int i = 1; i < n; i++)
        if (total_key(logits[i]) >= total_key(logits[best])) best = i;
    return best;
}

typedef struct {
    float prob;
    int index;
} ProbIndex;

typedef struct {
    ProbIndex *probindex;
    int vocab;
    float temperature, topp;
    uint64_t rng_state;
} OrcSampler;

/* sampler.rs:30-41 Sampler::new */
ORC_API OrcSampler *orc_sampler_new(int vocab, float temperature, float topp, uint64_t seed) {
    OrcSampler *s = (OrcSampler *)calloc(1, sizeof(OrcSampler));
    s->probindex = (ProbIndex *)calloc((size_t)vocab, sizeof(ProbIndex));
    s->vocab = vocab;
    s->temperature = temperature;
    s->topp = topp < 0.0f ? 0.0f : (topp > 1.0f ? 1.0f : topp);
    s->rng_state = seed;
    return s;
}
ORC_API void orc_sampler_free(OrcSampler *s) {
    if (s) {
        free(s->probindex);
        free(s);
    }
}
/* sampler.rs:44-49 random_u32 (xorshift64*) */
ORC_API uint32_t orc_sampler_random_u32(OrcSampler *s) {
    s->rng_state ^= s->rng_state >> 12;
    s->rng_state ^= s->rng_state << 25;
    s->rng_state ^= s->rng_state >> 27;
    return (uint32_t)((s->rng_state * 0x2545F4914F6CDD1DULL) >> 32);
}
/* sampler.rs:52-54 random_f32 */
ORC_API float orc_sampler_random_f32(OrcSampler *s) {
    return (float)(orc_sampler_random_u32(s) >> 8) / 16777216.0f;
}
/* sampler.rs:62-71 sample_mult */
static int sample_mult(const float *p, int n, float coin) {
    float cdf = 0.0f;
    for (int i = 0; i < n; i++) {
        cdf += p[i];
        if (coin < cdf) return i;
    }
    return n > 0 ? n - 1 : 0;
}
static int cmp_prob_desc(const void *a, const void *b) {
    int32_t ka = total_key(((const ProbIndex *)a)->prob), kb = total_key(((const ProbIndex *)b)->prob);
    return (kb > ka) - (kb < ka);
}
/* sampler.rs:74-110 sample_topp (sort_unstable: order among equal probs is unspecified in
 * the reference too). */
static int sample_topp(OrcSampler *s, const float *p, int n, float coin) {
    int denom = n - 1 > 1 ? n - 1 : 1;
    float cutoff = (1.0f - s->topp) / (float)denom;
    int n0 = 0;
    for (int i = 0; i < n; i++)
        if (p[i] >= cutoff) {
            s->probindex[n0].prob = p[i];
            s->probindex[n0].index = i;
            n0++;
        }
    qsort(s->probindex, (size_t)n0, sizeof(ProbIndex), cmp_prob_desc);
    float cum = 0.0f;
    int last = n0 > 0 ? n0 - 1 : 0;
    for (int i = 0; i < n0; i++) {
        cum += s->probindex[i].prob;
        if (cum > s->topp) {
            last = i;
            break;
        }
    }
    float r = coin * cum;
    float cdf = 0.0f;
    for (int i = 0; i <= last; i++) {
        cdf += s->probindex[i].prob;
        if (r < cdf) return s->probindex[i].index;
    }
    return s->probindex[last].index;
}
/* sampler.rs:116-136 sample(): mutates logits in place like the reference. */
ORC_API int orc_sampler_sample(OrcSampler *s, float *logits) {
    int n = s->vocab;
    if (s->temperature == 0.0f) return orc_argmax(logits, n);
    for (int i = 0; i < n; i++) logits[i] /= s->temperature;
    orc_softmax(logits, n);
    float coin = orc_sampler_random_f32(s);
    if (s->topp <= 0.0f || s->topp >= 1.0f) return sample_mult(logits, n, coin);
    return sample_topp(s, logits, n, coin);
}

/* ------------------------------------------------------------------------- */
/* qwen3-export/src/model_exporter.rs (fixture path)                          */
/* ------------------------------------------------------------------------- */

/* model_exporter.rs:321-338 round_half_to_even */
ORC_API float orc_round_half_to_even(float x) {
    float rounded = roundf(x);
    float diff = fabsf(x - rounded);
    if (diff != 0.5f) return rounded;
    if (((int32_t)rounded) % 2 == 0) return rounded;
    return x >= 0.0f ? rounded - 1.0f : rounded + 1.0f;
}

/* model_exporter.rs:48-57 find_optimal_group_size (MIN_GROUP_SIZE = 4, :37) */
ORC_API int orc_find_optimal_group_size(int hidden_dim, int requested) {
    const int MIN_GROUP_SIZE = 4;
    int size = requested < hidden_dim ? requested : hidden_dim;
    while (size >= MIN_GROUP_SIZE && hidden_dim % size != 0) size /= 2;
    return size > MIN_GROUP_SIZE ? size : MIN_GROUP_SIZE;
}

/* model_exporter.rs:104-161 quantize_q80: scale = max|w|/127 or 1.0 for an all-zero group,
 * q = clamp(round_half_to_even(w/scale), -127, 127) as i8.  Returns -1 when len is not a
 * multiple of gs (reference: Err("Weight length is not a multiple of group_size")). */
ORC_API int orc_quantize_q80(int8_t *q, float *s, float *max_error, const float *w, size_t len, int gs) {
    if (gs <= 0 || len % (size_t)gs != 0) {
        snprintf(g_err, sizeof g_err, "Weight length is not a multiple of group_size");
        return -1;
    }
    size_t ng = len / (size_t)gs;
    float maxerr = 0.0f;
#pragma omp parallel for schedule(static) reduction(max : maxerr)
    for (size_t g = 0; g < ng; g++) {
        const float *wg = w + g * (size_t)gs;
        float gmax = 0.0f;
        for (int i = 0; i < gs; i++) gmax = fmaxf(gmax, fabsf(wg[i])); /* fold(0.0, f32::max) */
        float scale = gmax > 0.0f ? gmax / 127.0f : 1.0f;
        s[g] = scale;
        for (int i = 0; i < gs; i++) {
            int8_t qi = 0;
            if (scale > 0.0f) {
                float r = orc_round_half_to_even(wg[i] / scale);
                /* f32::clamp(-127, 127): NaN stays NaN, then `as i8` -> 0 */
                if (r < -127.0f) r = -127.0f;
                if (r > 127.0f) r = 127.0f;
                qi = sat_i8(r);
            }
            q[g * (size_t)gs + i] = qi;
            float err = fabsf((float)qi * scale - wg[i]);
            maxerr = fmaxf(maxerr, err); /* f32::max drops NaN */
        }
    }
    if (max_error) *max_error = maxerr;
    return 0;
}

/* ------------------------------------------------------------------------- */
/* configuration.rs / utils.rs / models/mod.rs / models/qwen3.rs              */
/* ------------------------------------------------------------------------- */

/* POD mirror of ModelConfig (configuration.rs:18-30). */
typedef struct {
    int architecture_id, dim, hidden_dim, n_layers, n_heads, n_kv_heads, head_dim, seq_len, vocab_size,
        group_size, shared_classifier;
} OrcConfig;

typedef struct {
    const int8_t *q;
    const float *s;
} QT; /* QuantizedTensor borrowed from the mmap (tensor.rs:17-20) */

typedef struct {
    OrcConfig cfg;
    /* mmap (utils.rs:7-16) */
    uint8_t *map;
    size_t map_len;
    /* weights (qwen3.rs:199-277) */
    const float *rms_att, *rms_ffn, *rms_final, *q_ln, *k_ln;
    QT embed, *wq, *wk, *wv, *wo, *w1, *w2, *w3, wcls;
    float *token_embedding_table; /* dequantized at load, qwen3.rs:241-242 */
    /* buffers (qwen3.rs:412-445) */
    float *x, *xb, *xb2, *q, *att, *hb, *hb2, *key_cache, *value_cache, *logits;
    int8_t *xq_q, *hq_q;
    float *xq_s, *hq_s;
    /* optional per-layer trace for kernel-level parity tests */
    int trace_layer;
    float *xdump; /* optional [(L+1)][dim]: residual stream entering layer 0 and leaving every layer */
} OrcModel;

/* utils.rs:17-49: sequential cursor over the mapping with bounds checks. */
typedef struct {
    uint8_t *base;
    size_t len, off;
} Cursor;
static const void *cur_take(Cursor *c, size_t bytes, const char *what) {
    if (c->off + bytes > c->len) {
        snprintf(g_err, sizeof g_err, "Failed to read %s: Insufficient data: need %zu bytes, have %zu remaining",
                 what, bytes, c->len - c->off);
        return NULL;
    }
    const void *p = c->base + c->off;
    c->off += bytes;
    return p;
}

/* models/mod.rs:83-110 create_quantized_tensors: i8[size] then f32[size/gs], n times. */
static int take_qts(Cursor *c, QT *out, int n, size_t size_each, int gs, const char *what) {
    for (int i = 0; i < n; i++) {
        out[i].q = (const int8_t *)cur_take(c, size_each, what);
        if (!out[i].q) return -1;
        out[i].s = (const float *)cur_take(c, (size_each / (size_t)gs) * 4, what);
        if (!out[i].s) return -1;
    }
    return 0;
}

ORC_API void orc_model_free(OrcModel *m) {
    if (!m) return;
    free(m->wq); free(m->wk); free(m->wv); free(m->wo); free(m->w1); free(m->w2); free(m->w3);
    free(m->token_embedding_table);
    free(m->x); free(m->xb); free(m->xb2); free(m->q); free(m->att); free(m->hb); free(m->hb2);
    free(m->key_cache); free(m->value_cache); free(m->logits);
    free(m->xq_q); free(m->hq_q); free(m->xq_s); free(m->hq_s);
    if (m->map) munmap(m->map, m->map_len);
    free(m);
}

/* models/mod.rs:55-73 TransformerBuilder::build + configuration.rs:77-146 read_config /
 * validate_config + qwen3.rs:17-52,199-277,412-445.  ctx_len <= 0 means "no override". */
ORC_API OrcModel *orc_model_open(const char *path, int ctx_len) {
    int fd = open(path, O_RDONLY);
    if (fd < 0) {
        snprintf(g_err, sizeof g_err, "Failed to open checkpoint: %s", path);
        return NULL;
    }
    struct stat st;
    fstat(fd, &st);
    OrcModel *m = (OrcModel *)calloc(1, sizeof(OrcModel));
    m->trace_layer = -1;
    m->map_len = (size_t)st.st_size;
    m->map = (uint8_t *)mmap(NULL, m->map_len ? m->map_len : 1, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (m->map == MAP_FAILED) {
        m->map = NULL;
        snprintf(g_err, sizeof g_err, "Failed to create memory mapping");
        orc_model_free(m);
        return NULL;
    }
    Cursor c = {m->map, m->map_len, 0};
    const int32_t *h = (const int32_t *)cur_take(&c, 13 * 4, "config");
    if (!h) goto fail;
    /* configuration.rs:93-107 field order */
    int32_t magic = h[0], version = h[1];
    if (magic != CHECKPOINT_MAGIC) {
        snprintf(g_err, sizeof g_err, "Invalid checkpoint magic number: expected %#x, got %#x", CHECKPOINT_MAGIC, magic);
        goto fail;
    }
    if (version != CHECKPOINT_VERSION) {
        snprintf(g_err, sizeof g_err, "Unsupported checkpoint version: expected %d, got %d", CHECKPOINT_VERSION, version);
        goto fail;
    }
    {
        const char *names[8] = {"architecture_id", "dim", "n_layers", "n_heads", "n_kv_heads", "vocab_size", "seq_len", "head_dim"};
        int32_t vals[8] = {h[2], h[3], h[5], h[6], h[7], h[8], h[9], h[10]};
        for (int i = 0; i < 8; i++)
            if (vals[i] <= 0) {
                snprintf(g_err, sizeof g_err, "Invalid %s: must be positive, got %d", names[i], vals[i]);
                goto fail;
            }
    }
    if (!cur_take(&c, HEADER_SIZE - 13 * 4, "header padding")) goto fail; /* configuration.rs:110 */
    OrcConfig *cf = &m->cfg;
    cf->architecture_id = h[2]; cf->dim = h[3]; cf->hidden_dim = h[4]; cf->n_layers = h[5];
    cf->n_heads = h[6]; cf->n_kv_heads = h[7]; cf->vocab_size = h[8]; cf->seq_len = h[9];
    cf->head_dim = h[10]; cf->shared_classifier = h[11] != 0; cf->group_size = h[12];
    if (ctx_len > 0 && ctx_len < cf->seq_len) cf->seq_len = ctx_len; /* models/mod.rs:65-67 */
    if (cf->architecture_id != 1) {
        snprintf(g_err, sizeof g_err, "Unknown architecture_id: %d", cf->architecture_id);
        goto fail;
    }
    if (cf->group_size <= 0 || cf->hidden_dim <= 0) {
        snprintf(g_err, sizeof g_err, "Invalid group_size/hidden_dim");
        goto fail;
    }
    {
        int L = cf->n_layers, dim = cf->dim, hd = cf->head_dim, gs = cf->group_size;
        size_t AH = (size_t)cf->n_heads * hd, KV = (size_t)cf->n_kv_heads * hd, H = (size_t)cf->hidden_dim;
        size_t V = (size_t)cf->vocab_size;
        /* qwen3.rs:228-232 */
        if (!(m->rms_att = (const float *)cur_take(&c, (size_t)L * dim * 4, "attention normalization weights"))) goto fail;
        if (!(m->rms_ffn = (const float *)cur_take(&c, (size_t)L * dim * 4, "FFN normalization weights"))) goto fail;
        if (!(m->rms_final = (const float *)cur_take(&c, (size_t)dim * 4, "final normalization weights"))) goto fail;
        if (!(m->q_ln = (const float *)cur_take(&c, (size_t)L * hd * 4, "query layer norm weights"))) goto fail;
        if (!(m->k_ln = (const float *)cur_take(&c, (size_t)L * hd * 4, "key layer norm weights"))) goto fail;
        /* qwen3.rs:235-259 */
        if (take_qts(&c, &m->embed, 1, V * dim, gs, "token embedding")) goto fail;
        m->wq = calloc(L, sizeof(QT)); m->wk = calloc(L, sizeof(QT)); m->wv = calloc(L, sizeof(QT));
        m->wo = calloc(L, sizeof(QT)); m->w1 = calloc(L, sizeof(QT)); m->w2 = calloc(L, sizeof(QT));
        m->w3 = calloc(L, sizeof(QT));
        if (take_qts(&c, m->wq, L, (size_t)dim * AH, gs, "wq")) goto fail;
        if (take_qts(&c, m->wk, L, (size_t)dim * KV, gs, "wk")) goto fail;
        if (take_qts(&c, m->wv, L, (size_t)dim * KV, gs, "wv")) goto fail;
        if (take_qts(&c, m->wo, L, AH * dim, gs, "wo")) goto fail;
        if (take_qts(&c, m->w1, L, (size_t)dim * H, gs, "w1")) goto fail;
        if (take_qts(&c, m->w2, L, H * dim, gs, "w2")) goto fail;
        if (take_qts(&c, m->w3, L, (size_t)dim * H, gs, "w3")) goto fail;
        if (cf->shared_classifier) m->wcls = m->embed;
        else if (take_qts(&c, &m->wcls, 1, (size_t)dim * V, gs, "classifier")) goto fail;
        /* qwen3.rs:241-242 */
        m->token_embedding_table = (float *)malloc(V * dim * 4);
        orc_dequantize(m->token_embedding_table, m->embed.q, m->embed.s, V * dim, gs);
        /* qwen3.rs:420-444 (vec![0.0; ..] == calloc) */
        size_t S = (size_t)cf->seq_len;
        m->x = calloc(dim, 4); m->xb = calloc(AH, 4); m->xb2 = calloc(dim, 4);
        m->xq_q = calloc(AH > (size_t)dim ? AH : (size_t)dim, 1);
        m->xq_s = calloc((AH > (size_t)dim ? AH : (size_t)dim) / gs + 1, 4);
        m->q = calloc(AH, 4); m->att = calloc((size_t)cf->n_heads * S, 4);
        m->hb = calloc(H, 4); m->hb2 = calloc(H, 4); m->hq_q = calloc(H, 1); m->hq_s = calloc(H / gs + 1, 4);
        m->key_cache = calloc((size_t)L * S * KV, 4); m->value_cache = calloc((size_t)L * S * KV, 4);
        m->logits = calloc(V, 4);
        if (!m->key_cache || !m->value_cache || !m->att || !m->token_embedding_table) {
            snprintf(g_err, sizeof g_err, "out of memory allocating buffers");
            goto fail;
        }
    }
    return m;
fail:
    orc_model_free(m);
    return NULL;
}

ORC_API const OrcConfig *orc_model_config(const OrcModel *m) { return &m->cfg; }
ORC_API size_t orc_model_file_bytes(const OrcModel *m) { return m->map_len; }

/* layers.rs:346-372 apply_qk_normalization_and_rope */
static void qk_norm_rope(OrcModel *m, int l, size_t cur_off, const float *freqs) {
    const OrcConfig *cf = &m->cfg;
    int hd = cf->head_dim;
    float temp[1024];
    for (int h = 0; h < cf->n_heads; h++) {
        float *qs = m->q + (size_t)h * hd;
        memcpy(temp, qs, (size_t)hd * 4);
        orc_rmsnorm(qs, temp, m->q_ln + (size_t)l * hd, hd);
        orc_rope_apply(qs, freqs, hd);
    }
    for (int h = 0; h < cf->n_kv_heads; h++) {
        float *ks = m->key_cache + cur_off + (size_t)h * hd;
        memcpy(temp, ks, (size_t)hd * 4);
        orc_rmsnorm(ks, temp, m->k_ln + (size_t)l * hd, hd);
        orc_rope_apply(ks, freqs, hd);
    }
}

/* layers.rs:374-419 compute_attention: heads in parallel; per head scores over 0..=pos,
 * softmax, then out += a_t * v_t sequentially over t (mul then add, no FMA). */
static void compute_attention(OrcModel *m, int pos, size_t kv_off) {
    const OrcConfig *cf = &m->cfg;
    int hd = cf->head_dim, kv_mul = cf->n_heads / cf->n_kv_heads;
    size_t kv_dim = (size_t)cf->n_kv_heads * hd;
    float scale = 1.0f / sqrtf((float)hd); /* (head_dim as f32).sqrt().recip() */
#pragma omp parallel for schedule(static)
    for (int h = 0; h < cf->n_heads; h++) {
        float *att = m->att + (size_t)h * cf->seq_len;
        float *xb = m->xb + (size_t)h * hd;
        const float *qh = m->q + (size_t)h * hd;
        int kvh = h / kv_mul;
        for (int t = 0; t <= pos; t++) {
            const float *k = m->key_cache + kv_off + (size_t)t * kv_dim + (size_t)kvh * hd;
            float s = 0.0f;
            for (int i = 0; i < hd; i++) s += qh[i] * k[i];
            att[t] = s * scale;
        }
        orc_softmax(att, pos + 1);
        for (int i = 0; i < hd; i++) xb[i] = 0.0f;
        for (int t = 0; t <= pos; t++) {
            const float *v = m->value_cache + kv_off + (size_t)t * kv_dim + (size_t)kvh * hd;
            float a = att[t];
            for (int i = 0; i < hd; i++) xb[i] += a * v[i];
        }
    }
}

/* Optional trace of one layer's intermediates (kernel-level parity tests). */
typedef struct {
    int8_t *xq_attn_q; float *xq_attn_s;   /* quantize(rmsnorm_att(x))        [dim]  */
    float *q_post;                           /* q after qk-norm + rope          [AH]   */
    float *k_post, *v_row;                   /* cache rows written at pos       [kv]   */
    float *att_out;                          /* attention output xb             [AH]   */
    float *x_after_attn;                     /* x after first residual          [dim]  */
    float *hb_swiglu;                        /* hb after SwiGLU                 [H]    */
    int8_t *hq_q; float *hq_s;               /* quantize(hb)                    [H]    */
    float *x_out;                            /* x after the block               [dim]  */
} OrcTrace;
static OrcTrace g_trace;

ORC_API void orc_model_set_trace(OrcModel *m, int layer, OrcTrace *t) {
    m->trace_layer = layer;
    if (t) g_trace = *t;
}

/* qwen3.rs:131-176 TransformerBlock::forward, with layers.rs:328-344 and :466-480 inlined. */
static void block_forward(OrcModel *m, int l, int pos) {
    const OrcConfig *cf = &m->cfg;
    int dim = cf->dim, hd = cf->head_dim, gs = cf->group_size, H = cf->hidden_dim;
    int AH = cf->n_heads * hd, KV = cf->n_kv_heads * hd;
    int tr = (m->trace_layer == l);
    orc_rmsnorm(m->xb, m->x, m->rms_att + (size_t)l * dim, dim);   /* :134 */
    orc_quantize(m->xq_q, m->xq_s, m->xb, dim, gs);                /* :136 */
    if (tr && g_trace.xq_attn_q) { memcpy(g_trace.xq_attn_q, m->xq_q, dim); memcpy(g_trace.xq_attn_s, m->xq_s, (size_t)(dim / gs) * 4); }
    /* layers.rs:329-336 */
    size_t kv_off = (size_t)l * cf->seq_len * KV;
    size_t cur = kv_off + (size_t)pos * KV;
    orc_matmul(m->q, m->xq_q, m->xq_s, m->wq[l].q, m->wq[l].s, dim, AH, gs);
    orc_matmul(m->key_cache + cur, m->xq_q, m->xq_s, m->wk[l].q, m->wk[l].s, dim, KV, gs);
    orc_matmul(m->value_cache + cur, m->xq_q, m->xq_s, m->wv[l].q, m->wv[l].s, dim, KV, gs);
    float freqs[1024];
    orc_rope_freqs(freqs, pos, hd);                                 /* layers.rs:339 */
    qk_norm_rope(m, l, cur, freqs);                                 /* layers.rs:340 */
    if (tr && g_trace.q_post) { memcpy(g_trace.q_post, m->q, (size_t)AH * 4); memcpy(g_trace.k_post, m->key_cache + cur, (size_t)KV * 4); memcpy(g_trace.v_row, m->value_cache + cur, (size_t)KV * 4); }
    compute_attention(m, pos, kv_off);                              /* layers.rs:343 */
    if (tr && g_trace.att_out) memcpy(g_trace.att_out, m->xb, (size_t)AH * 4);
    orc_quantize(m->xq_q, m->xq_s, m->xb, AH, gs);                 /* :152 */
    orc_matmul(m->xb2, m->xq_q, m->xq_s, m->wo[l].q, m->wo[l].s, AH, dim, gs); /* :153 */
    for (int i = 0; i < dim; i++) m->x[i] += m->xb2[i];            /* :156 */
    if (tr && g_trace.x_after_attn) memcpy(g_trace.x_after_attn, m->x, (size_t)dim * 4);
    orc_rmsnorm(m->xb, m->x, m->rms_ffn + (size_t)l * dim, dim);   /* :159 */
    orc_quantize(m->xq_q, m->xq_s, m->xb, dim, gs);                /* :161 */
    /* layers.rs:466-480 */
    orc_matmul(m->hb, m->xq_q, m->xq_s, m->w1[l].q, m->w1[l].s, dim, H, gs);
    orc_matmul(m->hb2, m->xq_q, m->xq_s, m->w3[l].q, m->w3[l].s, dim, H, gs);
    for (int i = 0; i < H; i++) {
        float g = m->hb[i];
        float sw = g * (1.0f / (1.0f + expf(-g)));
        m->hb[i] = sw * m->hb2[i];
    }
    if (tr && g_trace.hb_swiglu) memcpy(g_trace.hb_swiglu, m->hb, (size_t)H * 4);
    orc_quantize(m->hq_q, m->hq_s, m->hb, H, gs);
    if (tr && g_trace.hq_q) { memcpy(g_trace.hq_q, m->hq_q, H); memcpy(g_trace.hq_s, m->hq_s, (size_t)(H / gs) * 4); }
    orc_matmul(m->xb, m->hq_q, m->hq_s, m->w2[l].q, m->w2[l].s, H, dim, gs);
    for (int i = 0; i < dim; i++) m->x[i] += m->xb[i];             /* :175 */
    if (tr && g_trace.x_out) memcpy(g_trace.x_out, m->x, (size_t)dim * 4);
}

/* qwen3.rs:62-79 Qwen3Transformer::forward.  Returns NULL (reference: panic on slice
 * index) when token/pos are out of range. */
ORC_API const float *orc_model_forward(OrcModel *m, int token, int pos) {
    const OrcConfig *cf = &m->cfg;
    if (token < 0 || token >= cf->vocab_size || pos < 0 || pos >= cf->seq_len) {
        snprintf(g_err, sizeof g_err, "index out of bounds: token %d pos %d", token, pos);
        return NULL;
    }
    int dim = cf->dim;
    memcpy(m->x, m->token_embedding_table + (size_t)token * dim, (size_t)dim * 4); /* layers.rs:72-76 */
    if (m->xdump) memcpy(m->xdump, m->x, (size_t)dim * 4);
    for (int l = 0; l < cf->n_layers; l++) {
        block_forward(m, l, pos);
        if (m->xdump) memcpy(m->xdump + (size_t)(l + 1) * dim, m->x, (size_t)dim * 4);
    }
    orc_rmsnorm(m->x, m->x, m->rms_final, dim);                                   /* :72 */
    orc_quantize(m->xq_q, m->xq_s, m->x, dim, cf->group_size);                    /* :75 */
    orc_matmul(m->logits, m->xq_q, m->xq_s, m->wcls.q, m->wcls.s, dim, cf->vocab_size, cf->group_size); /* :76 */
    return m->logits;
}

/* Layers [l0, l1) of one decode step on a caller-supplied residual stream (in place); with
 * run_head also final norm + lm_head.  Test helper mirroring q3_forward_layers: lets a test
 * compare one layer at a time with identical inputs (no error cascade). */
ORC_API const float *orc_model_forward_layers(OrcModel *m, int pos, int l0, int l1, float *x_io, int run_head) {
    const OrcConfig *cf = &m->cfg;
    int dim = cf->dim;
    memcpy(m->x, x_io, (size_t)dim * 4);
    for (int l = l0; l < l1; l++) block_forward(m, l, pos);
    memcpy(x_io, m->x, (size_t)dim * 4);
    if (!run_head) return NULL;
    orc_rmsnorm(m->x, m->x, m->rms_final, dim);
    orc_quantize(m->xq_q, m->xq_s, m->x, dim, cf->group_size);
    orc_matmul(m->logits, m->xq_q, m->xq_s, m->wcls.q, m->wcls.s, dim, cf->vocab_size, cf->group_size);
    return m->logits;
}

/* Zero the KV cache (fresh-transformer state; qwen3.rs:439-440). */
ORC_API void orc_model_reset(OrcModel *m) {
    const OrcConfig *cf = &m->cfg;
    size_t n = (size_t)cf->n_layers * cf->seq_len * cf->n_kv_heads * cf->head_dim;
    memset(m->key_cache, 0, n * 4);
    memset(m->value_cache, 0, n * 4);
}
/* Record the residual stream at every layer boundary of subsequent forwards into buf
 * ([(n_layers+1)][dim]); NULL stops recording. */
ORC_API void orc_model_set_xdump(OrcModel *m, float *buf) { m->xdump = buf; }
ORC_API const float *orc_model_key_cache(const OrcModel *m) { return m->key_cache; }
ORC_API const float *orc_model_value_cache(const OrcModel *m) { return m->value_cache; }

/* generation.rs:9-48 generate(), on token ids (tokenizer bypassed), greedy or sampled.
 * Prompt tokens except the last are NOT forwarded (:26-28) -- their cache rows stay zero
 * and still take part in the softmax.  `out_tokens` receives every token sampled by
 * generate_next_token() (:153-162) in order, up to max_new of them; generation stops early
 * at seq_len (:25) or when the sampled token is bos/eos (:170-172; pass -1 to disable; the
 * terminating token is not recorded).  margins (optional) gets the top1-top2 logit gap of
 * each recorded step.  Returns the number of tokens written, -1 on error. */
ORC_API int orc_generate(OrcModel *m, OrcSampler *sampler, const int *prompt, int n_prompt, int max_new,
                         int bos, int eos, int *out_tokens, float *margins) {
    if (n_prompt <= 0) {
        snprintf(g_err, sizeof g_err, "Please provide a prompt");
        return -1;
    }
    int seq_len = m->cfg.seq_len, V = m->cfg.vocab_size;
    int pos = 0, token = prompt[0], n_out = 0;
    float *copy = (float *)malloc((size_t)V * 4);
    while (pos < seq_len && n_out < max_new) {
        int next;
        if (pos < n_prompt - 1) {
            next = prompt[pos + 1];
        } else {
            const float *lg = orc_model_forward(m, token, pos);
            if (!lg) { free(copy); return -1; }
            memcpy(copy, lg, (size_t)V * 4); /* logits.to_vec() */
            if (margins) {
                int b = orc_argmax(copy, V);
                float second = -INFINITY;
                for (int i = 0; i < V; i++) if (i != b && copy[i] > second) second = copy[i];
                margins[n_out] = copy[b] - second;
            }
            next = orc_sampler_sample(sampler, copy);
            if (next == bos || next == eos) break;
            out_tokens[n_out++] = next;
        }
        token = next;
        pos++;
    }
    free(copy);
    return n_out;
}
