/*
 * cc_oracle.c -- CPU ORACLE for the compressed-KV decode path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file restates, in plain C, the algorithm of the reference
 * (AnswerDotAI/cold-compress, pure Python/PyTorch) for one decode step of one
 * attention layer.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it; the product path never does.
 *
 * Parity pin: the reference ships no tests or golden vectors, so the
 * oracle is pinned against outputs of the reference itself, executed unmodified in
 * the build container by tests/golden/make_golden.py; the fixtures are committed
 * under tests/golden/ and checked by tests/test_oracle_golden.py.
 *
 * The arithmetic lives in PyTorch (un-vendored dependency, reference pins
 * torch==2.5.0.dev20240723, fixtures generated with torch 2.11.0): bf16 tensors,
 * fp32 accumulation.  Where PyTorch leaves the fp32 evaluation order unspecified
 * (BLAS dot products, vectorised sums, Sleef exp) this file fixes ONE order, the
 * same one the CUDA kernel uses, so that oracle and kernel agree bit for bit:
 *   - q.k: 8-element fused-multiply-add chains combined by a xor-butterfly tree,
 *   - softmax denominator: exact 2^-40 fixed-point integer sum (order independent),
 *   - exp: the polynomial below, evaluated with explicit fmaf.
 * Against the reference these choices differ by <= 1 ulp(fp32) before each bf16
 * rounding, i.e. a rare 1-ulp(bf16) difference on individual probabilities; the
 * measured rates are recorded by tests/test_oracle_golden.py.
 *
 * Compile with -ffp-contract=off (see Makefile): contraction would change bits.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/cc_b200.h"

/* ------------------------------------------------------------------ bf16 */
static inline float bf2f(uint16_t h) {
    uint32_t u = (uint32_t)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}
/* round-to-nearest-even, as c10::BFloat16 / __float2bfloat16_rn */
static inline uint16_t f2bf(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x0040u);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
static inline float bfround(float f) { return bf2f(f2bf(f)); }

/* ------------------------------------------------------------ exp(x), x <= 0
 * Cephes-style: n = rint(x*log2e), r = x - n*ln2 (two-step), degree-6 polynomial,
 * scale by 2^n.  Only explicit fmaf/mul/add: identical bits on any IEEE machine. */
static inline float or_expf(float x) {
    if (!(x > -104.0f)) return 0.0f; /* also -inf and NaN */
    float t = x * 1.44269504088896341f;
    float n = rintf(t);
    float r = fmaf(n, -0.693359375f, x);
    r = fmaf(n, 2.12194440e-4f, r);
    float p = 1.9875691500e-4f;
    p = fmaf(p, r, 1.3981999507e-3f);
    p = fmaf(p, r, 8.3334519073e-3f);
    p = fmaf(p, r, 4.1665795894e-2f);
    p = fmaf(p, r, 1.6666665459e-1f);
    p = fmaf(p, r, 5.0000001201e-1f);
    float r2 = r * r;
    p = fmaf(p, r2, r);
    p = p + 1.0f;
    int ni = (int)n; /* -150 .. 0 */
    /* scale in two exact steps so results below 2^-126 round once, as a denormal */
    int n1 = ni < -100 ? -100 : ni;
    int n2 = ni - n1;
    uint32_t b1 = (uint32_t)(n1 + 127) << 23, b2 = (uint32_t)(n2 + 127) << 23;
    float s1, s2;
    memcpy(&s1, &b1, 4);
    memcpy(&s2, &b2, 4);
    return (p * s1) * s2;
}

/* ------------------------------------------------------- canonical dot product
 * D/8 partial sums, each an fma chain over 8 consecutive elements (ascending),
 * combined by a butterfly over lane distance D/16, ..., 2, 1. */
static float or_dot(const uint16_t* q, const uint16_t* k, int D) {
    float part[16];
    int nl = D / 8;
    for (int j = 0; j < nl; ++j) {
        float acc = 0.0f;
        for (int i = 0; i < 8; ++i) acc = fmaf(bf2f(q[8 * j + i]), bf2f(k[8 * j + i]), acc);
        part[j] = acc;
    }
    for (int d = nl / 2; d >= 1; d >>= 1) {
        float nxt[16];
        for (int j = 0; j < nl; ++j) nxt[j] = part[j] + part[j ^ d];
        memcpy(part, nxt, sizeof(float) * nl);
    }
    return part[0];
}

static __thread char g_err[256];
const char* cc_oracle_last_error(void) { return g_err; }
static int fail(int code, const char* msg) {
    strncpy(g_err, msg, sizeof(g_err) - 1);
    return code;
}

/* ---------------------------------------------------------- eviction policies */

/* KVCacheRecentGlobal._eviction_idx (cache.py:552-556): oldest position among slots >= g */
static int64_t evict_recent_global(const CcCacheState* st) {
    const int32_t* pos = st->pos;
    int g = st->global_tokens, L = st->max_cache_length;
    int64_t best = g;
    for (int s = g + 1; s < L; ++s)
        if (pos[s] < pos[best]) best = s;
    return best;
}

/* KVCacheL2._token_importances + KVCache._eviction_idx (cache.py:597-605, 366-379) */
static void evict_l2(const CcCacheState* st, int32_t input_pos, int64_t* idx) {
    int H = st->n_kv_heads, L = st->max_cache_length;
    const uint16_t* kn = (const uint16_t*)st->key_norm;
    float M = -INFINITY; /* key_norm.max() spans every head and slot of the layer */
    for (int i = 0; i < H * L; ++i) {
        float v = bf2f(kn[i]);
        if (v > M) M = v;
    }
    for (int h = 0; h < H; ++h) {
        const int32_t* pos = st->pos + (size_t)h * L;
        float best = 0;
        int64_t bi = -1;
        for (int s = 0; s < L; ++s) {
            float sc = bfround(M - bf2f(kn[(size_t)h * L + s])); /* bf16 - bf16 -> bf16 */
            if (pos[s] >= input_pos - st->recent_window) sc = INFINITY;
            if (s < st->global_tokens) sc = INFINITY; /* first g SLOTS */
            if (pos[s] == -1) sc = -INFINITY;
            if (bi < 0 || sc < best) { best = sc; bi = s; }
        }
        idx[h] = bi;
    }
}

/* bf16 ring sum (cache.py:855-859): torch accumulates the W bf16 values in fp32 in an
 * unspecified vectorised order and rounds the result to bf16.  Here the sum is taken in fp64
 * (exact for any realistic dynamic range, hence order independent), rounded to fp32, then to
 * bf16.  The CUDA kernel keeps the same fp64 value as a running sum instead of re-reading the
 * ring. */
static float ring_sum_bf16(const uint16_t* ring, int W) {
    double acc = 0.0;
    for (int w = 0; w < W; ++w) acc += (double)bf2f(ring[w]);
    return bfround((float)acc);
}

/* KVCacheHeavyHitter._eviction_idx (cache.py:725-765).  W == 1: fp64 running sum / max(denom, 1).
 * W > 1: sum of the slot's ring of the last W pooled probabilities (bf16 ring: cache.py:646-667; with
 * attn_thresholding a bool ring of "p >= 1 / cache_cts" indicators, whose sum is an integer count)
 * / clamp(denom, 1, W).  The evicted slot's ring and denominator are zeroed (753-763). */
static void evict_heavy_hitter(const CcCacheState* st, int32_t input_pos, int64_t* idx) {
    int H = st->n_kv_heads, L = st->max_cache_length, W = st->history_window;
    int thresh = st->attn_thresholding != 0;
    for (int h = 0; h < H; ++h) {
        const int32_t* pos = st->pos + (size_t)h * L;
        int32_t* den = st->attn_history_denom + (size_t)h * L;
        float best = 0;
        int64_t bi = -1;
        for (int s = 0; s < L; ++s) {
            float num;
            int32_t d = den[s] < 1 ? 1 : den[s];
            if (W == 1) {
                num = (float)((double*)st->attn_history_num)[(size_t)h * L + s];
            } else if (thresh) {
                const uint8_t* ring = (const uint8_t*)st->attn_history_num + ((size_t)h * L + s) * W;
                int cnt = 0;
                for (int w = 0; w < W; ++w) cnt += ring[w] != 0;
                num = (float)cnt;
                if (d > W) d = W;
            } else {
                num = ring_sum_bf16((const uint16_t*)st->attn_history_num + ((size_t)h * L + s) * W, W);
                if (d > W) d = W;
            }
            float avg = num / (float)d;
            if (pos[s] < st->global_tokens || pos[s] >= input_pos - st->recent_window) avg = 1.0f;
            if (pos[s] == -1) avg = 0.0f;
            if (bi < 0 || avg < best) { best = avg; bi = s; }
        }
        idx[h] = bi;
        /* history of the evicted slot restarts (cache.py:753-763) */
        if (W == 1) ((double*)st->attn_history_num)[(size_t)h * L + bi] = 0.0;
        else if (thresh) memset((uint8_t*)st->attn_history_num + ((size_t)h * L + bi) * W, 0, (size_t)W);
        else memset((uint16_t*)st->attn_history_num + ((size_t)h * L + bi) * W, 0, (size_t)W * 2);
        den[bi] = 0;
    }
}

/* KVCacheHybrid._select_fill_idx / _eviction_idx_for_head (cache.py:844-950).
 * Returns the slot, sets *evict and *stored. */
static int64_t hybrid_select(const CcCacheState* st, int h, int32_t input_pos, int is_punc,
                             int* evict, int* stored) {
    int L = st->max_cache_length, W = st->history_window;
    int bits = st->hyb_strategy_bits[h];
    int32_t cts = st->cache_cts[h];
    int64_t end_idx = cts < L - 1 ? cts : L - 1;
    *evict = 0;
    *stored = 1;
    if ((bits & CC_HYB_PUNC) && is_punc) return end_idx;
    if (bits & CC_HYB_FULL) return end_idx;
    int32_t budget = st->global_tokens;
    if (bits & CC_HYB_SPECIAL) budget += st->num_special ? st->num_special[0] : 0;
    if (bits & CC_HYB_PUNC) budget += st->num_punc ? st->num_punc[0] : 0;
    if (bits & CC_HYB_WINDOW) budget += st->hyb_recent_window[h];
    if (bits & CC_HYB_HEAVY_HITTER) budget += st->hyb_heavy_budget[h];
    if (cts < budget) return end_idx;
    if (!(bits & (CC_HYB_HEAVY_HITTER | CC_HYB_WINDOW))) {
        *stored = 0; /* special/punc-only head over budget: token is dropped */
        return L - 1;
    }
    const int32_t* pos = st->pos + (size_t)h * L;
    const int32_t* den = st->attn_history_denom + (size_t)h * L;
    const uint16_t* ring = (const uint16_t*)st->attn_history_num + (size_t)h * L * W;
    int32_t rw = (bits & CC_HYB_WINDOW) ? st->hyb_recent_window[h] : 0;
    float best = 0;
    int64_t bi = -1;
    int best_nan = 0;
    for (int s = 0; s < cts; ++s) {
        float sc;
        if (bits & CC_HYB_HEAVY_HITTER) {
            int32_t d = den[s] > W ? W : den[s];
            sc = ring_sum_bf16(ring + (size_t)s * W, W) / (float)d;
        } else {
            sc = (float)pos[s];
        }
        int save = s < st->global_tokens;
        if ((bits & CC_HYB_SPECIAL) && st->special_mask) save |= st->special_mask[(size_t)h * L + s];
        if ((bits & CC_HYB_PUNC) && st->punc_mask) save |= st->punc_mask[(size_t)h * L + s];
        if ((bits & CC_HYB_WINDOW) && pos[s] > input_pos - rw) save = 1; /* strict > */
        if (save) sc = INFINITY;
        /* torch.argmin: a NaN is the minimum, first one wins */
        int is_nan = sc != sc;
        if (bi < 0 || (!best_nan && (is_nan || sc < best))) {
            best = sc;
            bi = s;
            best_nan = is_nan;
        }
    }
    *evict = 1;
    return bi < 0 ? 0 : bi;
}

/* ------------------------------------------------------------------ attention
 * One KV head: logits over all L slots for the G query heads that share it,
 * softmax, P.V, GQA pooling.  probs_out (bf16 [L]) only in NAIVE mode. */
static void attend_head(const CcCacheState* st, int h, const uint16_t* q, uint16_t* y,
                        uint16_t* probs_out, float* lg /* [G*L] scratch */) {
    int L = st->max_cache_length, D = st->head_dim, G = st->q_per_kv;
    const uint16_t* K = (const uint16_t*)st->k_cache + (size_t)h * L * D;
    const uint16_t* V = (const uint16_t*)st->v_cache + (size_t)h * L * D;
    const uint8_t* mask = st->mask + (size_t)h * L;
    int naive = st->attn_mode == CC_ATTN_NAIVE_BF16;
    float scale = (float)(1.0 / sqrt((double)D));
    for (int g = 0; g < G; ++g) {
        const uint16_t* qg = q + (size_t)(h * G + g) * D;
        float* l = lg + (size_t)g * L;
        float m = -INFINITY;
        for (int s = 0; s < L; ++s) {
            float v;
            if (!mask[s]) {
                v = -INFINITY;
            } else {
                float d = or_dot(qg, K + (size_t)s * D, D);
                /* naive path rounds q@k^T to bf16, then *scale to bf16 (attention_utils.py:37) */
                v = naive ? bfround(bfround(d) * scale) : d * scale;
            }
            l[s] = v;
            if (v > m) m = v;
        }
        int64_t isum = 0;
        for (int s = 0; s < L; ++s) {
            float e = or_expf(l[s] - m);
            l[s] = e;
            isum += (int64_t)(e * 1099511627776.0f); /* 2^40, truncation */
        }
        float sum = (float)isum * 9.094947017729282e-13f; /* 2^-40 */
        for (int s = 0; s < L; ++s) {
            float p = l[s] / sum;
            l[s] = naive ? bfround(p) : p;
        }
        float acc[128];
        for (int d = 0; d < D; ++d) acc[d] = 0.0f;
        for (int s = 0; s < L; ++s) {
            float p = l[s];
            if (p == 0.0f) continue;
            const uint16_t* vr = V + (size_t)s * D;
            for (int d = 0; d < D; ++d) acc[d] = fmaf(p, bf2f(vr[d]), acc[d]);
        }
        uint16_t* yg = y + (size_t)(h * G + g) * D;
        for (int d = 0; d < D; ++d) yg[d] = f2bf(acc[d]);
    }
    if (naive && probs_out) {
        /* model.py:413-418: mean over the G query heads of the group, bf16 */
        float gf = (float)G; /* torch mean = fp32 sum / N, not * (1/N): pinned by hh_g7_bias */
        for (int s = 0; s < L; ++s) {
            float a = lg[s];
            for (int g = 1; g < G; ++g) a = a + lg[(size_t)g * L + s];
            probs_out[s] = f2bf(a / gf);
        }
    }
}

/* ------------------------------------------------------------- the decode step
 * Same contract as cc_decode_step, host pointers, phases honoured.
 * Follows model.py:389-427 -> cache.py:314-364 -> attention_utils.py:8-54 ->
 * cache.py:690-723. */
int cc_oracle_decode_step(const CcCacheState* st, const CcDecodeArgs* a) {
    if (!st || !a) return fail(CC_ERR_INVALID_ARG, "null state/args");
    int H = st->n_kv_heads, L = st->max_cache_length, D = st->head_dim, G = st->q_per_kv;
    if (D != 64 && D != 128) return fail(CC_ERR_UNSUPPORTED_SHAPE, "head_dim must be 64 or 128");
    if (H < 1 || G < 1 || L < 1) return fail(CC_ERR_INVALID_ARG, "bad shape");
    int32_t input_pos = a->input_pos[0];
    int phases = a->phases ? a->phases : CC_PHASE_ALL;
    int is_punc = a->is_punc_dev ? a->is_punc_dev[0] != 0 : a->is_punc != 0;
    int naive = st->attn_mode == CC_ATTN_NAIVE_BF16;
    int64_t idx[256];
    int stored[256], evict[256];
    if (H > 256) return fail(CC_ERR_UNSUPPORTED_SHAPE, "too many heads");

    if (phases & CC_PHASE_EVICT_FILL) {
        for (int h = 0; h < H; ++h) { stored[h] = 1; evict[h] = 0; }
        switch (st->policy) {
        case CC_POLICY_EXTERNAL:
            if (!a->fill_idx) return fail(CC_ERR_INVALID_ARG, "external policy needs fill_idx");
            for (int h = 0; h < H; ++h) idx[h] = a->fill_idx[st->pos_heads == 1 ? 0 : h];
            break;
        case CC_POLICY_RECENT_GLOBAL: {
            int64_t i = evict_recent_global(st);
            for (int h = 0; h < H; ++h) idx[h] = i;
            break;
        }
        case CC_POLICY_L2: evict_l2(st, input_pos, idx); break;
        case CC_POLICY_HEAVY_HITTER:
            evict_heavy_hitter(st, input_pos, idx);
            break;
        case CC_POLICY_HYBRID:
            for (int h = 0; h < H; ++h)
                idx[h] = hybrid_select(st, h, input_pos, is_punc, &evict[h], &stored[h]);
            break;
        default: return fail(CC_ERR_INVALID_ARG, "bad policy");
        }
        /* num_insertions = (pos[idx] == -1)  (cache.py:356-360), before the overwrite */
        int32_t incr[256];
        for (int h = 0; h < H; ++h) {
            const int32_t* pos = st->pos + (size_t)(st->pos_heads == 1 ? 0 : h) * L;
            incr[h] = pos[idx[h]] == -1;
        }
        /* _fill (cache.py:390-401 / 460-490) */
        for (int h = 0; h < H; ++h) {
            int64_t i = idx[h];
            memcpy((uint16_t*)st->k_cache + ((size_t)h * L + i) * D, (const uint16_t*)a->k_new + (size_t)h * D, 2 * D);
            memcpy((uint16_t*)st->v_cache + ((size_t)h * L + i) * D, (const uint16_t*)a->v_new + (size_t)h * D, 2 * D);
            if (st->pos_heads == 1) { if (h == 0) st->pos[i] = input_pos; }
            else st->pos[(size_t)h * L + i] = input_pos;
            if (st->policy == CC_POLICY_HYBRID) {
                /* mask only for heads that grew; evicting heads keep mask[idx] == 1; dropped untouched */
                if (stored[h] && !evict[h]) st->mask[(size_t)h * L + i] = 1;
                if (is_punc && st->punc_mask) st->punc_mask[(size_t)h * L + i] = 1;
            } else {
                st->mask[(size_t)h * L + i] = 1;
            }
            if (st->policy == CC_POLICY_L2) {
                /* key_norm[idx] = ||k_new||_2 (cache.py:592-593): fp32 sum of squares, sqrt, bf16 */
                const uint16_t* kr = (const uint16_t*)a->k_new + (size_t)h * D;
                float ss = or_dot(kr, kr, D);
                ((uint16_t*)st->key_norm)[(size_t)h * L + i] = f2bf(sqrtf(ss));
            }
        }
        if (st->policy == CC_POLICY_HYBRID) {
            for (int h = 0; h < H; ++h) st->cache_cts[h] += (stored[h] && !evict[h]) ? 1 : 0;
            if (is_punc && st->num_punc) st->num_punc[0] += 1;
        } else {
            st->cache_cts[0] += incr[0]; /* cache_cts += num_insertions[:1] (cache.py:330) */
        }
        if (a->fill_idx && st->policy != CC_POLICY_EXTERNAL)
            for (int h = 0; h < H; ++h) a->fill_idx[h] = idx[h];
    }

    uint16_t* probs = (uint16_t*)a->probs;
    uint16_t* own_probs = NULL;
    int need_probs = naive && (phases & (CC_PHASE_ATTEND | CC_PHASE_SCORE));
    if (need_probs && !probs) {
        own_probs = (uint16_t*)malloc((size_t)H * L * 2);
        probs = own_probs;
    }
    if (phases & CC_PHASE_ATTEND) {
        if (!a->q || !a->y) return fail(CC_ERR_INVALID_ARG, "attend needs q and y");
#pragma omp parallel for schedule(static)
        for (int h = 0; h < H; ++h) {
            float* lg = (float*)malloc((size_t)G * L * sizeof(float));
            attend_head(st, h, (const uint16_t*)a->q, (uint16_t*)a->y,
                        naive ? probs + (size_t)h * L : NULL, lg);
            free(lg);
        }
    }
    if ((phases & CC_PHASE_SCORE) && naive &&
        (st->policy == CC_POLICY_HEAVY_HITTER || st->policy == CC_POLICY_HYBRID)) {
        /* KVCacheHeavyHitter.update_state decode branch (cache.py:716-723) */
        int W = st->history_window;
        if (W == 1) {
            double* num = (double*)st->attn_history_num;
            for (size_t i = 0; i < (size_t)H * L; ++i) num[i] += (double)bf2f(probs[i]);
        } else if (st->policy == CC_POLICY_HEAVY_HITTER && st->attn_thresholding) {
            /* (attn >= 1 / cache_cts).int() stored into a bool ring (cache.py:708-709, 719): bf16 vs fp32 compare */
            uint8_t* ring = (uint8_t*)st->attn_history_num;
            int64_t w = st->attn_counter[0] % W;
            float thr = 1.0f / (float)st->cache_cts[0];
            for (size_t i = 0; i < (size_t)H * L; ++i) ring[i * W + w] = bf2f(probs[i]) >= thr;
        } else {
            uint16_t* ring = (uint16_t*)st->attn_history_num;
            int64_t w = st->attn_counter[0] % W;
            for (size_t i = 0; i < (size_t)H * L; ++i) ring[i * W + w] = probs[i];
        }
        for (size_t i = 0; i < (size_t)H * L; ++i) st->attn_history_denom[i] += 1;
        st->attn_counter[0] += 1;
    }
    free(own_probs);
    return CC_OK;
}

/* KVCacheL2.update_state prefill branch (cache.py:611-612): bf16 L2 norm of each cached key */
int cc_oracle_key_norms(const void* k, void* key_norm, int32_t n_rows, int32_t head_dim) {
    if (head_dim != 64 && head_dim != 128) return fail(CC_ERR_UNSUPPORTED_SHAPE, "head_dim");
    const uint16_t* K = (const uint16_t*)k;
    uint16_t* out = (uint16_t*)key_norm;
    for (int r = 0; r < n_rows; ++r) {
        const uint16_t* kr = K + (size_t)r * head_dim;
        out[r] = f2bf(sqrtf(or_dot(kr, kr, head_dim)));
    }
    return CC_OK;
}

/* exposed for unit tests of the shared numeric definitions */
float cc_oracle_expf(float x) { return or_expf(x); }
float cc_oracle_dot(const void* q, const void* k, int D) { return or_dot((const uint16_t*)q, (const uint16_t*)k, D); }
int cc_oracle_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
