/*
 * cc_b200.h -- C ABI of the B200-native compressed-KV decode path.
 *
 * The reference (AnswerDotAI/cold-compress) is pure Python and has no FFI; the
 * entry points below are what a binding for its hot path would call.  Each one
 * names the reference interface it replaces (file:line relative to the
 * reference checkout).  Plain pointers and sizes only: no torch types, no C++
 * exceptions across the boundary.  Every function returns 0 on success and a
 * negative CC_ERR_* code on failure; cc_last_error() gives the message of the
 * last failure on the calling thread.
 *
 * All buffer pointers are DEVICE pointers owned by the caller (in the Python
 * host they are the nn.Module buffers the reference registers in
 * cache.py:195-227 / 574 / 668-679).  The library owns nothing persistent.
 * Every call enqueues work on the caller's stream and returns without
 * synchronising, so a whole decode step can be captured in a CUDA graph.
 */
#ifndef CC_B200_H
#define CC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CC_ABI_VERSION 1

/* ---- error codes ------------------------------------------------------- */
#define CC_OK                     0
#define CC_ERR_INVALID_ARG       -1  /* NULL pointer, bad enum, bad size        */
#define CC_ERR_UNSUPPORTED_SHAPE -2  /* head_dim / group / length out of range  */
#define CC_ERR_CUDA              -3  /* a CUDA runtime call failed              */
#define CC_ERR_NO_DEVICE         -4  /* no sm_100 device / kernel image missing */

/* ---- eviction policies (cache.py:1444-1478 get_cache_constructor) ------- */
#define CC_POLICY_EXTERNAL       0  /* slot index supplied by the caller: full (cache.py:500-502),
                                       random (519-524), keep_it_odd (1437-1441) compute it host-side */
#define CC_POLICY_RECENT_GLOBAL  1  /* cache.py:552-556 */
#define CC_POLICY_L2             2  /* cache.py:580-605 */
#define CC_POLICY_HEAVY_HITTER   3  /* cache.py:690-765 */
#define CC_POLICY_HYBRID         4  /* cache.py:844-1019 (FastGen, per-head strategy) */

/* ---- attention numerics (attention_utils.py:8-54) ----------------------- */
#define CC_ATTN_MATH_FP32   0  /* F.scaled_dot_product_attention, MATH backend: fp32 inside,
                                  one rounding of y (attention_utils.py:27-35)              */
#define CC_ATTN_NAIVE_BF16  1  /* return_attn path: bf16 after q.k, after *scale, after
                                  softmax, after the GQA mean (attention_utils.py:36-54,
                                  model.py:413-418)                                         */

/* ---- phase flags: the fused step is EVICT|ATTEND|SCORE ------------------- */
#define CC_PHASE_EVICT_FILL  1  /* KVCache.update_kv decode branch (cache.py:314-364)       */
#define CC_PHASE_ATTEND      2  /* scaled_dot_product_attention + GQA pooling               */
#define CC_PHASE_SCORE       4  /* KVCache.update_state decode branch (cache.py:690-723)    */
#define CC_PHASE_ALL         7

/* ---- per-call flags ------------------------------------------------------ */
#define CC_FLAG_OVERLAP_PREDECESSOR 1  /* the kernel launched just before this one on the stream does
                                          not write this layer's cache or policy state (it is the q/k/v
                                          projection, or another layer's step): cc_decode_step may then
                                          start early (programmatic dependent launch) and fetch K/V and
                                          policy state while that kernel is still running               */

#define CC_FLAG_VIA_TOKEN_KERNEL     2  /* tests: run this step through the attention stage of the whole-token
                                          kernel (cc_decode_token) instead of the per-layer kernel: same results,
                                          synchronous, not capturable                                            */

/* ---- hybrid per-head strategy bits (cache.py:896-950 substrings of the name) */
#define CC_HYB_FULL          1
#define CC_HYB_WINDOW        2
#define CC_HYB_HEAVY_HITTER  4
#define CC_HYB_SPECIAL       8
#define CC_HYB_PUNC          16

/*
 * Per-layer cache state: mirrors the buffers of one KVCache module
 * (cache.py:162-227).  Shapes drop the leading batch dimension (always 1,
 * model.py:188-189).  bf16 tensors are passed as void*.
 */
typedef struct CcCacheState {
    int32_t policy;            /* CC_POLICY_*                                            */
    int32_t attn_mode;         /* CC_ATTN_*                                              */
    int32_t n_kv_heads;        /* Hkv: local KV heads (config.n_local_heads)             */
    int32_t q_per_kv;          /* G = n_head / n_local_heads (model.py:395-400)          */
    int32_t head_dim;          /* D: 64 or 128                                           */
    int32_t max_cache_length;  /* L: slots per head                                      */
    int32_t global_tokens;     /* g                                                      */
    int32_t recent_window;     /* per-layer int after setup_caches (generation_utils.py:362-372) */
    int32_t history_window;    /* W: 1 => fp64 running sum, >1 => bf16 ring (cache.py:661-667)   */
    int32_t pos_heads;         /* rows of `pos`: 1 (head-constant) or Hkv (head-specific)        */
    int32_t cts_heads;         /* rows of `cache_cts`: 1, or Hkv for variable_length (hybrid)    */
    int32_t attn_thresholding; /* heavy_hitter, W > 1: history holds bool indicators p >= 1 / cache_cts (cache.py:708) */

    void*    k_cache;          /* bf16 [Hkv, L, D]   cache.py:185,200                    */
    void*    v_cache;          /* bf16 [Hkv, L, D]   cache.py:186,201                    */
    int32_t* pos;              /* i32  [pos_heads, L], -1 = unfilled  cache.py:207-218   */
    uint8_t* mask;             /* bool [Hkv, L]      cache.py:226-227                    */
    int32_t* cache_cts;        /* i32  [cts_heads]   cache.py:219-222                    */

    /* heavy_hitter / hybrid (cache.py:646-679) */
    void*    attn_history_num;   /* f64 [Hkv, L] when W == 1; bf16 [Hkv, L, W] otherwise (bool [Hkv, L, W] with
                                    attn_thresholding)                                     */
    int32_t* attn_history_denom; /* i32 [Hkv, L]                                          */
    int64_t* attn_counter;       /* i64 [1]                                               */
    double*  attn_history_sum;   /* W > 1 (hybrid, windowed heavy_hitter): f64 [Hkv, L] running sum of each slot's ring,
                                    kept by the kernel so the ring is not re-summed every step (the
                                    reference re-sums it, cache.py:855-859); NULL otherwise            */

    /* l2 (cache.py:573-574) */
    void*    key_norm;           /* bf16 [Hkv, L]                                         */

    /* hybrid (cache.py:803-834, 896-950); NULL for other policies */
    const int32_t* hyb_strategy_bits;  /* i32 [Hkv] CC_HYB_* of the strategy chosen per head */
    const int32_t* hyb_recent_window;  /* i32 [Hkv] round(recent_window * L)                 */
    const int32_t* hyb_heavy_budget;   /* i32 [Hkv] round(heavy_hitter_frac * L)             */
    uint8_t* special_mask;             /* bool [Hkv, L] or NULL                              */
    uint8_t* punc_mask;                /* bool [Hkv, L] or NULL                              */
    int32_t* num_special;              /* i32 [1] or NULL                                    */
    int32_t* num_punc;                 /* i32 [1] or NULL                                    */

    /* 8-byte-aligned scratch of cc_sync_words() int32 words, zero-initialised once by the
       caller and after every reset(); orders shared-state writes after all heads have read. */
    int32_t* sync_words;
} CcCacheState;

/* Per-call arguments of one decode step (one new token, one layer). */
typedef struct CcDecodeArgs {
    const void*    q;          /* bf16 [Hkv*G, D], post-RoPE (model.py:382-385)          */
    const void*    k_new;      /* bf16 [Hkv, D],   post-RoPE                             */
    const void*    v_new;      /* bf16 [Hkv, D]                                          */
    const int32_t* input_pos;  /* i32 [1] on the device (generation_utils.py:482)        */
    void*          y;          /* out bf16 [Hkv*G, D]                                    */
    int64_t*       fill_idx;   /* policy EXTERNAL: in, i64 [pos_heads]; otherwise optional
                                  out i64 [Hkv] (the eviction trace), may be NULL        */
    void*          probs;      /* NAIVE mode: optional out bf16 [Hkv, L] GQA-pooled probabilities
                                  (what update_state receives); with phases == CC_PHASE_SCORE
                                  it is the input.  May be NULL in the fused step.        */
    int32_t        is_punc;    /* hybrid: 1 if the new token id is punctuation (cache.py:969-971) */
    int32_t        phases;     /* CC_PHASE_* mask; CC_PHASE_ALL for the fused step        */
    int32_t        flags;      /* CC_FLAG_*                                                */
    int32_t        n_prefetch; /* number of valid entries in prefetch[] (0..3)            */
    /* Optional L2 warm-up: read-only device ranges (16-byte aligned, e.g. the weights of the
       projections that run right after this step) that the kernel asks the memory system to pull
       into L2 while it is itself waiting on latencies and HBM is idle.  Pure hint. */
    struct { const void* ptr; uint64_t bytes; } prefetch[3];
    const uint8_t* is_punc_dev; /* hybrid: optional device flag (bool [1]) that overrides is_punc, so the
                                   step can be captured in a CUDA graph without a host read           */
} CcDecodeArgs;

/* ABI / build identification. */
int         cc_abi_version(void);
const char* cc_build_info(void);      /* e.g. "sm_100a nvcc 12.9" */
const char* cc_last_error(void);      /* message of the last failing call on this thread */
int         cc_sync_words(void);      /* number of int32 words `sync_words` must hold     */

/*
 * One decode step of one layer, as ONE kernel launch on `stream`:
 * eviction decision -> in-place slot overwrite -> masked attention over the
 * cache slots -> GQA pooling -> eviction-score update.
 * Replaces model.py:389-427 (decode branch) = KVCache.update_kv (cache.py:314-340)
 * + scaled_dot_product_attention (attention_utils.py:8-54) + KVCache.update_state.
 * `stream` is a cudaStream_t passed as void*.
 */
int cc_decode_step(const CcCacheState* state, const CcDecodeArgs* args, void* stream);

/* Launch-shape query for tests / the bench: cluster size the step would use. */
int cc_decode_cluster_size(const CcCacheState* state);

/* Diagnostics: device buffer of [n_CTAs][16] u64 that later cc_decode_step launches fill with
   %globaltimer stamps at their phase boundaries (profiles/trace_k1.py); NULL turns it off. */
void cc_debug_set_trace(void* device_buffer);

/*
 * Key L2 norms of cached keys: key_norm[h, s] = bf16(||k_cache[h, s, :]||_2).
 * Replaces KVCacheL2.update_state prefill branch (cache.py:607-612).
 */
int cc_key_norms(const void* k, void* key_norm, int32_t n_rows, int32_t head_dim, void* stream);

/*
 * Prompt compression gather (prompt_compression.py:82-88 + cache.py:381-401):
 * k_cache[h, j, :] = k_prompt[h, keep_idx[h or 0, j], :] for j < n_keep, same for V,
 * pos[h, j] = keep_idx[..], mask[h, j] = 1.
 */
int cc_prefill_gather(const CcCacheState* state, const void* k_prompt, const void* v_prompt,
                      const int64_t* keep_idx, int32_t keep_heads, int32_t n_keep,
                      int32_t prompt_len, void* stream);

/*
 * Quantised KV cache (--cache_bits 2 / 4 / 8): the round trip dequantize(quantize(cache)) the reference applies to
 * the WHOLE cache in every update_kv (cache.py:283-338; quantization_utils.py:4-46), in place on the bf16 working
 * cache [Hkv, L, D].  Groups are one slot across all heads and channels (axis=2); min/max, scale = (max - min)
 * .clamp(1e-6) / (2^n - 1), zero point = min + scale * 2^(n-1), code = round((x - min) / scale) clamped, value =
 * (code - 2^(n-1)) * scale + zero point, every step a bf16 op as in the reference.  Optional outputs: the packed codes
 * (Hkv*L*D*n_bit/8 bytes, quantization_utils.pack_low_bit_tensor order) and the per-slot statistics (bf16 [L]).
 */
int cc_requantize_cache(void* k_cache, void* v_cache, int32_t n_kv_heads, int32_t max_cache_length, int32_t head_dim,
                        int32_t n_bit, uint8_t* k_codes, uint8_t* v_codes, void* k_scales, void* k_zero_points,
                        void* v_scales, void* v_zero_points, void* stream);

/*
 * Causal prompt attention on the tensor cores (tcgen05) with the reductions of the GQA-pooled
 * probability map that prompt compression and the score-keeping caches consume, in one pass:
 * the [Hq, S, S] attention map of the reference is never materialised.
 * Replaces, in the prefill branch of Attention.forward (model.py:399-427):
 *   scaled_dot_product_attention(q, k, v, causal mask, return_attn=True)   attention_utils.py:36-54
 *   attn.view(1, Hkv, G, S, S).mean(2)                                     model.py:413-418
 *   attn[:, :, -obs_len:, :].mean(2), attn.sum(2)                          prompt_compression.py:175, 192; cache.py:704
 *   the sliding-window masses of KVCacheHybrid.build_masks                 cache.py:1066-1173
 * q [Hkv*G, S, D], k / v [Hkv, S, D], y [Hkv*G, S, D]: bf16, D contiguous, row / head strides in ELEMENTS
 * (multiples of 8), so the [S, H, D] layout the projections produce is taken as is.
 * Outputs (fp32, the caller applies the reference's final bf16 roundings):
 *   col_sum [Hkv, S]            sum over query rows i of pooled[h, i, j]        (NULL: attention only)
 *   obs_sum [Hkv, S]            the same over the last obs_len query rows       (NULL or obs_len 0: skipped)
 *   band_sum [n_windows, Hkv, S] sum over rows i with i - j < windows[w]        (n_windows 0: skipped)
 * where pooled[h, i, j] = bf16(mean over the G query heads of bf16 softmax probabilities).
 * attn_mode CC_ATTN_NAIVE_BF16 keeps the reference's naive-path roundings (bf16 after q.k, after *scale,
 * after softmax); CC_ATTN_MATH_FP32 keeps fp32 logits (F.scaled_dot_product_attention semantics).
 * workspace: device scratch of at least cc_prefill_attention_workspace(...) bytes when col_sum != NULL.
 */
typedef struct CcPrefillAttnArgs {
    const void* q;  int64_t q_stride_s, q_stride_h;
    const void* k;  int64_t k_stride_s, k_stride_h;
    const void* v;  int64_t v_stride_s, v_stride_h;
    void*       y;  int64_t y_stride_s, y_stride_h;
    int32_t seq_len, n_kv_heads, q_per_kv, head_dim;
    int32_t attn_mode;       /* CC_ATTN_*                                   */
    int32_t obs_len;         /* SnapKV observation window (prompt_compression.py:157) */
    int32_t n_windows;       /* 0..4                                        */
    int32_t windows[4];      /* window sizes in tokens                      */
    int32_t reserved0;
    float*  col_sum;
    float*  obs_sum;
    float*  band_sum;
    void*   workspace;
    uint64_t workspace_bytes;
} CcPrefillAttnArgs;

uint64_t cc_prefill_attention_workspace(int32_t seq_len, int32_t n_kv_heads, int32_t q_per_kv,
                                        int32_t obs_len, int32_t n_windows);
int cc_prefill_attention(const CcPrefillAttnArgs* args, void* stream);

/*
 * Tensor-parallel exchange fused into the projections (tp.py:134-138, 156-160: the all-reduce of the
 * row-parallel wo / w2 outputs).  Every rank owns an exchange buffer u64 [n_ops][world][dim / 2] in peer-mapped
 * (NVLink) memory, zero-initialised.  Word [op][r][i] = {bf16 x 2: partial sums of rows 2i, 2i+1 from rank r |
 * epoch << 32}: the flag travels IN the data word (8-byte stores are delivered atomically), so there are no
 * fences, flag arrays or counters on the path.  A projection called with send_op >= 0 does not write `out`: the
 * lane that owns a row pair stores its word straight into slot [send_op][rank] of EVERY rank's buffer.  A
 * projection called with recv_op >= 0 polls the `world` slots of that op until every word carries the current
 * epoch, then uses x = bf16(residual + bf16(sum over ranks, in rank order, of the partials)) as its input (and
 * CTA 0 writes x to recv_out).  The caller increments *epoch once per token after the last kernel; n_ops >= 2
 * distinct ops per token keep a slot from being rewritten before every rank has consumed it.
 */
typedef struct CcTpLink {
    int32_t world, rank, dim, n_ops;
    const uint64_t* peer_buf;    /* device array [world]: base address of each rank's exchange buffer */
    const uint64_t* peer_flag;   /* unused by the flag-in-data protocol (kept for ABI stability), may be NULL */
    int32_t*        cta_done;    /* unused, may be NULL                                                       */
    const int32_t*  epoch;       /* local device i32 [1], >= 1                                         */
} CcTpLink;

typedef struct CcTpArgs {
    const CcTpLink* link;
    int32_t send_op;             /* -1: write `out` locally as usual                                   */
    int32_t recv_op;             /* -1: read `x` as usual; else `x` is the residual stream             */
    void*   recv_out;            /* bf16 [dim]: the summed input, written by CTA 0 (may be NULL)       */
} CcTpArgs;

/*
 * The rest of the decode layer (one token): bf16 matrix-vector products with the elementwise
 * work of the layer fused in.  x/out are device vectors, weights are nn.Linear weights [N, K]
 * row-major.  norm_w != NULL applies the reference's RMSNorm (model.py:446-457) to x first.
 *
 * cc_rmsnorm_qkv_rope: attention_norm -> wqkv (+bias) -> rotary embedding at input_pos
 *   (model.py:376-387, 507-519).  freqs_cis = the precompute_freqs_cis table [max_pos, D/2, 2].
 * cc_gemv_residual:   out = [residual +] W x      (wo: model.py:431,317; w2: 443,326; output head: 290-291)
 * cc_rmsnorm_ffn_gate: out = silu(w1 xn) * (w3 xn), xn = ffn_norm(x)   (model.py:442-443)
 * tp (may be NULL) fuses the tensor-parallel all-reduce into producer / consumer, see CcTpArgs.
 */
int cc_rmsnorm_qkv_rope(const void* x, const void* norm_w, float eps, const void* wqkv, const void* bias,
                        const void* freqs_cis, const int32_t* input_pos, void* q, void* k, void* v,
                        int32_t dim, int32_t n_q_heads, int32_t n_kv_heads, int32_t head_dim,
                        const CcTpArgs* tp, void* stream);
int cc_gemv_residual(const void* x, const void* norm_w, float eps, const void* w, const void* residual,
                     void* out, int32_t K, int32_t N, const CcTpArgs* tp, void* stream);
int cc_rmsnorm_ffn_gate(const void* x, const void* norm_w, float eps, const void* w1, const void* w3,
                        void* out, int32_t K, int32_t N, const CcTpArgs* tp, void* stream);

/*
 * One decoded token as ONE persistent kernel (generation_utils.py:163-178 decode_one_token = Transformer.forward
 * model.py:268-292 with is_prefill=False, followed by greedy sampling generation_utils.py:136-142):
 * embedding row -> per layer [attention_norm + wqkv (+bias) + RoPE | eviction + overwrite + attention + pooling +
 * score update (the cc_decode_step stage) | wo + residual | ffn_norm + silu(w1 x) * w3 x | w2 + residual] -> norm ->
 * output head -> arg-max token (+ softmax probabilities).  One CTA per SM stays resident for the whole token; every
 * warp streams its share of the weight rows (and of the K/V rows) through a private TMA-fed shared-memory ring that
 * runs ahead across stage boundaries, and the stages hand vectors to each other through flag-in-data words
 * (u64 = {payload, token tag}) in `scratch`, so there is no grid barrier and no kernel boundary inside a token.
 * With world > 1 (tp.py head / column sharding) the row-parallel wo / w2 stages store their partial vectors into
 * every rank's scratch over NVLink and the consumer stage sums them: the all-reduces of tp.py:134-160 are part of
 * the same kernel.  Policies: recent_global, l2, heavy_hitter (full history); other caches use cc_decode_step.
 */
typedef struct CcTokenLayer {
    const void* attention_norm;   /* bf16 [dim]                          model.py:310-313 */
    const void* wqkv;             /* bf16 [(Hq + 2 Hkv) * D, dim]        model.py:339     */
    const void* bqkv;             /* bf16 [(Hq + 2 Hkv) * D] or NULL     (Qwen2)          */
    const void* wo;               /* bf16 [dim, Hq * D]                  model.py:340     */
    const void* ffn_norm;         /* bf16 [dim]                                          */
    const void* w1;               /* bf16 [intermediate, dim]            model.py:438-440 */
    const void* w3;               /* bf16 [intermediate, dim]                             */
    const void* w2;               /* bf16 [dim, intermediate]                             */
    float attention_norm_eps, ffn_norm_eps;
    CcCacheState cache;           /* this layer's KV cache                                */
    int64_t* fill_idx;            /* optional out i64 [Hkv]: the eviction trace           */
    void*    probs;               /* optional out bf16 [Hkv, L]: GQA-pooled probabilities */
} CcTokenLayer;

typedef struct CcTokenModel {
    int32_t n_layers;
    int32_t dim;                  /* model width (never sharded)                          */
    int32_t n_q_heads, n_kv_heads;/* LOCAL heads of this rank (tp.py:163-168)             */
    int32_t head_dim;
    int32_t intermediate;         /* LOCAL rows of w1 / w3                                */
    int32_t vocab;                /* rows of `output`                                     */
    float   norm_eps;
    const CcTokenLayer* layers;   /* HOST array [n_layers]                                */
    const void* tok_embeddings;   /* bf16 [vocab, dim]                                    */
    const void* norm;             /* bf16 [dim]                                           */
    const void* output;           /* bf16 [vocab, dim]                                    */
    const void* freqs_cis;        /* bf16 [max_pos, D/2, 2]                               */
    int32_t world, rank;          /* tensor-parallel ranks (1, 0 without)                 */
    const uint64_t* scratch_peers;/* HOST array [world]: device address of every rank's scratch (peer-mapped);
                                     entry [rank] is the local one                         */
    uint64_t scratch_bytes;       /* size of each scratch, >= cc_token_scratch_bytes()    */
} CcTokenModel;

/* Bytes of zero-initialised device scratch a plan for this model needs (exchange words + tags). */
uint64_t cc_token_scratch_bytes(const CcTokenModel* model);
/* Builds the stage program for the model on the current device; *plan is an opaque handle. */
int cc_token_plan_create(const CcTokenModel* model, void** plan);
int cc_token_plan_destroy(void* plan);
/* One decoded token.  token_in / input_pos: device i32 [1].  token_out: device i32 [1] (arg-max of the logits, first
 * index on ties).  logits_out: bf16 [vocab] or NULL.  probs_out: bf16 [vocab] softmax(logits) or NULL. */
int cc_decode_token(void* plan, const int32_t* token_in, const int32_t* input_pos, int32_t* token_out,
                    void* logits_out, void* probs_out, void* stream);
/* Tests / diagnostics: where a stage's output vector lives inside the scratch.  which: 0 q|k|v, 1 y, 2 wo partials
 * [world][dim], 3 h, 4 gate, 5 w2 partials [world][dim], 6 layer input x (layer >= 1; layer == n_layers: head input).
 * Words are u64 = {bf16 x 2 | tag << 32}. */
int cc_token_buffer(void* plan, int32_t layer, int32_t which, uint64_t* offset_bytes, int32_t* n_words);
/* Diagnostics: device buffer u64 [n_stages][n_ctas][12] of %globaltimer stamps: 0 stage entered, 1 input staged,
 * 2..9 warp w done (projection stages) / phase boundaries (attention stage). */
int cc_token_set_trace(void* plan, void* device_buffer);
int cc_token_info(void* plan, int32_t* n_ctas, int32_t* n_stages_program, int32_t* ring_stages, int32_t* smem_bytes);

#ifdef __cplusplus
}
#endif
#endif /* CC_B200_H */
