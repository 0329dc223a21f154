// decode_token.cu -- one decoded token as ONE persistent sm_100a kernel.
//
// Replaces the per-token launch chain of the reference's decode_one_token (generation_utils.py:163-178 ->
// Transformer.forward model.py:268-292 -> 32 x TransformerBlock.forward model.py:307-327 -> greedy sampling
// generation_utils.py:136-142), which round 1 ran as 5 launches per layer.  A decoded token streams 15 GB of bf16
// weights plus the live K/V rows exactly once; the only way to keep HBM busy across the ~160 dependent stages of a
// token is to never stop the stream at a stage boundary.  Design:
//
//   * one CTA per SM (cooperative launch: all co-resident), 8 warps; the rows of every stage are dealt to CTAs and,
//     inside a CTA, to warps, statically;
//   * every warp owns a private ring of 8 KiB shared-memory stages fed by the TMA engine (cp.async.bulk, UBLKCP):
//     the moment a warp has consumed a stage it re-arms it with the NEXT tile of its own static tile sequence --
//     the sequence runs across stage boundaries (weights of wo / w1 / w3 / w2 / the next layer, this layer's K/V
//     rows), so up to ~190 KiB per SM are in flight or resident while the warp waits for a predecessor stage;
//   * stages hand vectors to each other through flag-in-data words in global memory (u64 = {payload, token tag},
//     8-byte single-copy atomic): a consumer spins on the tag of the words it needs.  No grid barrier, no fence, no
//     kernel boundary; every stage output has its own buffer, so nothing is overwritten within a token;
//   * the attention stage is the cc_decode_step algorithm (decode_step.cu: eviction decision -> overwrite ->
//     masked attention -> GQA pooling -> score update, same bit-reproducible arithmetic) with the three cluster
//     exchanges replaced by flag-in-data exchanges between the CTAs of a KV head;
//   * with tensor parallelism the row-parallel stages (wo, w2) store their partial vectors into EVERY rank's
//     scratch (peer-mapped NVLink memory) and the consuming stage sums the `world` partial vectors: the all-reduce
//     of tp.py:134-160 costs one NVLink write latency and is hidden behind the weight stream like any other stage
//     hand-over.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/cc_b200.h"
#include "common.cuh"
#include "internal.h"

namespace cc {

#ifndef CC_TOKEN_WARPS
#define CC_TOKEN_WARPS 8
#endif
#ifndef CC_TOKEN_PRODUCER
#define CC_TOKEN_PRODUCER 1
#endif
constexpr int kTW = CC_TOKEN_WARPS;    // consumer warps per CTA (8 or 16)
constexpr int kTT = kTW * 32;          // consumer threads per CTA
constexpr bool kProducer = CC_TOKEN_PRODUCER != 0;   // one extra warp owns every ring's fetch cursor
// The producer warp comes with a whole warpgroup (setmaxnreg works on aligned groups of 4 warps): registers are
// dealt per scheduler, so a 9th warp alone would cap every thread at 168 registers; instead the kernel starts at
// 168 (384 threads), the producer warpgroup shrinks to kProducerRegs and the two consumer warpgroups grow to
// kConsumerRegs (128 * 40 + 256 * 232 = 64512 = 384 * 168).
constexpr int kLaunchThreads = kTT + (kProducer ? 128 : 0);
constexpr int kProducerRegs = 40, kConsumerRegs = 232;
// setmaxnreg moves registers inside the CTA's own launch allocation (384 x 168): what the producer warpgroup gives
// up must cover what the consumers take, or the consumers' setmaxnreg.inc waits forever (48 / 232 hangs the kernel)
static_assert(128 * (168 - kProducerRegs) >= 256 * (kConsumerRegs - 168), "setmaxnreg split exceeds the CTA's register allocation");
static_assert(!kProducer || kTW == 8, "the producer warpgroup layout assumes 8 consumer warps");
// CTA-wide barrier of the consumer warps (the producer warp never joins it)
__device__ __forceinline__ void cta_sync() {
    if (kProducer) asm volatile("bar.sync 1, %0;" :: "n"(kTT) : "memory");
    else __syncthreads();
}
constexpr uint32_t kStage = 65536 / kTW;   // bytes per ring stage: 8 KiB (8 warps) / 4 KiB (16 warps)
constexpr int kRedWarps = kTW < 8 ? kTW : 8;   // P.V partials are reduced through 8 shared-memory slots
constexpr int kTokMaxWorld = 8;
constexpr int kTokMaxG = 8;
constexpr uint64_t kNoKey64 = ~0ull;

enum : int { TOK_QKV = 0, TOK_ATTN = 1, TOK_GEMV = 2, TOK_GATE = 3, TOK_HEAD = 4 };
enum : int { IN_TOKEN = 0, IN_LL = 1 };
enum : int { RES_NONE = 0, RES_TOKEN = 1, RES_LL = 2 };

// One stage of the token program (device memory, read-only during the kernel).
struct TokOp {
    int32_t type;         // TOK_*
    int32_t K, N;         // matrix-vector stages: input length, output rows
    int32_t n_groups;     // row groups (a group = 2 output rows; the FFN gate reads 2 rows of w1 and 2 of w3)
    int32_t seg_len;      // elements of a row per ring stage (multiple of 256)
    int32_t n_seg;        // ceil(K / seg_len)
    int32_t rot;          // rotation of the CTA dealing (spreads the +1 row groups over CTAs from stage to stage)
    int32_t grp_base;     // n_groups / n_ctas (dealt on the host: no integer division in the kernel)
    int32_t in_kind;      // IN_TOKEN: embedding row of the input token; IN_LL: in_world partial vectors at in_off
    int32_t in_world;
    int32_t res_kind;     // residual added to the (summed) input: none / embedding row / flag-in-data vector at res_off
    int32_t publish_cta;  // the CTA that also writes the staged input vector to sum_off (-1: nobody)
    int32_t out_world;    // TOK_GEMV: ranks the partial output is stored to (1 = this rank only)
    int32_t attn;         // TOK_ATTN: index into the attention table
    float eps;
    int32_t grp_rem;      // n_groups % n_ctas
    uint64_t in_off, res_off, sum_off, out_off;   // byte offsets inside the scratch (identical on every rank)
    const __nv_bfloat16 *norm_w, *w, *w3, *bias;
};

struct TokAttn {
    CcCacheState st;
    int32_t c_shift;      // log2(CTAs per KV head)
    int32_t ch_shift;     // log2(rows per chunk)
    int32_t rot;
    int32_t lg_stride;    // floats per query head in the logit / probability array
    float scale;
    int32_t pad;
    uint64_t qkv_off, y_off, x1_off, x2_off, x3_off;
    int64_t* fill_idx;
    __nv_bfloat16* probs;
};

struct TokParams {
    const TokOp* ops;
    const TokAttn* attns;
    int32_t n_ops, world, rank, vocab;
    uint8_t* scratch;             // this rank's scratch
    const uint64_t* peers;        // device array [world]: scratch base address of every rank
    const int32_t* token_in;
    const int32_t* input_pos;
    int32_t* token_out;
    __nv_bfloat16* logits_out;
    __nv_bfloat16* probs_out;
    const __nv_bfloat16* emb;
    const __nv_bfloat16* freqs;
    int32_t dim, n_q_rows, n_kv_rows, head_dim;
    uint32_t xs_bytes;            // staging area at the start of shared memory (input vector / attention scratch)
    int32_t n_stages;             // ring stages per warp
    uint64_t rec1_off, rec2_off, rec3_off;  // logit arg-max / softmax-sum / probability arg-max records, one per CTA
    unsigned long long* trace;
};

// scratch header
struct TokHeader { uint32_t gen; uint32_t done; };

// diagnostics: trace[(stage * n_ctas + cta) * kTraceSlots + slot] = %globaltimer
constexpr int kTraceSlots = 12;
__device__ __forceinline__ void tstamp(const TokParams& P, int oi, int slot) {
    if (P.trace) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        P.trace[((size_t)oi * gridDim.x + blockIdx.x) * kTraceSlots + slot] = t;
    }
}

// The stage descriptors live in global memory; a cold read costs an L2 round trip (~0.5 us) and the tile cursors
// and the stages themselves read a dozen fields at every stage boundary.  A window of the next kOpWin stages is
// kept in shared memory: filled at kernel start, advanced by one entry at every stage entry.
constexpr int kOpWin = 16, kAttWin = 4;
struct DescWindow {
    TokOp op[kOpWin];
    TokAttn att[kAttWin];
    volatile int op_hi;    // stages < op_hi are in the window (slot = index % kOpWin)
    volatile int att_hi;   // attention records < att_hi are in the window (slot = index % kAttWin)
};
__device__ __forceinline__ const TokOp* op_at(const TokParams& P, const DescWindow* W, int oi) {
    if (oi < W->op_hi) { __threadfence_block(); return &W->op[oi % kOpWin]; }
    return P.ops + oi;
}
__device__ __forceinline__ const TokAttn* att_at(const TokParams& P, const DescWindow* W, int ai) {
    if (ai < W->att_hi) { __threadfence_block(); return &W->att[ai % kAttWin]; }
    return P.attns + ai;
}
// 16-byte chunk `c` of stage descriptor `oi` (and of its attention record) -> window; called by one warp
__device__ __forceinline__ void window_load(const TokParams& P, DescWindow* W, int oi, int lane, uint4* keep, int* keep_att) {
    // phase 1 (issue): returns the chunk in registers so that the caller can overlap the latency with a wait
    *keep_att = -1;
    if (oi >= P.n_ops) return;
    constexpr int nop = (int)(sizeof(TokOp) / 16), natt = (int)(sizeof(TokAttn) / 16);
    static_assert(sizeof(TokOp) % 16 == 0 && sizeof(TokAttn) % 16 == 0 && nop <= 32 && natt <= 32, "descriptor sizes");
    const TokOp* o = P.ops + oi;
    if (lane < nop) keep[0] = __ldg(reinterpret_cast<const uint4*>(o) + lane);
    if (o->type == TOK_ATTN) {
        *keep_att = o->attn;
        if (lane < natt) keep[1] = __ldg(reinterpret_cast<const uint4*>(P.attns + o->attn) + lane);
    }
}
__device__ __forceinline__ void window_store(const TokParams& P, DescWindow* W, int oi, int lane, const uint4* keep, int keep_att) {
    if (oi >= P.n_ops) return;
    constexpr int nop = (int)(sizeof(TokOp) / 16), natt = (int)(sizeof(TokAttn) / 16);
    if (lane < nop) reinterpret_cast<uint4*>(&W->op[oi % kOpWin])[lane] = keep[0];
    if (keep_att >= 0 && lane < natt) reinterpret_cast<uint4*>(&W->att[keep_att % kAttWin])[lane] = keep[1];
    __syncwarp();
    __threadfence_block();
    if (lane == 0) {
        if (keep_att >= 0) W->att_hi = keep_att + 1;
        W->op_hi = oi + 1;
    }
}

// ------------------------------------------------------------------ flag-in-data words
// u64 = {payload (low half), token tag (high half)}; naturally aligned 8-byte accesses are single-copy atomic, so a
// word whose tag is current carries valid data.  Words written by a peer GPU are accessed at system scope, words
// that only this GPU writes at gpu scope (both are served by L2; no L1).
__device__ __forceinline__ void ll_store(uint64_t* p, uint32_t data, uint32_t tag) {
    const uint64_t v = ((uint64_t)tag << 32) | data;
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ll_load(const uint64_t* p, bool sys = false) {
    uint64_t v;
    if (sys) asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    else asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void ll_load2(const uint64_t* p, uint64_t& a, uint64_t& b, bool sys = false) {
    if (sys) asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
    else asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ uint32_t ll_wait(const uint64_t* p, uint32_t tag, bool sys = false) {
    uint64_t v;
    do { v = ll_load(p, sys); } while ((uint32_t)(v >> 32) != tag);
    return (uint32_t)v;
}
// N groups of four consecutive words (each p[i] 16-byte aligned) = N x 8 bf16 values.  All 2N loads of a round are
// issued before any tag is looked at: one memory round trip when the data is there.
template <int N>
__device__ __forceinline__ void ll_wait4_n(const uint64_t* const (&p)[N], int n, uint32_t tag, uint4 (&o)[N], bool sys = false) {
    uint64_t v[4 * N];
    while (true) {
#pragma unroll
        for (int i = 0; i < N; ++i)
            if (i < n) {
                ll_load2(p[i], v[4 * i], v[4 * i + 1], sys);
                ll_load2(p[i] + 2, v[4 * i + 2], v[4 * i + 3], sys);
            }
        uint32_t bad = 0;
#pragma unroll
        for (int i = 0; i < N; ++i)
            if (i < n) {
#pragma unroll
                for (int j = 0; j < 4; ++j) bad |= (uint32_t)(v[4 * i + j] >> 32) ^ tag;
            }
        if (!bad) break;
    }
#pragma unroll
    for (int i = 0; i < N; ++i)
        o[i] = make_uint4((uint32_t)v[4 * i], (uint32_t)v[4 * i + 1], (uint32_t)v[4 * i + 2], (uint32_t)v[4 * i + 3]);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// ------------------------------------------------------------------ dealing of work to CTAs and warps
// matrix-vector stage: CTA c owns row groups [g0, g0 + n).  Inside the CTA the unit of work is a TILE = (group,
// segment) pair = one ring stage: tile t = j * n_seg + s of the CTA goes to warp t % kTW, so the warps of a CTA differ
// by at most one tile (dealing whole groups left warps with 1 or 2 groups of 7 tiles each in the w2 stage: the stage
// then ended with a few warps crawling through a latency-bound tail while the others sat on full rings).
__device__ __forceinline__ void deal_groups(const TokOp& op, int* g0, int* n) {
    const int nc = (int)gridDim.x;
    int c = (int)blockIdx.x + op.rot;
    if (c >= nc) c -= nc;
    *g0 = c * op.grp_base + min(c, op.grp_rem);
    *n = op.grp_base + (c < op.grp_rem ? 1 : 0);
}
// first tile of warp `w`: (j, s) with j * n_seg + s == w
__device__ __forceinline__ void first_tile(int w, int n_seg, int* j, int* s) {
    int jj = 0, ss = w;
    while (ss >= n_seg) { ss -= n_seg; ++jj; }
    *j = jj; *s = ss;
}
__device__ __forceinline__ void next_tile(int n_seg, int* j, int* s) {
    int jj = *j, ss = *s + kTW;
    while (ss >= n_seg) { ss -= n_seg; ++jj; }
    *j = jj; *s = ss;
}
// attention stage: CTA unit u < H * C works on KV head u / C as rank u % C of that head; chunk j of rank r is
// global chunk r + j * C; warp w owns rows [w * RW, (w + 1) * RW) of every chunk
struct AttnDeal {
    int active, h, rank, nlc;   // nlc: chunks of this rank
};
__device__ __forceinline__ AttnDeal deal_attn(const TokAttn& A) {
    AttnDeal d;
    const int nc = (int)gridDim.x;
    int u = (int)blockIdx.x + A.rot;
    if (u >= nc) u -= nc;
    const int C = 1 << A.c_shift;
    d.active = u < A.st.n_kv_heads * C;
    d.h = u >> A.c_shift;
    d.rank = u & (C - 1);
    const int n_chunks = (A.st.max_cache_length + (1 << A.ch_shift) - 1) >> A.ch_shift;
    d.nlc = (d.active && n_chunks > d.rank) ? (n_chunks - d.rank + C - 1) >> A.c_shift : 0;
    return d;
}
// first slot and number of valid rows of warp `w` in local chunk `jj`
__device__ __forceinline__ void attn_warp_rows(const TokAttn& A, int rank, int jj, int w, int* first, int* nvalid) {
    const int RW = (1 << A.ch_shift) / kTW;
    const int f = ((((jj << A.c_shift) + rank)) << A.ch_shift) + w * RW;
    *first = f;
    *nvalid = max(0, min(RW, A.st.max_cache_length - f));
}

// ------------------------------------------------------------------ per-warp tile sequence (the fetch cursor)
// Advancing inside a stage is a handful of adds (the cursor sits in lane 0's serial section of every ring release);
// only entering a new stage computes the dealing.
struct TileIter {
    int op;                 // current stage index (>= n_ops: exhausted)
    int type;
    int j, nj;              // matrix-vector: group index / groups of this CTA; attention: tile index / 2 * nlc
    int s, ns;              // segment index / count
    const uint8_t* base;    // matrix-vector: first row of the CTA's first group
    uint32_t seg_bytes, last_bytes;   // bytes of a full / of the last segment of a row
    uint32_t row_bytes;     // K * 2
    long long w3_delta;     // FFN gate: w3 - w1
    int rank, h;            // attention
    const TokAttn* a;
};

__device__ __forceinline__ bool iter_tile_valid(const TileIter& it, int warp) {
    if (it.type != TOK_ATTN) return true;
    const int nlc = it.nj >> 1;
    int first, nv;
    attn_warp_rows(*it.a, it.rank, it.j < nlc ? it.j : it.j - nlc, warp, &first, &nv);
    return nv > 0;
}
// position the cursor on the first tile of stage it.op or of the first later stage in which this warp has work
__device__ __forceinline__ void iter_enter(TileIter& it, const TokParams& P, const DescWindow* W, int warp) {
    for (; it.op < P.n_ops; ++it.op) {
        const TokOp* o = op_at(P, W, it.op);
        it.type = o->type;
        if (o->type == TOK_ATTN) {
            const TokAttn* a = att_at(P, W, o->attn);
            const AttnDeal d = deal_attn(*a);
            if (d.nlc == 0) continue;
            it.a = a; it.rank = d.rank; it.h = d.h;
            it.nj = 2 * d.nlc; it.s = 0; it.ns = 1;
            for (it.j = 0; it.j < it.nj && !iter_tile_valid(it, warp); ++it.j) {}
            if (it.j < it.nj) return;
        } else {
            int g0, n;
            deal_groups(*o, &g0, &n);
            first_tile(warp, o->n_seg, &it.j, &it.s);
            if (it.j >= n) continue;
            const int K = o->K;
            it.nj = n; it.ns = o->n_seg;
            it.row_bytes = (uint32_t)K * 2;
            it.seg_bytes = (uint32_t)o->seg_len * 2;
            it.last_bytes = (uint32_t)(K - (o->n_seg - 1) * o->seg_len) * 2;
            it.base = reinterpret_cast<const uint8_t*>(o->w) + (size_t)(2 * g0) * it.row_bytes;
            it.w3_delta = o->type == TOK_GATE ? reinterpret_cast<const uint8_t*>(o->w3) - reinterpret_cast<const uint8_t*>(o->w) : 0;
            return;
        }
    }
}
__device__ __forceinline__ void iter_advance(TileIter& it, const TokParams& P, const DescWindow* W, int warp) {
    if (it.type != TOK_ATTN) {
        next_tile(it.ns, &it.j, &it.s);
        if (it.j < it.nj) return;
    } else {
        for (++it.j; it.j < it.nj && !iter_tile_valid(it, warp); ++it.j) {}
        if (it.j < it.nj) return;
    }
    ++it.op;
    iter_enter(it, P, W, warp);
}
__device__ __forceinline__ const uint8_t* iter_src(const TileIter& it) {
    return it.base + (size_t)(2 * it.j) * it.row_bytes + (size_t)it.s * it.seg_bytes;
}
// issue the TMA copies of the cursor's tile into a ring stage (one lane).  Weights and cache rows are read once
// per token: lowest L2 retention priority, so that the 15 GB stream does not evict the exchange words and the
// kernel's own instructions.
__device__ __forceinline__ void iter_issue(const TileIter& it, int warp, uint8_t* stage, uint64_t* bar, uint64_t pol) {
    if (it.type == TOK_ATTN) {
        const TokAttn& A = *it.a;
        const int nlc = it.nj >> 1;
        const bool is_v = it.j >= nlc;
        int first, nv;
        attn_warp_rows(A, it.rank, is_v ? it.j - nlc : it.j, warp, &first, &nv);
        const uint32_t row_bytes = (uint32_t)A.st.head_dim * 2;
        const uint8_t* base = reinterpret_cast<const uint8_t*>(is_v ? A.st.v_cache : A.st.k_cache);
        const uint32_t bytes = (uint32_t)nv * row_bytes;
        mbar_expect_tx(bar, bytes);
        tma_load_1d_hint(stage, base + ((size_t)it.h * A.st.max_cache_length + first) * row_bytes, bytes, bar, pol);
        return;
    }
    const uint32_t bytes = it.s + 1 == it.ns ? it.last_bytes : it.seg_bytes;
    const uint8_t* src = iter_src(it);
    if (it.type == TOK_GATE) {
        mbar_expect_tx(bar, 4 * bytes);
        tma_load_1d_hint(stage, src, bytes, bar, pol);
        tma_load_1d_hint(stage + bytes, src + it.row_bytes, bytes, bar, pol);
        tma_load_1d_hint(stage + 2 * bytes, src + it.w3_delta, bytes, bar, pol);
        tma_load_1d_hint(stage + 3 * bytes, src + it.w3_delta + it.row_bytes, bytes, bar, pol);
    } else {
        mbar_expect_tx(bar, 2 * bytes);
        tma_load_1d_hint(stage, src, bytes, bar, pol);
        tma_load_1d_hint(stage + bytes, src + it.row_bytes, bytes, bar, pol);
    }
}

// ------------------------------------------------------------------ the per-warp ring
// Stage s of warp w: a "full" mbarrier (armed with expect_tx by whoever issues the copies, completed by the TMA
// engine) and, with the producer warp, an "empty" mbarrier (one arrival by the consumer warp's lane 0).
struct Ring {
    uint8_t* stages;      // this warp's stages
    uint64_t* bars;       // this warp's full barriers
    uint64_t* empty;      // this warp's empty barriers (producer mode)
    int ns, cs;           // number of stages, stage the consumer reads next
    uint32_t phases;      // parity of the next completion of each stage's full barrier
    const DescWindow* win;
    uint64_t policy;      // L2 eviction priority of the streamed tiles (created once)
    TileIter it;          // fetch cursor (self-refill mode: lane 0's copy; producer mode: unused)
};
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ const uint8_t* ring_wait(Ring& r) {
    mbar_wait(&r.bars[r.cs], (r.phases >> r.cs) & 1u);
    return r.stages + (size_t)r.cs * kStage;
}
// the warp is done with the current stage.  Producer mode: hand it back (one mbarrier arrival; the producer warp
// re-arms it with the next tile of this warp's sequence).  Self-refill mode: lane 0 re-arms it itself.
__device__ __forceinline__ void ring_release(Ring& r, const TokParams& P, int warp, int lane) {
    __syncwarp();
    // (no proxy fence: the stage was only READ through the generic proxy; the warp barrier above orders those
    // reads before the arrival / the copy that overwrites it, as in any TMA producer / consumer ring)
    if (kProducer) {
        if (lane == 0) mbar_arrive(&r.empty[r.cs]);
    } else if (lane == 0 && r.it.op < P.n_ops) {
        iter_issue(r.it, warp, r.stages + (size_t)r.cs * kStage, &r.bars[r.cs], r.policy);
        iter_advance(r.it, P, r.win, warp);
    }
    r.phases ^= 1u << r.cs;
    r.cs = r.cs + 1 == r.ns ? 0 : r.cs + 1;
}
// The producer warp: lane w owns the fetch cursor of consumer warp w's ring.  All lanes poll their ring's next
// empty barrier with ONE non-blocking test per round (the lanes stay converged), the lanes whose stage is free issue
// its copies and advance their cursor.  The ~500-cycle serial section of a refill (expect_tx, 2-4 bulk copies, cursor
// arithmetic) is off the consumers' path, and eight refills cost one pass.
__device__ void producer_loop(const TokParams& P, const DescWindow* W, uint8_t* ring_base, uint64_t* full, uint64_t* empty, int lane) {
    if (lane >= kTW) return;
    const int ns = P.n_stages;
    uint8_t* stages = ring_base + (size_t)lane * ns * kStage;
    uint64_t* fb = full + lane * ns;
    uint64_t* eb = empty + lane * ns;
    const uint64_t policy = l2_policy_evict_first();
    TileIter it;
    it.op = 0;
    iter_enter(it, P, W, lane);
    int ps = 0;
    uint32_t par = 0xffffffffu;   // a fresh barrier passes a wait on the phase "before the first": every stage starts free
    // (a second cursor prefetching the tiles beyond the ring into L2 -- cp.async.bulk.prefetch.L2, 2 / 4 / 8 / 16 tiles
    // per warp whenever the ring is full -- was measured 0.4 / 0.7 / 1.8 / 3.9 % SLOWER than none: the bulk
    // prefetches queue in front of the demand copies in the SM's TMA unit; profiles/r2/notes.md)
    while (it.op < P.n_ops) {
        if (mbar_test(&eb[ps], (par >> ps) & 1u)) {
            iter_issue(it, lane, stages + (size_t)ps * kStage, &fb[ps], policy);
            iter_advance(it, P, W, lane);
            par ^= 1u << ps;
            ps = ps + 1 == ns ? 0 : ps + 1;
        }
    }
}

// ------------------------------------------------------------------ input staging of a matrix-vector stage
// x = [residual +] sum over ranks of the predecessor's partial vectors  ->  shared memory (bf16), optionally
// RMS-normalised the way the reference's RMSNorm does it (model.py:452-457).
__device__ void stage_input(const TokParams& P, const TokOp& op, __nv_bfloat16* xs, float* scratch, uint32_t tag) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = op.K;
    const int token = (op.in_kind == IN_TOKEN || op.res_kind == RES_TOKEN) ? P.token_in[0] : 0;
    if (op.in_kind == IN_TOKEN) {
        const __nv_bfloat16* src = P.emb + (size_t)token * K;
        for (int i = tid * 8; i < K; i += kTT * 8)
            *reinterpret_cast<uint4*>(xs + i) = __ldg(reinterpret_cast<const uint4*>(src + i));
    } else {
        const int nw = K >> 1, W = op.in_world;
        const bool sys = W > 1;   // partial vectors written by peer GPUs
        const uint64_t* in = reinterpret_cast<const uint64_t*>(P.scratch + op.in_off);
        // Gate: one warp samples 32 words per source (with a back-off) before the CTA reads the whole vector, so a
        // CTA that is early does not hammer L2 with 256 spinning threads while the weight stream needs the bandwidth.
        if (warp == 0) {
            for (int sidx = 0; sidx < W; ++sidx) {
                const uint64_t* p = in + (size_t)sidx * nw + (size_t)lane * (nw >> 5);
                while ((uint32_t)(ll_load(p, sys) >> 32) != tag) __nanosleep(32);
            }
        }
        cta_sync();
        const uint64_t* res = op.res_kind == RES_LL ? reinterpret_cast<const uint64_t*>(P.scratch + op.res_off) : nullptr;
        const __nv_bfloat16* rest = op.res_kind == RES_TOKEN ? P.emb + (size_t)token * K : nullptr;
        uint64_t* sum_out = op.publish_cta == (int)blockIdx.x ? reinterpret_cast<uint64_t*>(P.scratch + op.sum_off) : nullptr;
        // x = [residual +] bf16(sum): h = x + attention(...) / out = h + feed_forward(...) are bf16 + bf16 -> bf16
        // (model.py:317, 326)
        auto finish = [&](int i, float (&acc)[8], const uint4* resv, bool summed) {
            if (op.res_kind != RES_NONE) {
                float f[8];
                unpack8(resv ? *resv : __ldg(reinterpret_cast<const uint4*>(rest + i)), f);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = __fadd_rn(f[j], bf16_round(acc[j]));
            } else if (summed) {
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = bf16_round(acc[j]);
            }
            const uint4 o = make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]),
                                       pack_bf16x2(acc[6], acc[7]));
            *reinterpret_cast<uint4*>(xs + i) = o;
            if (sum_out) {
                ll_store(sum_out + (i >> 1) + 0, o.x, tag);
                ll_store(sum_out + (i >> 1) + 1, o.y, tag);
                ll_store(sum_out + (i >> 1) + 2, o.z, tag);
                ll_store(sum_out + (i >> 1) + 3, o.w, tag);
            }
        };
        if (W == 1 && res) {
            // one source + residual vector: two 8-element groups of this thread per memory round trip
            for (int i0 = tid * 8; i0 < K; i0 += 2 * kTT * 8) {
                const int i1 = i0 + kTT * 8, i1c = i1 < K ? i1 : i0;
                const uint64_t* p[4] = {in + (i0 >> 1), res + (i0 >> 1), in + (i1c >> 1), res + (i1c >> 1)};
                uint4 o[4];
                ll_wait4_n<4>(p, i1 < K ? 4 : 2, tag, o);
                float acc[8];
                unpack8(o[0], acc);
                finish(i0, acc, &o[1], false);
                if (i1 < K) {
                    unpack8(o[2], acc);
                    finish(i1, acc, &o[3], false);
                }
            }
        } else if (W == 1) {
            // one source: four 8-element groups of this thread per memory round trip
            for (int i0 = tid * 8; i0 < K; i0 += 4 * kTT * 8) {
                const uint64_t* p[4];
                int n = 0;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + u * kTT * 8;
                    p[u] = in + ((i < K ? i : i0) >> 1);
                    n += i < K ? 1 : 0;
                }
                uint4 o[4];
                ll_wait4_n<4>(p, n, tag, o);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + u * kTT * 8;
                    if (i < K) {
                        if (op.res_kind == RES_NONE && !sum_out) {
                            *reinterpret_cast<uint4*>(xs + i) = o[u];
                        } else {
                            float acc[8];
                            unpack8(o[u], acc);
                            finish(i, acc, nullptr, false);
                        }
                    }
                }
            }
        } else {
            // tensor parallel: residual + the first four partial vectors of an 8-element group in one round trip,
            // the other ranks' in a second one; summed in rank order (the same bits on every rank)
            for (int i = tid * 8; i < K; i += kTT * 8) {
                float acc[8], f[8];
                uint4 o[5];
                {
                    const uint64_t* p[5];
                    const int n = min(4, W);
                    p[0] = res ? res + (i >> 1) : in + (i >> 1);
#pragma unroll
                    for (int u = 0; u < 4; ++u) p[1 + u] = in + (size_t)(u < n ? u : 0) * nw + (i >> 1);
                    ll_wait4_n<5>(p, 1 + n, tag, o, true);
                    unpack8(o[1], acc);
#pragma unroll
                    for (int u = 1; u < 4; ++u)
                        if (u < n) {
                            unpack8(o[1 + u], f);
#pragma unroll
                            for (int j = 0; j < 8; ++j) acc[j] += f[j];
                        }
                }
                if (W > 4) {
                    const uint64_t* p[4];
                    const int n = W - 4;
                    uint4 o2[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) p[u] = in + (size_t)(4 + (u < n ? u : 0)) * nw + (i >> 1);
                    ll_wait4_n<4>(p, n, tag, o2, true);
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (u < n) {
                            unpack8(o2[u], f);
#pragma unroll
                            for (int j = 0; j < 8; ++j) acc[j] += f[j];
                        }
                }
                finish(i, acc, res ? &o[0] : nullptr, true);
            }
        }
    }
    cta_sync();
    if (op.norm_w == nullptr) return;
    float ss = 0.f;
    for (int i = tid * 8; i < K; i += kTT * 8) {
        float f[8];
        unpack8(*reinterpret_cast<const uint4*>(xs + i), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) ss = fmaf(f[j], f[j], ss);
    }
    ss = warp_sum_f(ss);
    if (lane == 0) scratch[warp] = ss;
    cta_sync();
    float tot = 0.f;
    for (int w = 0; w < kTW; ++w) tot += scratch[w];
    const float rstd = rsqrtf(tot / (float)K + op.eps);
    auto normalise = [&](int i, const uint4& wq) {
        float f[8], wv[8];
        unpack8(*reinterpret_cast<const uint4*>(xs + i), f);
        unpack8(wq, wv);
        *reinterpret_cast<uint4*>(xs + i) =
            make_uint4(pack_bf16x2(__fmul_rn(bf16_round(__fmul_rn(f[0], rstd)), wv[0]), __fmul_rn(bf16_round(__fmul_rn(f[1], rstd)), wv[1])),
                       pack_bf16x2(__fmul_rn(bf16_round(__fmul_rn(f[2], rstd)), wv[2]), __fmul_rn(bf16_round(__fmul_rn(f[3], rstd)), wv[3])),
                       pack_bf16x2(__fmul_rn(bf16_round(__fmul_rn(f[4], rstd)), wv[4]), __fmul_rn(bf16_round(__fmul_rn(f[5], rstd)), wv[5])),
                       pack_bf16x2(__fmul_rn(bf16_round(__fmul_rn(f[6], rstd)), wv[6]), __fmul_rn(bf16_round(__fmul_rn(f[7], rstd)), wv[7])));
    };
    // (requesting the norm weights before the hand-over wait was measured: no change, 64 bytes of spills)
    for (int i = tid * 8; i < K; i += kTT * 8) normalise(i, __ldg(reinterpret_cast<const uint4*>(op.norm_w + i)));
    cta_sync();
}

// One ring stage against the staged input on the tensor cores (mma.sync m16n8k16, bf16 in, fp32 accumulate).
// The stage holds R row segments of `len` elements (R = 2, or 4 for the FFN gate: R * len <= 4096).  Every row is
// read as S = 16 / R STREAMS (8 or 4); stream p of a row takes the 32-element blocks p, p + S, p + 2S, ... of the
// segment.  The 16 rows of the MMA's A operand are 16 (row, stream) pairs -- rows 0-7 of A: the first R / 2 stage rows
// x S streams, rows 8-15: the other R / 2 stage rows x the same streams -- and column n of B is the input's stream
// n % S, so D[m][m % 8] accumulates the dot product of A row m's (row, stream) pair: the diagonal of the upper and of
// the lower 8 x 8 half.  Thread (gid = lane / 4, tig = lane % 4) loads 8 consecutive elements (one LDS.128) of its
// stream from the upper row, from the lower row and from the input per block and issues two MMAs (the 16 k-slots
// of an MMA can hold any 16 elements as long as A and B agree): 256 weight elements per MMA.  In every round the
// lanes of a quarter-warp read one contiguous 128-byte run of a row and the warp reads S * 64 contiguous bytes of
// the input: conflict free without padding, 12 (R = 2) or 10 (R = 4) shared-memory wavefronts per KiB of weights.
// (Round 2, first version: 8 A rows duplicated into 16 = half the weights per MMA and 16 wavefronts per KiB; a
// hand-written FFMA loop costs 10x the instructions.  Both left the warps barely faster than the HBM stream.)
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
template <int S_LOG2>   // log2(streams per row): 3 for two-row stages, 2 for the four-row FFN gate stage
__device__ __forceinline__ void consume_tile(const uint8_t* stage, const __nv_bfloat16* xseg, int len, float (&c)[8][4], int lane) {
    constexpr int S = 1 << S_LOG2;
    const int gid = lane >> 2, tig = lane & 3;
    const int nblk = len >> (5 + S_LOG2);   // 32-element blocks per stream
    const uint8_t* wp = stage + (size_t)(gid >> S_LOG2) * (len * 2) + (gid & (S - 1)) * 64 + tig * 16;
    const uint32_t lower = (uint32_t)(8 >> S_LOG2) * (uint32_t)(len * 2);   // A row gid + 8: the stage row R / 2 further on
    const uint8_t* xp = reinterpret_cast<const uint8_t*>(xseg) + (gid & (S - 1)) * 64 + tig * 16;
    // four blocks per round on eight independent accumulators
    int i = 0;
    for (; i + 4 <= nblk; i += 4) {
        uint4 wu[4], wl[4], x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            wu[u] = *reinterpret_cast<const uint4*>(wp + (i + u) * (64 * S));
            wl[u] = *reinterpret_cast<const uint4*>(wp + lower + (i + u) * (64 * S));
            x[u] = *reinterpret_cast<const uint4*>(xp + (i + u) * (64 * S));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            mma_bf16_16816(c[2 * u], wu[u].x, wl[u].x, wu[u].y, wl[u].y, x[u].x, x[u].y);
            mma_bf16_16816(c[2 * u + 1], wu[u].z, wl[u].z, wu[u].w, wl[u].w, x[u].z, x[u].w);
        }
    }
    for (; i < nblk; ++i) {
        const uint4 wu = *reinterpret_cast<const uint4*>(wp + i * (64 * S));
        const uint4 wl = *reinterpret_cast<const uint4*>(wp + lower + i * (64 * S));
        const uint4 x = *reinterpret_cast<const uint4*>(xp + i * (64 * S));
        mma_bf16_16816(c[0], wu.x, wl.x, wu.y, wl.y, x.x, x.y);
        mma_bf16_16816(c[1], wu.z, wl.z, wu.w, wl.w, x.z, x.w);
    }
}
// after the last stage of a row group: D[gid][gid] (upper rows) and D[gid + 8][gid] (lower rows) live in the thread
// with tig == gid / 2, elements gid & 1 and 2 + (gid & 1); every other thread contributes zero to the butterflies
// that follow.  Fixed summation order: deterministic.
__device__ __forceinline__ void gather_streams(const float (&c)[8][4], int lane, float* upper, float* lower) {
    const int gid = lane >> 2, tig = lane & 3;
    const int e = gid & 1;
    float u = e ? c[0][1] : c[0][0], l = e ? c[0][3] : c[0][2];
#pragma unroll
    for (int k = 1; k < 8; ++k) {
        u += e ? c[k][1] : c[k][0];
        l += e ? c[k][3] : c[k][2];
    }
    const bool own = tig == (gid >> 1);
    *upper = own ? u : 0.f;
    *lower = own ? l : 0.f;
}

// deferred writes of cross-head cache state (see attn_stage): performed by thread 0 once every head has read it
struct Pending {
    int32_t active;
    int32_t pos_val;
    int32_t* pos_ptr;
    uint16_t* kn_ptr;
    uint32_t kn_val;
};

struct HeadBest { float v; int idx; };

// ------------------------------------------------------------------ a matrix-vector stage
__device__ void gemv_stage(const TokParams& P, DescWindow* W, const TokOp& op, int oi, uint8_t* smem, Ring& ring, uint32_t tag,
                           float* scratch, Pending* pend, HeadBest* best) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(smem);
    cta_sync();   // every warp is done with the staging area of the previous stage
    if (tid == 0) tstamp(P, oi, 0);
    if (warp == 1) {   // advance the descriptor window (this warp only waits for warp 0's poll otherwise)
        uint4 keep[2];
        int ka;
        window_load(P, W, oi + kOpWin - 1, lane, keep, &ka);
        window_store(P, W, oi + kOpWin - 1, lane, keep, ka);
    }
    stage_input(P, op, xs, scratch, tag);
    if (tid == 0) tstamp(P, oi, 1);
    // cross-head cache state of the attention stage just before this one: its input being complete means every
    // head's CTAs have finished reading the old state
    if (tid == 0 && pend->active) {
        if (pend->pos_ptr) *pend->pos_ptr = pend->pos_val;
        if (pend->kn_ptr) *pend->kn_ptr = (uint16_t)pend->kn_val;
        pend->active = 0;
    }
    int g0, nloc;
    deal_groups(op, &g0, &nloc);
    const int type = op.type, K = op.K, n_seg = op.n_seg;
    // partial sums of every tile of this CTA, [group][segment] x {row sums}: behind the staged input
    float4* part = reinterpret_cast<float4*>(smem + (((size_t)K * 2 + 15) & ~(size_t)15));
#ifdef CC_TOKEN_MICROTRACE
    int dbg_n = 0;
    const unsigned long long dbg_t0 = clock64();
#endif
    {
        int j, sg;
        first_tile(warp, n_seg, &j, &sg);
        while (j < nloc) {
            const int off = sg * op.seg_len, len = min(op.seg_len, K - off);
            float c[8][4];
#pragma unroll
            for (int k = 0; k < 8; ++k) c[k][0] = c[k][1] = c[k][2] = c[k][3] = 0.f;
#ifdef CC_TOKEN_MICROTRACE
            unsigned long long* dbg = (P.trace && blockIdx.x == 0 && lane == 0 && dbg_n < 8)
                ? P.trace + (size_t)P.n_ops * gridDim.x * kTraceSlots + ((size_t)oi * kTW + warp) * 32 + 1 + dbg_n * 3 : nullptr;
            if (dbg) dbg[-1 - dbg_n * 3] = dbg_t0;
#endif
            const uint8_t* st = ring_wait(ring);
#ifdef CC_TOKEN_MICROTRACE
            if (dbg) dbg[0] = clock64();
#endif
            if (type == TOK_GATE) consume_tile<2>(st, xs + off, len, c, lane);
            else consume_tile<3>(st, xs + off, len, c, lane);
#ifdef CC_TOKEN_MICROTRACE
            if (dbg) dbg[1] = clock64();
#endif
            ring_release(ring, P, warp, lane);
#ifdef CC_TOKEN_MICROTRACE
            if (dbg) { dbg[2] = clock64(); ++dbg_n; }
#endif
            float up, lo;
            gather_streams(c, lane, &up, &lo);
#pragma unroll
            for (int d = 1; d <= 8; d <<= 1) {
                up += __shfl_xor_sync(0xffffffffu, up, d);
                lo += __shfl_xor_sync(0xffffffffu, lo, d);
            }
            // gate: upper = w1 rows r0 (lanes 0-15) and r0 + 1 (lanes 16-31), lower = the same rows of w3;
            // otherwise: upper = row r0, lower = row r0 + 1, each split over the two half-warps
            const float up_hi = __shfl_sync(0xffffffffu, up, 16), lo_hi = __shfl_sync(0xffffffffu, lo, 16);
            if (lane == 0)
                part[j * n_seg + sg] = type == TOK_GATE ? make_float4(up, up_hi, lo, lo_hi) : make_float4(up + up_hi, lo + lo_hi, 0.f, 0.f);
            next_tile(n_seg, &j, &sg);
        }
    }
    if (lane == 0) tstamp(P, oi, 2 + warp);
    cta_sync();
    // ---- epilogue: one thread per row group; segment sums in segment order (deterministic, independent of the
    // warp that computed a tile)
    const int pos = type == TOK_QKV ? P.input_pos[0] : 0;
    uint64_t* out = reinterpret_cast<uint64_t*>(P.scratch + op.out_off);
    for (int j = tid; j < nloc; j += kTT) {
        const int g = g0 + j, r0 = 2 * g;
        float4 a4 = part[j * n_seg];
        for (int sg = 1; sg < n_seg; ++sg) {
            const float4 b4 = part[j * n_seg + sg];
            a4.x += b4.x; a4.y += b4.y; a4.z += b4.z; a4.w += b4.w;
        }
        if (type == TOK_QKV) {
            // nn.Linear (+bias) output is a bf16 tensor; q and k rows rotate in fp32 (model.py:507-519), v rows do not
            float e0 = 0.f, e1 = 0.f;
            if (op.bias) { e0 = __bfloat162float(op.bias[r0]); e1 = __bfloat162float(op.bias[r0 + 1]); }
            float a = bf16_round(a4.x + e0), b = bf16_round(a4.y + e1);
            if (r0 < P.n_q_rows + P.n_kv_rows) {
                const int half = P.head_dim >> 1, i = (r0 & (P.head_dim - 1)) >> 1;
                const float e2 = __bfloat162float(P.freqs[((size_t)pos * half + i) * 2]);
                const float e3 = __bfloat162float(P.freqs[((size_t)pos * half + i) * 2 + 1]);
                const float ra = __fsub_rn(__fmul_rn(a, e2), __fmul_rn(b, e3));
                const float rb = __fadd_rn(__fmul_rn(b, e2), __fmul_rn(a, e3));
                a = ra; b = rb;
            }
            ll_store(out + g, pack_bf16x2(a, b), tag);
        } else if (type == TOK_GEMV) {
            // row-parallel partial outputs go to every rank (consecutive threads = consecutive words: whole lines)
            const uint32_t d = pack_bf16x2(bf16_round(a4.x), bf16_round(a4.y));
            const size_t word = (size_t)P.rank * (op.N >> 1) + g;
            for (int pr = 0; pr < op.out_world; ++pr)
                ll_store(reinterpret_cast<uint64_t*>(P.peers[pr] + op.out_off) + word, d, tag);
        } else if (type == TOK_GATE) {
            float o[2];
            const float w1v[2] = {a4.x, a4.y}, w3v[2] = {a4.z, a4.w};
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const float gt = bf16_round(w1v[q]);                      // w1(x), bf16 tensor
                const float sg = bf16_round(gt / (1.0f + expf(-gt)));     // F.silu in fp32, bf16 out
                o[q] = __fmul_rn(sg, bf16_round(w3v[q]));                 // * w3(x)
            }
            ll_store(out + g, pack_bf16x2(o[0], o[1]), tag);
        } else {  // TOK_HEAD: logits (model.py:291), running arg-max with torch.argmax's first-index rule
            const float a = bf16_round(a4.x), b = bf16_round(a4.y);
            if (P.logits_out) *reinterpret_cast<uint32_t*>(P.logits_out + r0) = pack_bf16x2(a, b);
            if (a > best->v || (a == best->v && r0 < best->idx)) { best->v = a; best->idx = r0; }
            if (b > best->v || (b == best->v && r0 + 1 < best->idx)) { best->v = b; best->idx = r0 + 1; }
        }
    }
}

// ------------------------------------------------------------------ the attention stage
// The cc_decode_step algorithm (decode_step.cu) for policies recent_global / l2 / heavy_hitter (full history), with
// C CTAs per KV head exchanging {candidate, row max}, {softmax sums} and {P.V partials} through flag-in-data words.
// attention scratch inside the staging area, in 4-byte words / bytes as noted
struct AttnSmem {
    float* lg; float* red; uint32_t* rec; uint32_t* sums; float* yx; unsigned long long* ssum; uint8_t* maskb; float* misc;
};
__host__ __device__ inline uint32_t attn_smem_bytes(int D, int G, int C, int stride, int local_rows) {
    uint32_t off = 0;
    off += (uint32_t)G * stride * 4;            // lg
    off += (uint32_t)kRedWarps * G * D * 4;     // red
    off += (uint32_t)C * (2 + 2 * G) * 4;       // rec
    off += (uint32_t)C * 2 * G * 4;             // sums (lo, hi)
    off += (uint32_t)G * D * 4;                 // yx
    off += kTokMaxG * 8;                        // ssum
    off += ((uint32_t)local_rows + 15) / 16 * 16;  // maskb
    off += 128 * 4;                             // misc
    return off;
}
enum : int { kTmNewLg = 0, kTmMyRec = 16, kTmWarpKey = 48, kTmWarpMax = 80, kTmWords = 128 };

template <int D, int G>
__device__ void attn_stage(const TokParams& P, DescWindow* W, const TokAttn& A, int oi, uint8_t* smem, Ring& ring, uint32_t tag, Pending* pend) {
    constexpr int LPR = D / 8;           // lanes per cache row
    constexpr int RPI = 32 / LPR;        // rows per warp iteration
    constexpr uint32_t kRowBytes = D * 2;
    constexpr int NWR = 2 + 2 * G;       // words of an X1 record: key lo, key hi, lmax[G], candidate logits[G]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int li = lane % LPR, sub = lane / LPR;
    const AttnDeal dl = deal_attn(A);
    cta_sync();   // the staging area of the previous stage is free
    if (tid == 0) tstamp(P, oi, 0);
    if (warp == 1) {   // advance the descriptor window
        uint4 keep[2];
        int ka;
        window_load(P, W, oi + kOpWin - 1, lane, keep, &ka);
        window_store(P, W, oi + kOpWin - 1, lane, keep, ka);
    }
    if (!dl.active) return;
    const CcCacheState& st = A.st;
    const int csh = A.c_shift, C = 1 << csh, chs = A.ch_shift, CH = 1 << chs, RW = CH / kTW;
    const int h = dl.h, rank = dl.rank, nlc = dl.nlc;
    const int H = st.n_kv_heads, L = st.max_cache_length;
    const int policy = st.policy;
    const bool naive = st.attn_mode == CC_ATTN_NAIVE_BF16;
    const int local_rows = nlc << chs, stride = A.lg_stride;
    auto slot_of = [&](int lr) { return ((((lr >> chs) << csh) + rank) << chs) + (lr & (CH - 1)); };

    AttnSmem sm;
    {
        uint8_t* p = smem;
        sm.lg = reinterpret_cast<float*>(p);    p += (size_t)G * stride * 4;
        sm.red = reinterpret_cast<float*>(p);   p += (size_t)kRedWarps * G * D * 4;
        sm.rec = reinterpret_cast<uint32_t*>(p); p += (size_t)C * NWR * 4;
        sm.sums = reinterpret_cast<uint32_t*>(p); p += (size_t)C * 2 * G * 4;
        sm.yx = reinterpret_cast<float*>(p);    p += (size_t)G * D * 4;
        sm.ssum = reinterpret_cast<unsigned long long*>(p); p += kTokMaxG * 8;
        sm.maskb = p;                           p += ((size_t)local_rows + 15) / 16 * 16;
        sm.misc = reinterpret_cast<float*>(p);
    }
    float* lg = sm.lg;
    float* misc = sm.misc;
    uint32_t* miscu = reinterpret_cast<uint32_t*>(misc);

    uint64_t* x1 = reinterpret_cast<uint64_t*>(P.scratch + A.x1_off) + (size_t)h * C * NWR;
    uint64_t* x2 = reinterpret_cast<uint64_t*>(P.scratch + A.x2_off) + (size_t)h * C * 2 * G;
    uint64_t* x3 = reinterpret_cast<uint64_t*>(P.scratch + A.x3_off) + (size_t)h * C * G * D;
    const uint64_t* qkv = reinterpret_cast<const uint64_t*>(P.scratch + A.qkv_off);

    // ---- A. everything that does not need this token's q / k / v: mask bytes, arg-min candidate
    const int ipos = P.input_pos[0];
    const int pos_row = st.pos_heads == 1 ? 0 : h;
    const int32_t* posp = st.pos + (size_t)pos_row * L;
    for (int lr = tid; lr < local_rows; lr += kTT) {
        const int s = slot_of(lr);
        sm.maskb[lr] = s < L ? st.mask[(size_t)h * L + s] : 0;
    }
    if (tid < kTokMaxG) sm.ssum[tid] = 0ull;
    float l2_max = 0.f;
    if (policy == CC_POLICY_L2) {
        // key_norm.max() over every head and slot (cache.py:602): 2 * H * L bytes from L2, every CTA for itself
        const uint16_t* kn = reinterpret_cast<const uint16_t*>(st.key_norm);
        float m = -CUDART_INF_F;
        for (int i = tid; i < H * L; i += kTT) m = fmaxf(m, bf16_bits_to_float(kn[i]));
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
        if (lane == 0) misc[kTmWarpMax + warp] = m;
        cta_sync();
        l2_max = misc[kTmWarpMax];
        for (int w = 1; w < kTW; ++w) l2_max = fmaxf(l2_max, misc[kTmWarpMax + w]);
    }
    uint64_t ckey = kNoKey64;
    if (policy == CC_POLICY_HEAVY_HITTER) {
        const double* num = reinterpret_cast<const double*>(st.attn_history_num) + (size_t)h * L;
        const int32_t* den = st.attn_history_denom + (size_t)h * L;
        for (int lr = tid; lr < local_rows; lr += kTT) {
            const int s = slot_of(lr);
            if (s >= L) continue;
            const int ps = posp[s];
            // cache.py:727-749: mean accumulated attention; globals & recent -> 1.0; unfilled -> 0.0
            float avg = __fdiv_rn(__double2float_rn(num[s]), (float)max(den[s], 1));
            if (ps < st.global_tokens || ps >= ipos - st.recent_window) avg = 1.0f;
            if (ps == -1) avg = 0.0f;
            const uint64_t k = ((uint64_t)float_order_key(avg) << 32) | (uint32_t)s;
            ckey = k < ckey ? k : ckey;
        }
    } else if (policy == CC_POLICY_L2) {
        const uint16_t* kn = reinterpret_cast<const uint16_t*>(st.key_norm) + (size_t)h * L;
        for (int lr = tid; lr < local_rows; lr += kTT) {
            const int s = slot_of(lr);
            if (s >= L) continue;
            const int ps = posp[s];
            // cache.py:597-605 then 366-379: bf16(max - norm); recent -> +inf; first g slots -> +inf; unfilled -> -inf
            float sc = bf16_round(__fsub_rn(l2_max, bf16_bits_to_float(kn[s])));
            if (ps >= ipos - st.recent_window) sc = CUDART_INF_F;
            if (s < st.global_tokens) sc = CUDART_INF_F;
            if (ps == -1) sc = -CUDART_INF_F;
            const uint64_t k = ((uint64_t)float_order_key(sc) << 32) | (uint32_t)s;
            ckey = k < ckey ? k : ckey;
        }
    } else {  // CC_POLICY_RECENT_GLOBAL, cache.py:552-556: oldest position among slots >= g (unfilled = -1 first)
        for (int lr = tid; lr < local_rows; lr += kTT) {
            const int s = slot_of(lr);
            if (s >= L || s < st.global_tokens) continue;
            const uint64_t k = ((uint64_t)((uint32_t)posp[s] ^ 0x80000000u) << 32) | (uint32_t)s;
            ckey = k < ckey ? k : ckey;
        }
    }
    ckey = warp_min_u64(ckey);
    if (lane == 0) reinterpret_cast<unsigned long long*>(miscu + kTmWarpKey)[warp] = ckey;
    cta_sync();
    ckey = lane < kTW ? reinterpret_cast<unsigned long long*>(miscu + kTmWarpKey)[lane] : kNoKey64;
    ckey = warp_min_u64(ckey);
    int cand_local = -1;
    if (ckey != kNoKey64) {
        const int s = (int)(uint32_t)(ckey & 0xffffffffu);
        cand_local = (((s >> chs) >> csh) << chs) + (s & (CH - 1));
    }

    // ---- B. this token's query rows of the KV group, the new key and value slices (flag-in-data words of the
    // projection stage): 4 words per 8 elements
    float q[G][8];
    uint4 knew8, vnew8;
    {
        // q rows in rounds of up to four groups (registers), then k and v: 1-2 memory round trips
        constexpr int QB = G <= 4 ? G : 4;
#pragma unroll
        for (int g0 = 0; g0 < G; g0 += QB) {
            const uint64_t* p[QB];
            uint4 o[QB];
#pragma unroll
            for (int u = 0; u < QB; ++u) p[u] = qkv + (((size_t)(h * G + min(g0 + u, G - 1)) * D + li * 8) >> 1);
            ll_wait4_n<QB>(p, min(QB, G - g0), tag, o);
#pragma unroll
            for (int u = 0; u < QB; ++u)
                if (g0 + u < G) unpack8(o[u], q[g0 + u]);
        }
        const uint64_t* p[2] = {qkv + (((size_t)P.n_q_rows + (size_t)h * D + li * 8) >> 1),
                                qkv + (((size_t)P.n_q_rows + P.n_kv_rows + (size_t)h * D + li * 8) >> 1)};
        uint4 o[2];
        ll_wait4_n<2>(p, 2, tag, o);
        knew8 = o[0];
        vnew8 = o[1];
    }
    if (tid == 0) tstamp(P, oi, 1);
    if (warp == kTW - 1) {   // logits of the new token (it attends to itself, model.py:392 then 402)
        float kf[8], part[G];
        unpack8(knew8, kf);
#pragma unroll
        for (int g = 0; g < G; ++g) {
            float acc = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) acc = __fmaf_rn(q[g][i], kf[i], acc);
            part[g] = acc;
        }
        int gsel;
        const float dot = group_reduce_multi<LPR, G>(part, li, &gsel);
        const float v = naive ? bf16_round(__fmul_rn(bf16_round(dot), A.scale)) : __fmul_rn(dot, A.scale);
        if (sub == 0 && gsel < G) misc[kTmNewLg + gsel] = v;
    }
    // ---- logits of this warp's rows of every local chunk (canonical dot tree), K tiles from the ring
    for (int jj = 0; jj < nlc; ++jj) {
        int first, nv;
        attn_warp_rows(A, rank, jj, warp, &first, &nv);
        const uint8_t* kt = nv > 0 ? ring_wait(ring) : nullptr;
        const int lr0 = (jj << chs) + warp * RW;
        for (int rb = 0; rb < RW; rb += RPI) {   // warp-uniform trip count: the dot tree shuffles over all lanes
            const int r = rb + sub;
            const uint4 kv = r < nv ? *reinterpret_cast<const uint4*>(kt + r * kRowBytes + li * 16) : make_uint4(0, 0, 0, 0);
            float kf[8], part[G];
            unpack8(kv, kf);
#pragma unroll
            for (int g = 0; g < G; ++g) {
                float acc = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) acc = __fmaf_rn(q[g][i], kf[i], acc);
                part[g] = acc;
            }
            int gsel;
            const float dot = group_reduce_multi<LPR, G>(part, li, &gsel);
            // naive path rounds q@k^T to bf16, then * scale to bf16 (attention_utils.py:37)
            const float v = naive ? bf16_round(__fmul_rn(bf16_round(dot), A.scale)) : __fmul_rn(dot, A.scale);
            const int lr = lr0 + r;
            if (gsel < G && r < RW) lg[gsel * stride + lr] = (r < nv && sm.maskb[lr]) ? v : -CUDART_INF_F;
        }
        if (nv > 0) ring_release(ring, P, warp, lane);
    }
    cta_sync();
    if (tid == 0) tstamp(P, oi, 2);
    // ---- row max without the local candidate, and the candidate's own logits -> this CTA's X1 record
    for (int g = warp; g < G; g += kTW) {
        float m = -CUDART_INF_F;
        for (int lr = lane; lr < local_rows; lr += 32)
            if (lr != cand_local) m = fmaxf(m, lg[g * stride + lr]);
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
        if (lane == 0) {
            misc[kTmMyRec + 2 + g] = m;
            misc[kTmMyRec + 2 + G + g] = cand_local >= 0 ? lg[g * stride + cand_local] : -CUDART_INF_F;
        }
    }
    if (tid == 0) {
        miscu[kTmMyRec + 0] = (uint32_t)(ckey & 0xffffffffu);
        miscu[kTmMyRec + 1] = (uint32_t)(ckey >> 32);
    }
    cta_sync();
    // ---- X1: every CTA of the head learns every candidate and every partial max
    if (tid < NWR) ll_store(x1 + (size_t)rank * NWR + tid, miscu[kTmMyRec + tid], tag);
    for (int idx = tid; idx < C * NWR; idx += kTT) sm.rec[idx] = ll_wait(x1 + idx, tag);
    cta_sync();
    if (tid == 0) tstamp(P, oi, 3);
    // winner + global max per query head (every warp redundantly)
    uint64_t wkey = kNoKey64;
    for (int r = lane; r < C; r += 32) {
        const uint64_t k = (uint64_t)sm.rec[r * NWR] | ((uint64_t)sm.rec[r * NWR + 1] << 32);
        wkey = k < wkey ? k : wkey;
    }
    wkey = warp_min_u64(wkey);
    const bool has_slot = wkey != kNoKey64;   // a slot is overwritten this step, and the new token attends to itself there
    const int widx = has_slot ? (int)(uint32_t)(wkey & 0xffffffffu) : -1;
    const int owner = has_slot ? (widx >> chs) & (C - 1) : -1;
    const bool i_own = has_slot && owner == rank;
    const int widx_local = has_slot ? (((widx >> chs) >> csh) << chs) + (widx & (CH - 1)) : -1;
    float gmax[G], newlg[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
        float m = -CUDART_INF_F;
        for (int r = lane; r < C; r += 32) {
            m = fmaxf(m, __uint_as_float(sm.rec[r * NWR + 2 + g]));
            const bool r_has_cand = (sm.rec[r * NWR] & sm.rec[r * NWR + 1]) != 0xffffffffu;
            // a candidate that is not overwritten keeps its logit
            if (r_has_cand && r != owner) m = fmaxf(m, __uint_as_float(sm.rec[r * NWR + 2 + G + g]));
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
        newlg[g] = has_slot ? misc[kTmNewLg + g] : -CUDART_INF_F;
        gmax[g] = fmaxf(m, newlg[g]);
    }
    const int patch_row = i_own ? widx_local : -1;  // this row now holds the new token

    // ---- exp and the exact fixed-point softmax denominator = X2
    {
        const int rows_pad = (local_rows + 31) & ~31;
        for (int it = tid; it < G * rows_pad; it += kTT) {
            const int g = it / rows_pad, lr = it - g * rows_pad;
            long long part = 0;
            if (lr < local_rows) {
                float gm = gmax[0], nl = newlg[0];
#pragma unroll
                for (int gg = 1; gg < G; ++gg) if (g == gg) { gm = gmax[gg]; nl = newlg[gg]; }
                const float l = lr == patch_row ? nl : lg[g * stride + lr];
                const float e = det_expf(__fsub_rn(l, gm));
                lg[g * stride + lr] = e;
                part = __float2ll_rz(__fmul_rn(e, 1099511627776.0f));
            }
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
            if (lane == 0 && part) atomicAdd(&sm.ssum[g], (unsigned long long)part);  // integers: order independent
        }
    }
    cta_sync();
    if (tid < 2 * G) ll_store(x2 + (size_t)rank * 2 * G + tid, (uint32_t)(sm.ssum[tid >> 1] >> (32 * (tid & 1))), tag);
    if (tid == 0) tstamp(P, oi, 4);
    for (int idx = tid; idx < C * 2 * G; idx += kTT) sm.sums[idx] = ll_wait(x2 + idx, tag);
    cta_sync();
    if (tid == 0) tstamp(P, oi, 5);
    float sumf[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
        long long t = 0;
        for (int r = lane; r < C; r += 32)
            t += (long long)((unsigned long long)sm.sums[(r * G + g) * 2] | ((unsigned long long)sm.sums[(r * G + g) * 2 + 1] << 32));
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
        sumf[g] = __fmul_rn(__ll2float_rn(t), 9.094947017729282e-13f);  // 2^-40
    }
    for (int it = tid; it < G * local_rows; it += kTT) {
        const int g = it / local_rows, lr = it - g * local_rows;
        float sf = sumf[0];
#pragma unroll
        for (int gg = 1; gg < G; ++gg) if (g == gg) sf = sumf[gg];
        float pr = __fdiv_rn(lg[g * stride + lr], sf);
        if (naive) pr = bf16_round(pr);
        lg[g * stride + lr] = pr;
    }
    cta_sync();
    if (tid == 0) tstamp(P, oi, 6);

    // ---- y = P.V over this warp's V tiles, warp -> CTA -> head reduction = X3
    {
        float acc[G][8];
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[g][i] = 0.f;
        for (int jj = 0; jj < nlc; ++jj) {
            int first, nv;
            attn_warp_rows(A, rank, jj, warp, &first, &nv);
            if (nv <= 0) continue;
            const uint8_t* vt = ring_wait(ring);
            const int lr0 = (jj << chs) + warp * RW;
            for (int r = sub; r < nv; r += RPI) {
                const int lr = lr0 + r;
                uint4 vv = *reinterpret_cast<const uint4*>(vt + r * kRowBytes + li * 16);
                if (lr == patch_row) vv = vnew8;
                float vf[8];
                unpack8(vv, vf);
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const float pr = lg[g * stride + lr];
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[g][i] = fmaf(pr, vf[i], acc[g][i]);
                }
            }
            ring_release(ring, P, warp, lane);
        }
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float v = acc[g][i];
#pragma unroll
                for (int d = 16; d >= LPR; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
                acc[g][i] = v;
            }
        if (kTW > kRedWarps) {   // upper warps first, then the lower ones add their own on top (fixed order)
            if (warp >= kRedWarps && lane < LPR) {
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    float4* dst = reinterpret_cast<float4*>(&sm.red[(warp - kRedWarps) * G * D + g * D + lane * 8]);
                    dst[0] = make_float4(acc[g][0], acc[g][1], acc[g][2], acc[g][3]);
                    dst[1] = make_float4(acc[g][4], acc[g][5], acc[g][6], acc[g][7]);
                }
            }
            cta_sync();
        }
        if (warp < kRedWarps && lane < LPR) {
#pragma unroll
            for (int g = 0; g < G; ++g) {
                float4* dst = reinterpret_cast<float4*>(&sm.red[warp * G * D + g * D + lane * 8]);
                float4 a = make_float4(acc[g][0], acc[g][1], acc[g][2], acc[g][3]);
                float4 b = make_float4(acc[g][4], acc[g][5], acc[g][6], acc[g][7]);
                if (kTW > kRedWarps) {
                    const float4 pa = dst[0], pb = dst[1];
                    a.x += pa.x; a.y += pa.y; a.z += pa.z; a.w += pa.w;
                    b.x += pb.x; b.y += pb.y; b.z += pb.z; b.w += pb.w;
                }
                dst[0] = a;
                dst[1] = b;
            }
        }
    }
    cta_sync();
    if (tid == 0) tstamp(P, oi, 7);
    const int SL = (G * D) >> csh;  // outputs reduced by each CTA of the head (a multiple of 4)
    for (int o = tid; o < G * D; o += kTT) {
        float s = sm.red[o];
#pragma unroll
        for (int w = 1; w < kRedWarps; ++w) s += sm.red[w * G * D + o];
        const int dest = o / SL, off = o - dest * SL;
        ll_store(x3 + ((size_t)dest * C + rank) * SL + off, __float_as_uint(s), tag);
    }

    // ======== while the partial sums travel: global state writes ========
    // owner CTA: in-place slot overwrite (cache.py:390-401 / 460-490) + policy state
    if (i_own && warp == kTW - 1) {
        if (lane < LPR) {
            uint16_t* kd = reinterpret_cast<uint16_t*>(st.k_cache) + ((size_t)h * L + widx) * D + lane * 8;
            uint16_t* vd = reinterpret_cast<uint16_t*>(st.v_cache) + ((size_t)h * L + widx) * D + lane * 8;
            *reinterpret_cast<uint4*>(kd) = knew8;   // lanes < LPR have sub == 0, li == lane
            *reinterpret_cast<uint4*>(vd) = vnew8;
        }
        float knorm = 0.f;
        if (policy == CC_POLICY_L2) {
            // key_norm[idx] = ||k_new||_2 (cache.py:592-593), canonical tree, correctly rounded sqrt
            float kf[8];
            unpack8(knew8, kf);
            float acc = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) acc = __fmaf_rn(kf[i], kf[i], acc);
            knorm = __fsqrt_rn(group_butterfly<LPR>(acc));
        }
        if (lane == 0) {
            const int32_t pos_old = posp[widx];
            st.mask[(size_t)h * L + widx] = 1;
            if (A.fill_idx) A.fill_idx[h] = widx;
            // cross-head state (head-constant pos, key_norm: read by the CTAs of every head above) is written once
            // the NEXT stage's input is complete, i.e. when all heads are past their reads
            pend->pos_ptr = nullptr; pend->kn_ptr = nullptr;
            if (st.pos_heads != 1) st.pos[(size_t)pos_row * L + widx] = ipos;
            else if (h == 0) { pend->pos_ptr = st.pos + widx; pend->pos_val = ipos; }
            if (policy == CC_POLICY_L2) {
                pend->kn_ptr = reinterpret_cast<uint16_t*>(st.key_norm) + (size_t)h * L + widx;
                const __nv_bfloat16 kb = __float2bfloat16_rn(knorm);
                pend->kn_val = *reinterpret_cast<const uint16_t*>(&kb);
            }
            pend->active = (pend->pos_ptr || pend->kn_ptr) ? 1 : 0;
            if (h == 0) st.cache_cts[0] += (pos_old == -1) ? 1 : 0;  // cache.py:330, 356-360
        }
    }
    // GQA pooling (model.py:413-418) and the score update (cache.py:716-723)
    if (naive) {
        const bool hh = policy == CC_POLICY_HEAVY_HITTER;
        for (int lr = tid; lr < local_rows; lr += kTT) {
            const int s = slot_of(lr);
            if (s >= L) continue;
            float acc = lg[lr];
#pragma unroll
            for (int g = 1; g < G; ++g) acc = __fadd_rn(acc, lg[g * stride + lr]);
            const float pooled = bf16_round(__fdiv_rn(acc, (float)G));  // torch mean: fp32 sum / G, then bf16
            if (A.probs) A.probs[(size_t)h * L + s] = __float2bfloat16_rn(pooled);
            if (hh) {
                double* num = reinterpret_cast<double*>(st.attn_history_num) + (size_t)h * L;
                int32_t* den = st.attn_history_denom + (size_t)h * L;
                const bool fresh = has_slot && s == widx;  // evicted slot's history restarts (cache.py:753-763)
                num[s] = (fresh ? 0.0 : num[s]) + (double)pooled;
                den[s] = (fresh ? 0 : den[s]) + 1;
            }
        }
        if (hh && h == 0 && rank == 0 && tid == 0) st.attn_counter[0] += 1;
    }
    // ---- X3 arrives: this CTA's slice of y (all words requested at once, summed in rank order)
    {
        const uint64_t* mine = x3 + (size_t)rank * C * SL;
        if (tid == 0) tstamp(P, oi, 8);
        for (int idx = tid; idx < C * SL; idx += kTT) sm.yx[idx] = __uint_as_float(ll_wait(mine + idx, tag));
        cta_sync();
        if (tid == 0) tstamp(P, oi, 9);
        uint64_t* yout = reinterpret_cast<uint64_t*>(P.scratch + A.y_off) + (((size_t)h * G * D + (size_t)rank * SL) >> 1);
        for (int t = tid; t < (SL >> 1); t += kTT) {
            float s0 = 0.f, s1 = 0.f;
            for (int r = 0; r < C; ++r) {
                s0 += sm.yx[r * SL + 2 * t];
                s1 += sm.yx[r * SL + 2 * t + 1];
            }
            ll_store(yout + t, pack_bf16x2(s0, s1), tag);
        }
    }
}

// ------------------------------------------------------------------ the kernel
template <int D, int G>
__global__ void __launch_bounds__(kLaunchThreads, 1) cc_token_kernel(const __grid_constant__ TokParams P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ float scratch[kTW];
    __shared__ Pending pend;
    __shared__ HeadBest wbest[kTW];
    __shared__ __align__(16) DescWindow win;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t ring_off = (P.xs_bytes + 127u) & ~127u;
    uint8_t* ring_base = smem + ring_off;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring_base + (size_t)kTW * P.n_stages * kStage);
    uint64_t* ebars = bars + kTW * P.n_stages;

    Ring ring;
    ring.ns = P.n_stages;
    ring.cs = 0;
    ring.phases = 0;
    ring.stages = ring_base + (size_t)(warp < kTW ? warp : 0) * P.n_stages * kStage;
    ring.bars = bars + (warp < kTW ? warp : 0) * P.n_stages;
    ring.empty = ebars + (warp < kTW ? warp : 0) * P.n_stages;
    if (lane == 0 && warp < kTW) {
        for (int s = 0; s < P.n_stages; ++s) {
            mbar_init(&ring.bars[s], 1);
            if (kProducer) mbar_init(&ring.empty[s], 1);
        }
        mbar_fence_init();
    }
    if (tid == 0) { pend.active = 0; win.op_hi = 0; win.att_hi = 0; }
    __syncthreads();
    // descriptor window: the first kOpWin - 1 stages, one per warp round
    if (warp < kTW) {
        for (int oi = warp; oi < kOpWin - 1; oi += kTW) {
            uint4 keep[2];
            int ka;
            window_load(P, &win, oi, lane, keep, &ka);
            if (oi < P.n_ops) {
                constexpr int nop = (int)(sizeof(TokOp) / 16), natt = (int)(sizeof(TokAttn) / 16);
                if (lane < nop) reinterpret_cast<uint4*>(&win.op[oi % kOpWin])[lane] = keep[0];
                if (ka >= 0 && lane < natt) reinterpret_cast<uint4*>(&win.att[ka % kAttWin])[lane] = keep[1];
            }
        }
    }
    __syncthreads();
    if (tid == 0) {
        const int n = min(kOpWin - 1, P.n_ops);
        int na = 0;
        for (int oi = 0; oi < n; ++oi) na += win.op[oi].type == TOK_ATTN ? 1 : 0;
        win.att_hi = na;
        win.op_hi = n;
    }
    __syncthreads();
    ring.win = &win;
    // the weight / cache stream starts before anything else: fill the whole ring
    // (an L2 prefetch cursor running further ahead -- cp.async.bulk.prefetch.L2 of the tiles after the ring -- was
    // measured SLOWER: 301 -> 273 tokens/s at 2 tiles per warp, 240 at 6; profiles/r2/notes.md)
    if (kProducer) {
        if (warp >= kTW) {
            asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(kProducerRegs));
            if (warp == kTW) producer_loop(P, &win, ring_base, bars, ebars, lane);
            return;   // every later CTA-wide barrier is the consumers' named barrier
        }
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(kConsumerRegs));
    } else {
        ring.policy = l2_policy_evict_first();
        ring.it.op = 0;
        if (lane == 0) {
            iter_enter(ring.it, P, &win, warp);
            for (int s = 0; s < P.n_stages && ring.it.op < P.n_ops; ++s) {
                iter_issue(ring.it, warp, ring.stages + (size_t)s * kStage, &ring.bars[s], ring.policy);
                iter_advance(ring.it, P, &win, warp);
            }
        }
        __syncwarp();
    }
    TokHeader* hdr = reinterpret_cast<TokHeader*>(P.scratch);
    const uint32_t tag = *reinterpret_cast<volatile uint32_t*>(&hdr->gen) + 1u;
    HeadBest best = {-CUDART_INF_F, 0x7fffffff};
    bool has_head = false;

    for (int oi = 0; oi < P.n_ops; ++oi) {
        const TokOp& op = *op_at(P, &win, oi);
        if (op.type == TOK_ATTN) {
            attn_stage<D, G>(P, &win, *att_at(P, &win, op.attn), oi, smem, ring, tag, &pend);
        } else {
            gemv_stage(P, &win, op, oi, smem, ring, tag, scratch, &pend, &best);
            has_head |= op.type == TOK_HEAD;
        }
    }
    cta_sync();
    if (tid == 0 && pend.active) {   // a program that ends with an attention stage (tests)
        if (pend.pos_ptr) *pend.pos_ptr = pend.pos_val;
        if (pend.kn_ptr) *pend.kn_ptr = (uint16_t)pend.kn_val;
        pend.active = 0;
    }

    // ---- greedy sampling (generation_utils.py:136-142): arg-max over all CTAs' rows, optional softmax
    if (has_head) {
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {   // every thread tracked the rows it finished
            const float ov = __shfl_xor_sync(0xffffffffu, best.v, d);
            const int oidx = __shfl_xor_sync(0xffffffffu, best.idx, d);
            if (ov > best.v || (ov == best.v && oidx < best.idx)) { best.v = ov; best.idx = oidx; }
        }
        if (lane == 0) wbest[warp] = best;
        cta_sync();
        HeadBest cb = wbest[0];
        for (int w = 1; w < kTW; ++w)
            if (wbest[w].v > cb.v || (wbest[w].v == cb.v && wbest[w].idx < cb.idx)) cb = wbest[w];
        uint64_t* rec1 = reinterpret_cast<uint64_t*>(P.scratch + P.rec1_off);
        const int nc = (int)gridDim.x;
        if (tid == 0) {
            ll_store(rec1 + 2 * blockIdx.x, __float_as_uint(cb.v), tag);
            ll_store(rec1 + 2 * blockIdx.x + 1, (uint32_t)cb.idx, tag);
        }
        // every CTA reduces all records (it needs the global max for the softmax anyway); all words requested at once
        float* recs = reinterpret_cast<float*>(smem);   // the staging area is free: 2 * n_ctas words
        for (int i = tid; i < 2 * nc; i += kTT) recs[i] = __uint_as_float(ll_wait(rec1 + i, tag));
        cta_sync();
        HeadBest gb = {-CUDART_INF_F, 0x7fffffff};
        for (int c = lane; c < nc; c += 32) {
            const float v = recs[2 * c];
            const int idx = (int)__float_as_uint(recs[2 * c + 1]);
            if (v > gb.v || (v == gb.v && idx < gb.idx)) { gb.v = v; gb.idx = idx; }
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, gb.v, d);
            const int oidx = __shfl_xor_sync(0xffffffffu, gb.idx, d);
            if (ov > gb.v || (ov == gb.v && oidx < gb.idx)) { gb.v = ov; gb.idx = oidx; }
        }
        if (!(P.probs_out && P.logits_out)) {
            if (blockIdx.x == 0 && tid == 0 && P.token_out) P.token_out[0] = gb.idx;   // arg-max of the logits
        } else {
            // softmax over the bf16 logits in fp32 (torch.softmax on a bf16 tensor), one bf16 rounding
            const TokOp& hop = *op_at(P, &win, P.n_ops - 1);
            int g0, nloc;
            deal_groups(hop, &g0, &nloc);
            float part = 0.f;
            for (int i = tid; i < 2 * nloc; i += kTT)
                part += expf(__bfloat162float(P.logits_out[2 * g0 + i]) - gb.v);
            part = warp_sum_f(part);
            if (lane == 0) scratch[warp] = part;
            cta_sync();
            uint64_t* rec2 = reinterpret_cast<uint64_t*>(P.scratch + P.rec2_off);
            if (tid == 0) {
                float tot = 0.f;
                for (int w = 0; w < kTW; ++w) tot += scratch[w];
                ll_store(rec2 + blockIdx.x, __float_as_uint(tot), tag);
            }
            for (int i = tid; i < nc; i += kTT) recs[i] = __uint_as_float(ll_wait(rec2 + i, tag));
            cta_sync();
            float tot = 0.f;
            for (int c = 0; c < nc; ++c) tot += recs[c];  // CTA order: deterministic
            // The reference samples arg-max of the bf16 PROBABILITIES (generation_utils.py:136-142): logits that
            // differ by less than the bf16 resolution of their probabilities tie, and the first index wins.
            HeadBest pb = {-1.f, 0x7fffffff};
            for (int i = tid; i < 2 * nloc; i += kTT) {
                const float l = __bfloat162float(P.logits_out[2 * g0 + i]);
                const __nv_bfloat16 pr = __float2bfloat16_rn(__fdiv_rn(expf(l - gb.v), tot));
                P.probs_out[2 * g0 + i] = pr;
                const float pv = __bfloat162float(pr);
                if (pv > pb.v) { pb.v = pv; pb.idx = 2 * g0 + i; }   // i ascends: the first of equal values stays
            }
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, pb.v, d);
                const int oidx = __shfl_xor_sync(0xffffffffu, pb.idx, d);
                if (ov > pb.v || (ov == pb.v && oidx < pb.idx)) { pb.v = ov; pb.idx = oidx; }
            }
            cta_sync();   // wbest / recs of the logit arg-max have been read by every warp
            if (lane == 0) wbest[warp] = pb;
            cta_sync();
            uint64_t* rec3 = reinterpret_cast<uint64_t*>(P.scratch + P.rec3_off);
            if (tid == 0) {
                HeadBest cb3 = wbest[0];
                for (int w = 1; w < kTW; ++w)
                    if (wbest[w].v > cb3.v || (wbest[w].v == cb3.v && wbest[w].idx < cb3.idx)) cb3 = wbest[w];
                ll_store(rec3 + 2 * blockIdx.x, __float_as_uint(cb3.v), tag);
                ll_store(rec3 + 2 * blockIdx.x + 1, (uint32_t)cb3.idx, tag);
            }
            if (blockIdx.x == 0 && P.token_out) {
                for (int i = tid; i < 2 * nc; i += kTT) recs[i] = __uint_as_float(ll_wait(rec3 + i, tag));
                cta_sync();
                if (warp == 0) {
                    HeadBest g3 = {-1.f, 0x7fffffff};
                    for (int c = lane; c < nc; c += 32) {
                        const float v = recs[2 * c];
                        const int idx = (int)__float_as_uint(recs[2 * c + 1]);
                        if (v > g3.v || (v == g3.v && idx < g3.idx)) { g3.v = v; g3.idx = idx; }
                    }
#pragma unroll
                    for (int d = 16; d >= 1; d >>= 1) {
                        const float ov = __shfl_xor_sync(0xffffffffu, g3.v, d);
                        const int oidx = __shfl_xor_sync(0xffffffffu, g3.idx, d);
                        if (ov > g3.v || (ov == g3.v && oidx < g3.idx)) { g3.v = ov; g3.idx = oidx; }
                    }
                    if (lane == 0) P.token_out[0] = g3.idx;
                }
            }
        }
    }
    // ---- the last CTA to finish advances the token tag
    cta_sync();
    if (tid == 0) {
        __threadfence();
        const uint32_t prev = atomicAdd(&hdr->done, 1u);
        if (prev == gridDim.x - 1) {
            hdr->done = 0;
            __threadfence();
            *reinterpret_cast<volatile uint32_t*>(&hdr->gen) = tag;
        }
    }
}

// pack / unpack between plain bf16 vectors and flag-in-data words (tests: CC_FLAG_VIA_TOKEN_KERNEL)
__global__ void tok_pack_kernel(const uint32_t* __restrict__ src, uint64_t* dst, int n_words, const TokHeader* hdr) {
    const uint32_t tag = hdr->gen + 1u;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += gridDim.x * blockDim.x)
        dst[i] = ((uint64_t)tag << 32) | src[i];
}
__global__ void tok_unpack_kernel(const uint64_t* __restrict__ src, uint32_t* dst, int n_words) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += gridDim.x * blockDim.x)
        dst[i] = (uint32_t)src[i];
}

// --------------------------------------------------------------------------- host side
using TokKernelFn = void (*)(const TokParams);

template <int D>
static TokKernelFn pick_tok_group(int G) {
    switch (G) {
    case 1: return cc_token_kernel<D, 1>;
    case 2: return cc_token_kernel<D, 2>;
    case 4: return cc_token_kernel<D, 4>;
    case 7: return cc_token_kernel<D, 7>;
    case 8: return cc_token_kernel<D, 8>;
    default: return nullptr;
    }
}
static TokKernelFn pick_tok_kernel(int D, int G) {
    if (D == 128) return pick_tok_group<128>(G);
    if (D == 64) return pick_tok_group<64>(G);
    return nullptr;
}

struct BufMap { uint64_t off; int32_t words; };

struct TokenPlan {
    TokParams p;
    TokKernelFn fn;
    TokOp* d_ops;
    TokAttn* d_attns;
    uint64_t* d_peers;
    int n_ctas;
    uint32_t smem;
    int n_layers;
    std::vector<BufMap> bufs;   // [layer][7]
    uint64_t scratch_bytes;
};

static uint64_t align_up64(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

struct AttnShape { int c_shift, ch_shift, stride, local_rows; uint32_t smem; };

static int attn_shape(const CcCacheState& st, int n_ctas, AttnShape* out) {
    const int H = st.n_kv_heads, L = st.max_cache_length, G = st.q_per_kv, D = st.head_dim;
    if (H > n_ctas) return set_error(CC_ERR_UNSUPPORTED_SHAPE, "more KV heads than SMs");
    int csh = 0;
    while ((2 << csh) * H <= n_ctas && (2 << csh) <= 128) ++csh;
    while (csh > 0 && (1 << csh) > (G * D) / 4) --csh;                 // every CTA reduces >= 4 outputs of y
    constexpr int min_chs = kTW <= 8 ? 4 : 5;                           // every warp owns >= 2 rows of a chunk
    while (csh > 0 && (1 << csh) > (L + (1 << min_chs) - 1) >> min_chs) --csh;   // every CTA gets >= 1 chunk
    int chs = 7;
    while (chs > min_chs && ((L + (1 << chs) - 1) >> chs) < (1 << csh)) --chs;
    const int C = 1 << csh, n_chunks = (L + (1 << chs) - 1) >> chs;
    const int mlc = (n_chunks + C - 1) / C;
    out->c_shift = csh; out->ch_shift = chs;
    out->local_rows = mlc << chs;
    out->stride = (out->local_rows + 31) / 32 * 32;
    out->smem = attn_smem_bytes(D, G, C, out->stride, out->local_rows);
    return CC_OK;
}

struct Layout {
    uint64_t total;
    std::vector<BufMap> bufs;   // per layer 7 entries, then one extra layer entry for the head input
    std::vector<uint64_t> x1, x2, x3;
    uint64_t rec1, rec2, rec3;
};

static int token_policy_ok(const CcCacheState& st) {
    if (st.policy == CC_POLICY_RECENT_GLOBAL || st.policy == CC_POLICY_L2) return st.attn_mode == CC_ATTN_MATH_FP32;
    if (st.policy == CC_POLICY_HEAVY_HITTER)
        return st.attn_mode == CC_ATTN_NAIVE_BF16 && st.history_window == 1 && !st.attn_thresholding;
    return 0;
}

static int build_layout(const CcTokenModel* m, int n_ctas, Layout* lay) {
    uint64_t off = 256;   // header
    auto take = [&](uint64_t words) { const uint64_t o = off; off = align_up64(off + words * 8, 256); return o; };
    const int D = m->head_dim, world = m->world;
    const int qkv_rows = (m->n_q_heads + 2 * m->n_kv_heads) * D;
    lay->bufs.assign((size_t)(m->n_layers + 1) * 7, BufMap{0, 0});
    for (int l = 0; l < m->n_layers; ++l) {
        BufMap* b = &lay->bufs[(size_t)l * 7];
        b[0] = {take(qkv_rows / 2), qkv_rows / 2};
        b[1] = {take(m->n_q_heads * D / 2), m->n_q_heads * D / 2};
        b[2] = {take((uint64_t)world * m->dim / 2), world * m->dim / 2};
        b[3] = {take(m->dim / 2), m->dim / 2};
        b[4] = {take(m->intermediate / 2), m->intermediate / 2};
        b[5] = {take((uint64_t)world * m->dim / 2), world * m->dim / 2};
        if (l > 0) b[6] = {take(m->dim / 2), m->dim / 2};
        const CcCacheState& st = m->layers[l].cache;
        AttnShape sh;
        if (int rc = attn_shape(st, n_ctas, &sh)) return rc;
        const int C = 1 << sh.c_shift, G = st.q_per_kv;
        lay->x1.push_back(take((uint64_t)st.n_kv_heads * C * (2 + 2 * G)));
        lay->x2.push_back(take((uint64_t)st.n_kv_heads * C * 2 * G));
        lay->x3.push_back(take((uint64_t)st.n_kv_heads * C * G * D));
    }
    lay->bufs[(size_t)m->n_layers * 7 + 6] = {take(m->dim / 2), m->dim / 2};
    lay->rec1 = take(2 * (uint64_t)n_ctas);
    lay->rec2 = take((uint64_t)n_ctas);
    lay->rec3 = take(2 * (uint64_t)n_ctas);
    lay->total = off;
    return CC_OK;
}

static int validate_model(const CcTokenModel* m) {
    if (!m || !m->layers || m->n_layers < 1) return set_error(CC_ERR_INVALID_ARG, "cc_token: null model or no layers");
    if (m->world < 1 || m->world > kTokMaxWorld || m->rank < 0 || m->rank >= m->world)
        return set_error(CC_ERR_INVALID_ARG, "cc_token: bad world / rank");
    const bool attn_only_prog = m->n_layers == 1 && m->layers[0].wqkv == nullptr;   // tests: one attention stage
    if (!attn_only_prog && (m->dim % 256 || m->intermediate % 256 || (m->n_q_heads * m->head_dim) % 256))
        return set_error(CC_ERR_UNSUPPORTED_SHAPE, "cc_token: dim, intermediate size and n_q_heads * head_dim must be multiples of 256");
    if ((m->n_q_heads * m->head_dim) % 8) return set_error(CC_ERR_UNSUPPORTED_SHAPE, "cc_token: n_q_heads * head_dim must be a multiple of 8");
    if (m->head_dim != 64 && m->head_dim != 128) return set_error(CC_ERR_UNSUPPORTED_SHAPE, "cc_token: head_dim must be 64 or 128");
    if (m->n_kv_heads < 1 || m->n_q_heads % m->n_kv_heads) return set_error(CC_ERR_INVALID_ARG, "cc_token: bad head counts");
    if (m->vocab & 1) return set_error(CC_ERR_UNSUPPORTED_SHAPE, "cc_token: vocabulary rows must be even");
    const int G = m->n_q_heads / m->n_kv_heads;
    if (!pick_tok_kernel(m->head_dim, G)) return set_error(CC_ERR_UNSUPPORTED_SHAPE, "cc_token: unsupported q_per_kv (1, 2, 4, 7, 8)");
    for (int l = 0; l < m->n_layers; ++l) {
        const CcTokenLayer& y = m->layers[l];
        const CcCacheState& st = y.cache;
        if (!y.wqkv || !y.attention_norm) {   // attention-only program (tests)
            if (m->n_layers != 1) return set_error(CC_ERR_INVALID_ARG, "cc_token: null layer weights");
        } else if (!y.wo || !y.ffn_norm || !y.w1 || !y.w3 || !y.w2) {
            return set_error(CC_ERR_INVALID_ARG, "cc_token: null layer weights");
        }
        if (st.n_kv_heads != m->n_kv_heads || st.q_per_kv != G || st.head_dim != m->head_dim)
            return set_error(CC_ERR_INVALID_ARG, "cc_token: cache shape does not match the model");
        if (!token_policy_ok(st))
            return set_error(CC_ERR_UNSUPPORTED_SHAPE, "cc_token: cache policy not supported by the whole-token kernel "
                                                       "(recent_global, l2, heavy_hitter with full history)");
        if (!st.k_cache || !st.v_cache || !st.pos || !st.mask || !st.cache_cts) return set_error(CC_ERR_INVALID_ARG, "cc_token: null cache buffer");
        if (st.pos_heads != 1 && st.pos_heads != st.n_kv_heads) return set_error(CC_ERR_INVALID_ARG, "cc_token: pos_heads must be 1 or n_kv_heads");
        if (st.policy == CC_POLICY_HEAVY_HITTER && (!st.attn_history_num || !st.attn_history_denom || !st.attn_counter))
            return set_error(CC_ERR_INVALID_ARG, "cc_token: null heavy_hitter state");
        if (st.policy == CC_POLICY_L2 && !st.key_norm) return set_error(CC_ERR_INVALID_ARG, "cc_token: null key_norm");
        if (st.policy == CC_POLICY_RECENT_GLOBAL && st.global_tokens >= st.max_cache_length)
            return set_error(CC_ERR_INVALID_ARG, "cc_token: global_tokens must be < max_cache_length");
    }
    return CC_OK;
}

static int token_device(int* n_sms, int* smem_optin) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return set_error(CC_ERR_NO_DEVICE, "no CUDA device");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return set_error(CC_ERR_NO_DEVICE, "cudaGetDeviceProperties failed");
    if (prop.major != 10) return set_error(CC_ERR_NO_DEVICE, "this library is built for sm_100a (Blackwell B200) only");
    if (!prop.cooperativeLaunch) return set_error(CC_ERR_NO_DEVICE, "device does not support cooperative launches");
    *n_sms = prop.multiProcessorCount;
    *smem_optin = (int)prop.sharedMemPerBlockOptin;
    return CC_OK;
}

unsigned long long token_scratch_bytes_impl(const CcTokenModel* m) {
    if (validate_model(m)) return 0;
    int n_sms = 148, optin = 0;
    if (const char* e = getenv("CC_B200_TOKEN_CTAS")) n_sms = atoi(e);
    else if (token_device(&n_sms, &optin)) n_sms = 148;   // sizing without a device (tests of the host logic)
    Layout lay;
    if (build_layout(m, n_sms, &lay)) return 0;
    return lay.total;
}

static void seg_plan(int K, int rg, int* seg_len, int* n_seg) {
    const int cap = (int)(kStage / 2) / rg;   // elements per row per stage
    int sl = K < cap ? K : cap;
    sl = sl / 256 * 256;
    *seg_len = sl;
    *n_seg = (K + sl - 1) / sl;
}

int token_plan_create_impl(const CcTokenModel* m, void** out_plan) {
    if (!out_plan) return set_error(CC_ERR_INVALID_ARG, "cc_token_plan_create: null out pointer");
    if (int rc = validate_model(m)) return rc;
    if (!m->scratch_peers) return set_error(CC_ERR_INVALID_ARG, "cc_token_plan_create: null scratch_peers");
    int n_sms = 0, optin = 0;
    if (int rc = token_device(&n_sms, &optin)) return rc;
    int n_ctas = n_sms;
    if (const char* e = getenv("CC_B200_TOKEN_CTAS")) n_ctas = atoi(e) > 0 ? atoi(e) : n_ctas;   // experiments only
    Layout lay;
    if (int rc = build_layout(m, n_ctas, &lay)) return rc;
    if (m->scratch_bytes < lay.total) return set_error(CC_ERR_INVALID_ARG, "cc_token_plan_create: scratch smaller than cc_token_scratch_bytes()");
    const int D = m->head_dim, G = m->n_q_heads / m->n_kv_heads, world = m->world;
    const bool attn_only = m->layers[0].wqkv == nullptr;
    std::vector<TokOp> ops;
    std::vector<TokAttn> attns;
    uint32_t xs_bytes = 1024;
    int rot = 0;
    auto next_rot = [&](int n_groups) { const int r = rot; rot = (rot + n_groups % n_ctas) % n_ctas; return r; };
    auto base_op = [&](int type, int K, int N, int rg, int layer) {
        TokOp o;
        memset(&o, 0, sizeof(o));
        (void)layer;
        o.type = type; o.K = K; o.N = N; o.n_groups = N / 2;
        seg_plan(K, rg, &o.seg_len, &o.n_seg);
        o.rot = next_rot(o.n_groups);
        o.grp_base = o.n_groups / n_ctas; o.grp_rem = o.n_groups % n_ctas;
        o.publish_cta = -1; o.in_world = 1; o.out_world = 1; o.attn = -1;
        // staged input + one float4 of partial sums per tile of a CTA
        const uint32_t need = (((uint32_t)K * 2 + 15u) & ~15u) + (uint32_t)(o.grp_base + 1) * o.n_seg * 16u;
        if (need > xs_bytes) xs_bytes = need;
        return o;
    };
    const int qkv_rows = (m->n_q_heads + 2 * m->n_kv_heads) * D;
    for (int l = 0; l < m->n_layers; ++l) {
        const CcTokenLayer& y = m->layers[l];
        const BufMap* b = &lay.bufs[(size_t)l * 7];
        if (!attn_only) {
            TokOp o = base_op(TOK_QKV, m->dim, qkv_rows, 2, l);
            if (l == 0) {
                o.in_kind = IN_TOKEN;
            } else {   // x_l = h_{l-1} + all-reduce(w2 partials of layer l-1)
                const BufMap* pb = &lay.bufs[(size_t)(l - 1) * 7];
                o.in_kind = IN_LL; o.in_world = world; o.in_off = pb[5].off;
                o.res_kind = RES_LL; o.res_off = pb[3].off;
                o.publish_cta = (2 * l) % n_ctas; o.sum_off = b[6].off;
            }
            o.norm_w = (const __nv_bfloat16*)y.attention_norm; o.eps = y.attention_norm_eps;
            o.w = (const __nv_bfloat16*)y.wqkv; o.bias = (const __nv_bfloat16*)y.bqkv;
            o.out_off = b[0].off;
            ops.push_back(o);
        }
        {
            TokOp o;
            memset(&o, 0, sizeof(o));
            o.type = TOK_ATTN; o.attn = (int)attns.size(); o.publish_cta = -1;
            TokAttn a;
            memset(&a, 0, sizeof(a));
            a.st = y.cache;
            AttnShape sh;
            if (int rc = attn_shape(a.st, n_ctas, &sh)) return rc;
            a.c_shift = sh.c_shift; a.ch_shift = sh.ch_shift; a.lg_stride = sh.stride;
            a.rot = (l * 37) % n_ctas;
            a.scale = (float)(1.0 / sqrt((double)D));
            a.qkv_off = b[0].off; a.y_off = b[1].off;
            a.x1_off = lay.x1[l]; a.x2_off = lay.x2[l]; a.x3_off = lay.x3[l];
            a.fill_idx = y.fill_idx; a.probs = (__nv_bfloat16*)y.probs;
            if (sh.smem > xs_bytes) xs_bytes = sh.smem;
            attns.push_back(a);
            ops.push_back(o);
        }
        if (attn_only) continue;
        {   // wo: row-parallel under tensor parallelism, partial sums to every rank
            TokOp o = base_op(TOK_GEMV, m->n_q_heads * D, m->dim, 2, l);
            o.in_kind = IN_LL; o.in_off = b[1].off;
            o.w = (const __nv_bfloat16*)y.wo;
            o.out_off = b[2].off; o.out_world = world;
            ops.push_back(o);
        }
        {   // h = x + all-reduce(wo partials); ffn_norm; silu(w1 h) * w3 h
            TokOp o = base_op(TOK_GATE, m->dim, m->intermediate, 4, l);
            o.in_kind = IN_LL; o.in_world = world; o.in_off = b[2].off;
            if (l == 0) o.res_kind = RES_TOKEN; else { o.res_kind = RES_LL; o.res_off = b[6].off; }
            o.publish_cta = (2 * l + 1) % n_ctas; o.sum_off = b[3].off;
            o.norm_w = (const __nv_bfloat16*)y.ffn_norm; o.eps = y.ffn_norm_eps;
            o.w = (const __nv_bfloat16*)y.w1; o.w3 = (const __nv_bfloat16*)y.w3;
            o.out_off = b[4].off;
            ops.push_back(o);
        }
        {   // w2: row-parallel
            TokOp o = base_op(TOK_GEMV, m->intermediate, m->dim, 2, l);
            o.in_kind = IN_LL; o.in_off = b[4].off;
            o.w = (const __nv_bfloat16*)y.w2;
            o.out_off = b[5].off; o.out_world = world;
            ops.push_back(o);
        }
    }
    if (!attn_only) {
        if (!m->output || !m->norm || !m->tok_embeddings || !m->freqs_cis)
            return set_error(CC_ERR_INVALID_ARG, "cc_token_plan_create: null embedding / norm / output / freqs_cis");
        const int l = m->n_layers;
        const BufMap* pb = &lay.bufs[(size_t)(l - 1) * 7];
        TokOp o = base_op(TOK_HEAD, m->dim, m->vocab, 2, l);
        o.in_kind = IN_LL; o.in_world = world; o.in_off = pb[5].off;
        o.res_kind = RES_LL; o.res_off = pb[3].off;
        o.publish_cta = (2 * l) % n_ctas; o.sum_off = lay.bufs[(size_t)l * 7 + 6].off;
        o.norm_w = (const __nv_bfloat16*)m->norm; o.eps = m->norm_eps;
        o.w = (const __nv_bfloat16*)m->output;
        ops.push_back(o);
    }
    xs_bytes = (xs_bytes + 127u) & ~127u;
    TokKernelFn fn = pick_tok_kernel(D, G);
    cudaFuncAttributes fattr;
    if (cudaFuncGetAttributes(&fattr, fn) != cudaSuccess) {
        cudaGetLastError();
        return set_error(CC_ERR_NO_DEVICE, "cc_token: no kernel image for this device");
    }
    const int max_dyn = optin - (int)fattr.sharedSizeBytes;   // static shared memory counts against the same limit
    // shared memory: staging area | ring (8 warps x n_stages x 8 KiB) | mbarriers
    const int avail = max_dyn - (int)xs_bytes - (8 * kTW * 16 + 64) /* full + empty barriers of up to 8 stages */;
    int n_stages = avail / (int)(kTW * kStage);
    if (n_stages > 8) n_stages = 8;
    if (const char* e = getenv("CC_B200_TOKEN_STAGES")) n_stages = atoi(e) < n_stages && atoi(e) > 0 ? atoi(e) : n_stages;
    if (n_stages < 2) return set_error(CC_ERR_UNSUPPORTED_SHAPE, "cc_token: staging area leaves fewer than 2 ring stages per warp");
    const uint32_t smem = ((xs_bytes + 127u) & ~127u) + (uint32_t)n_stages * kTW * kStage + (uint32_t)n_stages * kTW * 16 + 64;

    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn) != cudaSuccess) {
        cudaGetLastError();
        return set_error(CC_ERR_CUDA, "cc_token: cannot raise the dynamic shared memory limit");
    }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kLaunchThreads, smem) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        return set_error(CC_ERR_UNSUPPORTED_SHAPE, "cc_token: the kernel does not fit an SM with this shared-memory plan");
    }
    if (n_ctas > n_sms * per_sm) return set_error(CC_ERR_UNSUPPORTED_SHAPE, "cc_token: more CTAs than can be co-resident");

    TokenPlan* plan = new TokenPlan();
    plan->fn = fn; plan->n_ctas = n_ctas; plan->smem = smem; plan->n_layers = m->n_layers;
    plan->bufs = lay.bufs; plan->scratch_bytes = lay.total;
    cudaError_t e = cudaMalloc(&plan->d_ops, ops.size() * sizeof(TokOp));
    if (e == cudaSuccess) e = cudaMalloc(&plan->d_attns, attns.size() * sizeof(TokAttn));
    if (e == cudaSuccess) e = cudaMalloc(&plan->d_peers, kTokMaxWorld * sizeof(uint64_t));
    if (e == cudaSuccess) e = cudaMemcpy(plan->d_ops, ops.data(), ops.size() * sizeof(TokOp), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(plan->d_attns, attns.data(), attns.size() * sizeof(TokAttn), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(plan->d_peers, m->scratch_peers, world * sizeof(uint64_t), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { delete plan; return set_cuda_error(e, "cc_token_plan_create"); }
    TokParams& p = plan->p;
    memset(&p, 0, sizeof(p));
    p.ops = plan->d_ops; p.attns = plan->d_attns; p.n_ops = (int)ops.size();
    p.world = world; p.rank = m->rank; p.vocab = m->vocab;
    p.scratch = reinterpret_cast<uint8_t*>(m->scratch_peers[m->rank]);
    p.peers = plan->d_peers;
    p.emb = (const __nv_bfloat16*)m->tok_embeddings; p.freqs = (const __nv_bfloat16*)m->freqs_cis;
    p.dim = m->dim; p.n_q_rows = m->n_q_heads * D; p.n_kv_rows = m->n_kv_heads * D; p.head_dim = D;
    p.xs_bytes = xs_bytes; p.n_stages = n_stages;


    p.rec1_off = lay.rec1; p.rec2_off = lay.rec2; p.rec3_off = lay.rec3;
    *out_plan = plan;
    return CC_OK;
}

int token_plan_destroy_impl(void* vp) {
    TokenPlan* plan = static_cast<TokenPlan*>(vp);
    if (!plan) return CC_OK;
    cudaFree(plan->d_ops); cudaFree(plan->d_attns); cudaFree(plan->d_peers);
    delete plan;
    return CC_OK;
}

int decode_token_impl(void* vp, const int32_t* token_in, const int32_t* input_pos, int32_t* token_out, void* logits_out,
                      void* probs_out, cudaStream_t stream) {
    TokenPlan* plan = static_cast<TokenPlan*>(vp);
    if (!plan || !input_pos) return set_error(CC_ERR_INVALID_ARG, "cc_decode_token: null plan or input_pos");
    if (probs_out && !logits_out) return set_error(CC_ERR_INVALID_ARG, "cc_decode_token: probs_out needs logits_out");
    TokParams p = plan->p;
    p.token_in = token_in; p.input_pos = input_pos; p.token_out = token_out;
    p.logits_out = (__nv_bfloat16*)logits_out; p.probs_out = (__nv_bfloat16*)probs_out;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(plan->n_ctas);
    cfg.blockDim = dim3(kLaunchThreads);
    cfg.dynamicSmemBytes = plan->smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeCooperative;   // the stages spin on each other's output: every CTA must be resident
    attr.val.cooperative = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, plan->fn, p);
    if (e != cudaSuccess) return set_cuda_error(e, "cc_decode_token launch");
    return CC_OK;
}

int token_buffer_impl(void* vp, int layer, int which, uint64_t* off, int32_t* words) {
    TokenPlan* plan = static_cast<TokenPlan*>(vp);
    if (!plan || layer < 0 || layer > plan->n_layers || which < 0 || which > 6) return set_error(CC_ERR_INVALID_ARG, "cc_token_buffer: bad argument");
    const BufMap& b = plan->bufs[(size_t)layer * 7 + which];
    if (!b.words) return set_error(CC_ERR_INVALID_ARG, "cc_token_buffer: no such buffer");
    if (off) *off = b.off;
    if (words) *words = b.words;
    return CC_OK;
}

int token_set_trace_impl(void* vp, void* buf) {
    TokenPlan* plan = static_cast<TokenPlan*>(vp);
    if (!plan) return set_error(CC_ERR_INVALID_ARG, "cc_token_set_trace: null plan");
    plan->p.trace = static_cast<unsigned long long*>(buf);
    return CC_OK;
}

int token_info_impl(void* vp, int32_t* n_ctas, int32_t* n_ops, int32_t* ring_stages, int32_t* smem) {
    TokenPlan* plan = static_cast<TokenPlan*>(vp);
    if (!plan) return set_error(CC_ERR_INVALID_ARG, "cc_token_info: null plan");
    if (n_ctas) *n_ctas = plan->n_ctas;
    if (n_ops) *n_ops = plan->p.n_ops;
    if (ring_stages) *ring_stages = plan->p.n_stages;
    if (smem) *smem = (int32_t)plan->smem;
    return CC_OK;
}

// cc_decode_step with CC_FLAG_VIA_TOKEN_KERNEL: one attention stage of the whole-token kernel on caller-supplied
// q / k / v (tests: the stage is held to the same bit-exact bar as the per-layer kernel).  Synchronous.
int decode_step_via_token_impl(const CcCacheState* st, const CcDecodeArgs* a, cudaStream_t stream) {
    if (a->phases != CC_PHASE_ALL) return set_error(CC_ERR_INVALID_ARG, "the whole-token kernel runs the fused step only (phases = ALL)");
    CcTokenLayer layer;
    memset(&layer, 0, sizeof(layer));
    layer.cache = *st;
    layer.fill_idx = a->fill_idx;
    layer.probs = a->probs;
    CcTokenModel m;
    memset(&m, 0, sizeof(m));
    m.n_layers = 1; m.dim = 256; m.intermediate = 256; m.vocab = 2;
    m.n_kv_heads = st->n_kv_heads; m.n_q_heads = st->n_kv_heads * st->q_per_kv; m.head_dim = st->head_dim;
    m.layers = &layer; m.world = 1; m.rank = 0;
    if (int rc = validate_model(&m)) return rc;
    const uint64_t bytes = token_scratch_bytes_impl(&m);
    if (!bytes) return CC_ERR_INVALID_ARG;
    uint8_t* scratch = nullptr;
    cudaError_t e = cudaMalloc(&scratch, bytes);
    if (e != cudaSuccess) return set_cuda_error(e, "cc_decode_step (token kernel) scratch");
    cudaMemsetAsync(scratch, 0, bytes, stream);
    uint64_t peer = reinterpret_cast<uint64_t>(scratch);
    m.scratch_peers = &peer; m.scratch_bytes = bytes;
    void* plan = nullptr;
    int rc = token_plan_create_impl(&m, &plan);
    if (rc == CC_OK) {
        TokenPlan* tp = static_cast<TokenPlan*>(plan);
        const int D = st->head_dim, nq = m.n_q_heads * D, nkv = m.n_kv_heads * D;
        uint64_t* qkv = reinterpret_cast<uint64_t*>(scratch + tp->bufs[0].off);
        const TokHeader* hdr = reinterpret_cast<const TokHeader*>(scratch);
        tok_pack_kernel<<<8, 256, 0, stream>>>((const uint32_t*)a->q, qkv, nq / 2, hdr);
        tok_pack_kernel<<<4, 256, 0, stream>>>((const uint32_t*)a->k_new, qkv + nq / 2, nkv / 2, hdr);
        tok_pack_kernel<<<4, 256, 0, stream>>>((const uint32_t*)a->v_new, qkv + (nq + nkv) / 2, nkv / 2, hdr);
        rc = decode_token_impl(plan, nullptr, a->input_pos, nullptr, nullptr, nullptr, stream);
        if (rc == CC_OK) {
            tok_unpack_kernel<<<8, 256, 0, stream>>>(reinterpret_cast<const uint64_t*>(scratch + tp->bufs[1].off), (uint32_t*)a->y, nq / 2);
            e = cudaStreamSynchronize(stream);
            if (e != cudaSuccess) rc = set_cuda_error(e, "cc_decode_step (token kernel)");
        }
        token_plan_destroy_impl(plan);
    }
    cudaFree(scratch);
    return rc;
}

}  // namespace cc
