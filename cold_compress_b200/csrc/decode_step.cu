// decode_step.cu -- the fused per-layer decode step for sm_100a.
//
// One launch = eviction decision -> in-place slot overwrite -> masked attention over the live
// cache slots -> GQA pooling -> eviction-score update, for every KV head of one layer.
// Replaces the reference's decode branch of Attention.forward (model.py:389-427):
// KVCache.update_kv (cache.py:314-364), scaled_dot_product_attention
// (attention_utils.py:8-54), the group mean (model.py:413-418) and KVCache.update_state
// (cache.py:690-723), which run there as ~25 ATen ops and materialise 4x K/V copies.
//
// Shape of the launch (HBM-latency bound: 4 MiB of K/V per layer at L=1024):
//   * one thread-block CLUSTER per KV head, C CTAs (<=16) per cluster; the cache rows of the head
//     are cut into chunks of CH rows (16/32/64) dealt round-robin to the CTAs of the cluster, so
//     a head whose live slots are a prefix (hybrid, caches still filling) stays balanced
//     -> Hkv*C CTAs (128 of 148 SMs for Llama-3-8B);
//   * the K and V rows of each chunk are streamed by the TMA engine (cp.async.bulk, UBLKCP) into
//     a shared-memory ring, as many chunks in flight as the ring holds from the first instruction;
//   * three DSMEM all-to-all exchanges per step: {arg-min candidate, row max}, {softmax sum},
//     {P.V partials}; the first one carries both the eviction decision and the softmax max, so
//     the evicted slot's old logit is excluded / the new token's logit included without a
//     second pass over K;
//   * arithmetic order is the bit-reproducible definition in common.cuh (canonical dot tree,
//     det_expf, 2^-40 fixed-point sum), so eviction indices do not depend on C, CH or the
//     thread count.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/cc_b200.h"
#include "common.cuh"
#include "internal.h"

namespace cc {

#ifndef CC_DECODE_THREADS
#define CC_DECODE_THREADS 256
#endif
constexpr int kThreads = CC_DECODE_THREADS;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxCluster = 16;
constexpr int kMaxG = 8;
constexpr int kRecWords = 2 + 2 * kMaxG;  // X1 record: key64, lmax[G], candlogit[G]
constexpr uint64_t kNoKey = ~0ull;

struct SmemLayout {
    uint32_t kring, vring, lg, red, yx, rec, sums, ssum, x0, maskb, bars, misc, total;
};

struct DecodeParams {
    SmemLayout lay;     // shared-memory carve-up, computed once on the host
    CcCacheState st;
    CcDecodeArgs a;
    int ch_shift;       // log2(rows per chunk)
    int max_local_chunks;  // chunks dealt to rank 0 (the most any CTA gets)
    int n_stages;       // ring depth for K and for V
    int lg_stride;      // floats per query head in the logit/prob array
    float scale;        // (float)(1/sqrt(D))
    int c_shift;        // log2(cluster size)
    unsigned long long* trace;  // diagnostics: [gridDim.x][16] globaltimer stamps (NULL = off)
};

__host__ __device__ inline uint32_t align_up(uint32_t x, uint32_t a) { return (x + a - 1) / a * a; }

__host__ __device__ inline SmemLayout make_layout(int D, int G, int chunk_rows, int local_rows, int n_stages, int lg_stride) {
    SmemLayout s;
    uint32_t off = 0;
    const uint32_t chunk_bytes = chunk_rows * D * 2;
    s.kring = off; off += align_up(n_stages * chunk_bytes, 128);
    s.vring = off; off += align_up(n_stages * chunk_bytes, 128);
    s.lg = off;    off += align_up(G * lg_stride * 4, 128);
    s.red = off;   off += kWarps * G * D * 4;
    s.yx = off;    off += G * D * 4;
    s.rec = off;   off += kMaxCluster * kRecWords * 4;
    s.sums = off;  off += kMaxCluster * kMaxG * 8;
    s.ssum = off;  off += kMaxG * 8;
    s.x0 = off;    off += 128;
    s.maskb = off; off += align_up(local_rows, 128);
    s.bars = off;  off += align_up((2 * n_stages + 4) * 8, 128);  // K ring, V ring, 4 exchange barriers
    s.misc = off;  off += 512;
    s.total = off;
    return s;
}

// misc area, indices in 4-byte words
enum : int {
    kMiscNewLg = 0,     // [kMaxG] logits of the new token
    kMiscLmax = 8,      // [kMaxG]
    kMiscCandLg = 16,   // [kMaxG]
    kMiscGmax = 24,     // [kMaxG]
    kMiscSumF = 32,     // [kMaxG]
    kMiscMyRec = 40,    // [kRecWords]
    kMiscWarpKey = 64,  // [kWarps] u64 (32 words)
    kMiscWarpMax = 96,  // [kWarps]
    kMiscWords = 128,
};

// hybrid per-head decision (reference cache.py:896-950)
enum : int { kHybAppend = 0, kHybEvict = 1, kHybDrop = 2 };

__device__ __forceinline__ void stamp(const DecodeParams& p, int k) {
    if (p.trace && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.trace[(size_t)blockIdx.x * 16 + k] = t;
    }
}

template <int D, int G, int POLICY, bool NAIVE>
__global__ void __launch_bounds__(kThreads, 1)
cc_decode_kernel(const __grid_constant__ DecodeParams p) {
    constexpr int LPR = D / 8;           // lanes per cache row (16-byte slice each)
    constexpr int RPW = 32 / LPR;        // rows per warp iteration
    constexpr int NRG = kThreads / LPR;  // row groups in the P.V phase
    constexpr uint32_t kRowBytes = D * 2;
    constexpr bool hybrid = POLICY == CC_POLICY_HYBRID;

    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int li = lane % LPR, sub = lane / LPR;
    const int csh = p.c_shift, C = 1 << csh, rank = (int)cluster_ctarank();
    const int h = blockIdx.x >> csh;
    const CcCacheState& st = p.st;
    const int L = st.max_cache_length, H = st.n_kv_heads;
    const int chs = p.ch_shift, CH = 1 << chs;
    const uint32_t chunk_bytes = CH * kRowBytes;
    const int phases = p.a.phases;
    const bool do_evict = phases & CC_PHASE_EVICT_FILL, do_attend = phases & CC_PHASE_ATTEND,
               do_score = phases & CC_PHASE_SCORE;
    const int NS = p.n_stages, stride = p.lg_stride;
    // shared (cross-head) state is read by every cluster and written by one: order with an epoch barrier
    // heavy hitter with a history window (W > 1): the shared step counter selects the ring column every head writes
    const bool hh_ring = POLICY == CC_POLICY_HEAVY_HITTER && st.history_window > 1;
    const bool needs_epoch = (do_evict && (POLICY == CC_POLICY_RECENT_GLOBAL || POLICY == CC_POLICY_L2)) ||
                             ((hybrid || hh_ring) && (do_evict || do_score));

    // chunk-cyclic ownership: local chunk j <-> global chunk rank + j*C; local row lr = j*CH + r
    const int n_chunks_total = (L + CH - 1) >> chs;
    const int nlc = n_chunks_total > rank ? (n_chunks_total - rank + C - 1) >> csh : 0;
    const int local_rows = nlc << chs;
    auto slot_of = [&](int lr) { return ((((lr >> chs) << csh) + rank) << chs) + (lr & (CH - 1)); };

    const SmemLayout& lay = p.lay;
    uint8_t* kring = smem + lay.kring;
    uint8_t* vring = smem + lay.vring;
    float* lg = reinterpret_cast<float*>(smem + lay.lg);
    float* red = reinterpret_cast<float*>(smem + lay.red);
    float* yx = reinterpret_cast<float*>(smem + lay.yx);
    uint32_t* rec = reinterpret_cast<uint32_t*>(smem + lay.rec);
    unsigned long long* sums = reinterpret_cast<unsigned long long*>(smem + lay.sums);
    float* x0 = reinterpret_cast<float*>(smem + lay.x0);
    uint8_t* maskb = smem + lay.maskb;
    uint64_t* kbar = reinterpret_cast<uint64_t*>(smem + lay.bars);
    uint64_t* vbar = kbar + NS;
    uint64_t* xbar = vbar + NS;  // [0] X0 (L2 max), [1] X1, [2] X2, [3] X3: single-use exchange barriers
    float* misc = reinterpret_cast<float*>(smem + lay.misc);
    uint32_t* miscu = reinterpret_cast<uint32_t*>(misc);
    unsigned long long* ssum = reinterpret_cast<unsigned long long*>(smem + lay.ssum);  // [kMaxG] local fixed-point sums

    stamp(p, 0);
    constexpr int NW = 2 + 2 * G;  // words of an X1 record
    if (tid == 0) {
        for (int s = 0; s < 2 * NS + 4; ++s) mbar_init(&kbar[s], 1);
        mbar_fence_init();
        // every exchange delivers a known number of bytes into this CTA: arm the barriers right away
        // (only the exchanges this launch performs: an armed barrier nobody completes is a synccheck finding)
        if (POLICY == CC_POLICY_L2 && do_evict) mbar_expect_tx(&xbar[0], (uint32_t)C * 4);
        mbar_expect_tx(&xbar[1], (uint32_t)C * NW * 4);
        if (do_attend) {
            mbar_expect_tx(&xbar[2], (uint32_t)C * G * 8);
            mbar_expect_tx(&xbar[3], (uint32_t)G * D * 4);
        }
    }
    // B0: this CTA has started and its exchange barriers exist; peers wait for this before they send
    cluster_arrive_release();
    bool b0_waited = false;

    // ---- hybrid: per-head decision from scalars (every CTA of the cluster computes the same)
    int hyb_mode = kHybAppend, hyb_bits = 0, hyb_cts = 0, hyb_rw = 0;
    int64_t hyb_counter = 0;
    int live_limit = L;  // slots >= live_limit cannot be live: their chunks are never loaded
    const bool is_punc = hybrid && (p.a.is_punc_dev ? p.a.is_punc_dev[0] != 0 : p.a.is_punc != 0);
    if constexpr (hybrid) {
        hyb_bits = st.hyb_strategy_bits[h];
        hyb_cts = st.cache_cts[h];
        hyb_rw = (hyb_bits & CC_HYB_WINDOW) ? st.hyb_recent_window[h] : 0;
        if (st.attn_counter) hyb_counter = st.attn_counter[0];
        if (do_evict) {
            if (((hyb_bits & CC_HYB_PUNC) && is_punc) || (hyb_bits & CC_HYB_FULL)) {
                hyb_mode = kHybAppend;
            } else {
                int budget = st.global_tokens;
                if ((hyb_bits & CC_HYB_SPECIAL) && st.num_special) budget += st.num_special[0];
                if ((hyb_bits & CC_HYB_PUNC) && st.num_punc) budget += st.num_punc[0];
                if (hyb_bits & CC_HYB_WINDOW) budget += st.hyb_recent_window[h];
                if (hyb_bits & CC_HYB_HEAVY_HITTER) budget += st.hyb_heavy_budget[h];
                hyb_mode = hyb_cts < budget ? kHybAppend
                         : (hyb_bits & (CC_HYB_HEAVY_HITTER | CC_HYB_WINDOW)) ? kHybEvict : kHybDrop;
            }
        }
        live_limit = min(L, hyb_cts + ((do_evict && hyb_mode == kHybAppend) ? 1 : 0));
    }
    int32_t hh_cts = 0;   // cache_cts before this step (thresholded history compares against 1 / cache_cts)
    if (hh_ring) {
        hyb_counter = st.attn_counter[0];
        hh_cts = st.cache_cts[0];
    }
    const int live_chunks_total = (live_limit + CH - 1) >> chs;
    const int nlc_live = do_attend ? (live_chunks_total > rank ? (live_chunks_total - rank + C - 1) >> csh : 0) : 0;
    const int live_rows = nlc_live << chs;

    const uint16_t* kc = reinterpret_cast<const uint16_t*>(st.k_cache) + (size_t)h * L * D;
    const uint16_t* vc = reinterpret_cast<const uint16_t*>(st.v_cache) + (size_t)h * L * D;
    auto chunk_rows_valid = [&](int j) { return min(CH, L - ((((j << csh) + rank)) << chs)); };

    // ---- 0. barriers + as many K/V chunks in flight as the ring holds
    if (tid == 0) {
        fence_proxy_async();
        const int pre = min(NS, nlc_live);
        for (int j = 0; j < pre; ++j) {
            const uint32_t bytes = chunk_rows_valid(j) * kRowBytes;
            mbar_expect_tx(&kbar[j], bytes);
            tma_load_1d(kring + j * chunk_bytes, kc + ((size_t)((j << csh) + rank) << chs) * D, bytes, &kbar[j]);
        }
        for (int j = 0; j < pre; ++j) {
            const uint32_t bytes = chunk_rows_valid(j) * kRowBytes;
            mbar_expect_tx(&vbar[j], bytes);
            tma_load_1d(vring + j * chunk_bytes, vc + ((size_t)((j << csh) + rank) << chs) * D, bytes, &vbar[j]);
        }
    }
    if (tid < kMaxG) ssum[tid] = 0ull;
    stamp(p, 1);
    // generation of the cross-cluster "shared reads done" barrier, read by the thread that may
    // later have to wait on it (lane 0 of the last warp) BEFORE this cluster arrives
    int32_t gen = 0;
    if (needs_epoch && tid == kThreads - 32) gen = ld_acquire_gpu(&st.sync_words[1]);

    const int ipos = p.a.input_pos[0];
    const int pos_row = (st.pos_heads == 1) ? 0 : h;
    const int32_t* posp = st.pos + (size_t)pos_row * L;
    for (int lr = tid; lr < local_rows; lr += kThreads) {
        const int s = slot_of(lr);
        maskb[lr] = (s < L) ? st.mask[(size_t)h * L + s] : 0;
    }

    pdl_launch_dependents();  // the next kernel in the stream may start its own independent prologue

    // ---- 1. L2 only: key_norm.max() over every head and slot (cache.py:602) = X0 exchange
    float l2_max = 0.f;
    if constexpr (POLICY == CC_POLICY_L2) {
        if (do_evict) {
            const uint16_t* kn = reinterpret_cast<const uint16_t*>(st.key_norm);
            float m = -CUDART_INF_F;
            for (int hh = 0; hh < H; ++hh)
                for (int lr = tid; lr < local_rows; lr += kThreads) {
                    const int s = slot_of(lr);
                    if (s < L) m = fmaxf(m, bf16_bits_to_float(kn[(size_t)hh * L + s]));
                }
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
            if (lane == 0) misc[kMiscWarpMax + warp] = m;
            __syncthreads();
            cluster_wait_acquire();
            b0_waited = true;
            if (tid < C) {
                float mm = misc[kMiscWarpMax];
                for (int w = 1; w < kWarps; ++w) mm = fmaxf(mm, misc[kMiscWarpMax + w]);
                st_async_u32(map_to_rank(smem_u32(&x0[rank]), tid), __float_as_uint(mm), map_to_rank(smem_u32(&xbar[0]), tid));
            }
            mbar_wait(&xbar[0], 0);
            l2_max = x0[0];
            for (int r = 1; r < C; ++r) l2_max = fmaxf(l2_max, x0[r]);
        }
    }

    // ---- 2. local arg-min candidate, key = (order(score) << 32) | slot  (first index wins ties)
    uint64_t ckey = kNoKey;
    if (do_evict) {
        if constexpr (POLICY == CC_POLICY_HEAVY_HITTER) {
            const double* num = reinterpret_cast<const double*>(st.attn_history_num) + (size_t)h * L;
            const int32_t* den = st.attn_history_denom + (size_t)h * L;
            for (int lr = tid; lr < local_rows; lr += kThreads) {
                const int s = slot_of(lr);
                if (s >= L) continue;
                const int ps = posp[s];
                // cache.py:727-749: mean accumulated attention; globals & recent -> 1.0; unfilled -> 0.0
                float avg;
                if (!hh_ring) {
                    avg = __fdiv_rn(__double2float_rn(num[s]), (float)max(den[s], 1));
                } else {
                    // sum of the slot's ring (kept as an exact fp64 running sum): bf16 ring -> bf16(fp32 sum),
                    // thresholded (bool) ring -> integer count; / clamp(denom, 1, W)   (cache.py:727-738)
                    const float hs = __double2float_rn(st.attn_history_sum[(size_t)h * L + s]);
                    avg = __fdiv_rn(st.attn_thresholding ? hs : bf16_round(hs), (float)min(max(den[s], 1), st.history_window));
                }
                if (ps < st.global_tokens || ps >= ipos - st.recent_window) avg = 1.0f;
                if (ps == -1) avg = 0.0f;
                const uint64_t k = ((uint64_t)float_order_key(avg) << 32) | (uint32_t)s;
                ckey = k < ckey ? k : ckey;
            }
        } else if constexpr (POLICY == CC_POLICY_L2) {
            const uint16_t* kn = reinterpret_cast<const uint16_t*>(st.key_norm) + (size_t)h * L;
            for (int lr = tid; lr < local_rows; lr += kThreads) {
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
        } else if constexpr (POLICY == CC_POLICY_RECENT_GLOBAL) {
            // cache.py:552-556: oldest position among slots >= g (unfilled = -1 first)
            for (int lr = tid; lr < local_rows; lr += kThreads) {
                const int s = slot_of(lr);
                if (s >= L || s < st.global_tokens) continue;
                const uint64_t k = ((uint64_t)((uint32_t)posp[s] ^ 0x80000000u) << 32) | (uint32_t)s;
                ckey = k < ckey ? k : ckey;
            }
        } else if constexpr (hybrid) {
            if (hyb_mode == kHybEvict) {
                // cache.py:844-894 over the live prefix [0, cts): mean ring attention (or position for
                // window-only heads); first g slots, special, punctuation and the recent window are saved
                const int W = st.history_window;
                const int32_t* den = st.attn_history_denom + (size_t)h * L;
                const double* hsum = st.attn_history_sum + (size_t)h * L;
                for (int lr = tid; lr < local_rows; lr += kThreads) {
                    const int s = slot_of(lr);
                    if (s >= hyb_cts) continue;
                    const int ps = posp[s];
                    float sc;
                    if (hyb_bits & CC_HYB_HEAVY_HITTER) {
                        const float d = (float)min(den[s], W);
                        sc = __fdiv_rn(bf16_round(__double2float_rn(hsum[s])), d);
                    } else {
                        sc = (float)ps;
                    }
                    bool save = s < st.global_tokens;
                    if ((hyb_bits & CC_HYB_SPECIAL) && st.special_mask) save |= st.special_mask[(size_t)h * L + s] != 0;
                    if ((hyb_bits & CC_HYB_PUNC) && st.punc_mask) save |= st.punc_mask[(size_t)h * L + s] != 0;
                    if ((hyb_bits & CC_HYB_WINDOW) && ps > ipos - hyb_rw) save = true;  // strict >
                    if (save) sc = CUDART_INF_F;
                    // torch.argmin treats NaN (0/0) as the minimum
                    const uint32_t hi = (sc != sc) ? 0u : float_order_key(sc);
                    const uint64_t k = ((uint64_t)hi << 32) | (uint32_t)s;
                    ckey = k < ckey ? k : ckey;
                }
            } else {
                // append at min(L-1, cts); a dropped token is parked in slot L-1 (cache.py:973-979)
                const int idx = hyb_mode == kHybAppend ? min(L - 1, hyb_cts) : L - 1;
                if (tid == 0 && ((idx >> chs) & (C - 1)) == rank) ckey = (uint64_t)(uint32_t)idx;
            }
        } else {  // CC_POLICY_EXTERNAL: the caller chose the slot
            const int64_t idx = p.a.fill_idx[st.pos_heads == 1 ? 0 : h];
            if (tid == 0 && idx >= 0 && idx < L && (((int)idx >> chs) & (C - 1)) == rank) ckey = (uint64_t)(uint32_t)idx;
        }
        ckey = warp_min_u64(ckey);
        if (lane == 0) reinterpret_cast<unsigned long long*>(miscu + kMiscWarpKey)[warp] = ckey;
        __syncthreads();
        ckey = lane < kWarps ? reinterpret_cast<unsigned long long*>(miscu + kMiscWarpKey)[lane] : kNoKey;
        ckey = warp_min_u64(ckey);
    }
    stamp(p, 2);
    // L2 warm-up hints: from here on this kernel mostly waits on latencies and HBM is idle; ask the
    // memory system to stream the next projections' weights into L2 meanwhile (4 KiB per request,
    // requests dealt round-robin over all CTAs).  Issued after this kernel's own loads.
    for (int i = 0; i < p.a.n_prefetch; ++i) {
        const char* base = reinterpret_cast<const char*>(p.a.prefetch[i].ptr);
        const unsigned long long bytes = p.a.prefetch[i].bytes;
        const unsigned long long n_req = (bytes + 4095) >> 12;
        for (unsigned long long c = (unsigned long long)tid * gridDim.x + blockIdx.x; c < n_req; c += (unsigned long long)kThreads * gridDim.x) {
            const unsigned long long off = c << 12;
            l2_prefetch_bulk(base + off, (uint32_t)min(4096ull, bytes - off));
        }
    }
    // Everything above reads only this layer's cache and policy state, which no kernel in flight
    // writes (CC_FLAG_OVERLAP_PREDECESSOR); q / k_new / v_new come from the projection kernel launched
    // just before this one: wait for it here, with the K/V tiles already on their way.
    pdl_wait();
    // query rows of this KV group and the new value slice, in registers
    float q[G][8];
    uint4 vnew8 = make_uint4(0, 0, 0, 0);
    if (do_attend) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const uint4 v = *reinterpret_cast<const uint4*>(
                reinterpret_cast<const uint16_t*>(p.a.q) + (size_t)(h * G + g) * D + li * 8);
            unpack8(v, q[g]);
        }
        if (do_evict)
            vnew8 = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.a.v_new) + (size_t)h * D + li * 8);
    }
    // local row of this CTA's candidate (it owns the slot by construction)
    int cand_local = -1;
    if (ckey != kNoKey) {
        const int s = (int)(uint32_t)(ckey & 0xffffffffu);
        cand_local = (((s >> chs) >> csh) << chs) + (s & (CH - 1));
    }
    // a dropped token does not take part in attention (its slot's mask stays as it was)
    const bool attend_new = do_evict && !(hybrid && hyb_mode == kHybDrop);

    // ---- 3. logits over the K ring (canonical dot tree), new-token logits by the last warp
    if (do_attend) {
        if (attend_new && warp == kWarps - 1) {
            float kf[8];
            const uint4 kv = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.a.k_new) + (size_t)h * D + li * 8);
            unpack8(kv, kf);
            float part[G];
#pragma unroll
            for (int g = 0; g < G; ++g) {
                float acc = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) acc = __fmaf_rn(q[g][i], kf[i], acc);
                part[g] = acc;
            }
            int gsel;
            const float dot = group_reduce_multi<LPR, G>(part, li, &gsel);
            const float v = NAIVE ? bf16_round(__fmul_rn(bf16_round(dot), p.scale)) : __fmul_rn(dot, p.scale);
            if (sub == 0 && gsel < G) misc[kMiscNewLg + gsel] = v;  // replicated lanes write the same value
        }
        // rows of chunks that cannot be live
        for (int lr = live_rows + tid; lr < local_rows; lr += kThreads)
#pragma unroll
            for (int g = 0; g < G; ++g) lg[g * stride + lr] = -CUDART_INF_F;
        __syncthreads();  // maskb visible
        for (int j = 0; j < nlc_live; ++j) {
            const int stage = j % NS;
            mbar_wait(&kbar[stage], (j / NS) & 1);
            if (j == 0) stamp(p, 3);
            const int rows_c = chunk_rows_valid(j);
            const uint8_t* kbase = kring + stage * chunk_bytes;
            // up to 4 row pairs per warp per pass, all fetched before the dependent math (ILP)
            constexpr int UR = 4;
            for (int r0 = warp * RPW; r0 < CH; r0 += UR * kWarps * RPW) {
                uint4 kv[UR];
#pragma unroll
                for (int u = 0; u < UR; ++u) {
                    const int r = r0 + u * kWarps * RPW + sub;
                    kv[u] = (r < rows_c) ? *reinterpret_cast<const uint4*>(kbase + r * kRowBytes + li * 16)
                                         : make_uint4(0, 0, 0, 0);
                }
#pragma unroll
                for (int u = 0; u < UR; ++u) {
                    const int r = r0 + u * kWarps * RPW + sub;
                    float kf[8];
                    unpack8(kv[u], kf);
                    float part[G];
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
                    const float v = NAIVE ? bf16_round(__fmul_rn(bf16_round(dot), p.scale)) : __fmul_rn(dot, p.scale);
                    const int lr = (j << chs) + r;
                    if (gsel < G && r < CH) lg[gsel * stride + lr] = (r < rows_c && maskb[lr]) ? v : -CUDART_INF_F;
                }
            }
            __syncthreads();
            if (tid == 0 && j + NS < nlc_live) {
                fence_proxy_async();
                const int jn = j + NS;
                const uint32_t bytes = chunk_rows_valid(jn) * kRowBytes;
                mbar_expect_tx(&kbar[stage], bytes);
                tma_load_1d(kring + stage * chunk_bytes, kc + ((size_t)((jn << csh) + rank) << chs) * D, bytes, &kbar[stage]);
            }
        }
        // ---- 4. row max without the local candidate, and the candidate's own logits -> my X1 record
        for (int g = warp; g < G; g += kWarps) {
            float m = -CUDART_INF_F;
            for (int lr = lane; lr < live_rows; lr += 32)
                if (lr != cand_local) m = fmaxf(m, lg[g * stride + lr]);
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
            if (lane == 0) {
                misc[kMiscMyRec + 2 + g] = m;
                misc[kMiscMyRec + 2 + G + g] = cand_local >= 0 ? lg[g * stride + cand_local] : -CUDART_INF_F;
            }
        }
    } else if (tid < 2 * G) {
        misc[kMiscMyRec + 2 + tid] = -CUDART_INF_F;
    }
    stamp(p, 4);
    // ---- 5. X1: every CTA learns every candidate and every partial max
    if (tid == 0) {
        miscu[kMiscMyRec + 0] = (uint32_t)(ckey & 0xffffffffu);
        miscu[kMiscMyRec + 1] = (uint32_t)(ckey >> 32);
    }
    __syncthreads();
    if (!b0_waited) cluster_wait_acquire();
    for (int idx = tid; idx < C * NW; idx += kThreads) {
        const int dest = idx / NW, w = idx % NW;
        st_async_u32(map_to_rank(smem_u32(&rec[rank * kRecWords + w]), dest), miscu[kMiscMyRec + w],
                     map_to_rank(smem_u32(&xbar[1]), dest));
    }
    mbar_wait(&xbar[1], 0);
    stamp(p, 5);

    // ---- 6. winner + global max per query head (every warp redundantly: no block barrier needed).
    // Lane l looks at record r = l & 15; the two half-warps split the query heads (even / odd).
    const int rl = lane & 15, gh = lane >> 4;
    uint64_t wkey = kNoKey;
    if (rl < C) wkey = (uint64_t)rec[rl * kRecWords] | ((uint64_t)rec[rl * kRecWords + 1] << 32);
    const bool lane_has_cand = wkey != kNoKey;
#pragma unroll
    for (int d = 8; d >= 1; d >>= 1) {
        const uint64_t o = __shfl_xor_sync(0xffffffffu, wkey, d);
        wkey = o < wkey ? o : wkey;
    }
    const bool has_slot = do_evict && wkey != kNoKey;  // a slot is overwritten this step
    const int widx = has_slot ? (int)(uint32_t)(wkey & 0xffffffffu) : -1;
    const int owner = has_slot ? (widx >> chs) & (C - 1) : -1;
    const bool i_own = has_slot && owner == rank;
    const int widx_local = has_slot ? (((widx >> chs) >> csh) << chs) + (widx & (CH - 1)) : -1;
    const bool has_new = has_slot && attend_new;  // ... and the new token attends to itself there
    if (needs_epoch && rank == 0 && tid == 0) red_release_gpu_add(&st.sync_words[0], 1);  // shared reads done
    float gmax[G], newlg[G];
    if (do_attend) {
        constexpr int GH = (G + 1) / 2;  // query heads per half-warp
        float part[GH];
#pragma unroll
        for (int j = 0; j < GH; ++j) {
            const int g = 2 * j + gh;
            float m = -CUDART_INF_F;
            if (rl < C && g < G) {
                m = __uint_as_float(rec[rl * kRecWords + 2 + g]);
                // a candidate that is not overwritten keeps its logit; the overwritten one only if the new
                // token does not replace it in attention (hybrid drop: slot L-1 stays as it was)
                if (lane_has_cand && (rl != owner || !has_new))
                    m = fmaxf(m, __uint_as_float(rec[rl * kRecWords + 2 + G + g]));
            }
            part[j] = m;
        }
#pragma unroll
        for (int d = 8; d >= 1; d >>= 1)
#pragma unroll
            for (int j = 0; j < GH; ++j) part[j] = fmaxf(part[j], __shfl_xor_sync(0xffffffffu, part[j], d));
#pragma unroll
        for (int g = 0; g < G; ++g) {
            newlg[g] = has_new ? misc[kMiscNewLg + g] : -CUDART_INF_F;
            gmax[g] = fmaxf(__shfl_sync(0xffffffffu, part[g / 2], (g & 1) * 16), newlg[g]);
        }
    }
    const int patch_row = (i_own && has_new) ? widx_local : -1;  // this row now holds the new token
    stamp(p, 6);

    if (do_attend) {
        // ---- 8. exp and the exact fixed-point softmax denominator = X2.  Items are (g, row) pairs laid
        // out g-major with the row count padded to a warp multiple, so a warp never straddles two g.
        const int rows_pad = (live_rows + 31) & ~31;
        for (int it = tid; it < G * rows_pad; it += kThreads) {
            const int g = it / rows_pad, lr = it - g * rows_pad;
            long long part = 0;
            if (lr < live_rows) {
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
            if (lane == 0 && part) atomicAdd(&ssum[g], (unsigned long long)part);  // integers: order independent
        }
        __syncthreads();
        stamp(p, 7);
        for (int idx = tid; idx < C * G; idx += kThreads) {
            const int dest = idx / G, g = idx % G;
            st_async_u64(map_to_rank(smem_u32(&sums[rank * kMaxG + g]), dest), ssum[g], map_to_rank(smem_u32(&xbar[2]), dest));
        }
        mbar_wait(&xbar[2], 0);
        stamp(p, 8);
        // ---- 9a. probabilities (every warp sums the C partial denominators itself)
        float sumf[G];
#pragma unroll
        for (int g = 0; g < G; ++g) {
            long long t = lane < C ? (long long)sums[lane * kMaxG + g] : 0ll;
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
            sumf[g] = __fmul_rn(__ll2float_rn(t), 9.094947017729282e-13f);  // 2^-40
        }
        for (int it = tid; it < G * live_rows; it += kThreads) {
            const int g = it / live_rows, lr = it - g * live_rows;
            float sf = sumf[0];
#pragma unroll
            for (int gg = 1; gg < G; ++gg) if (g == gg) sf = sumf[gg];
            float pr = __fdiv_rn(lg[g * stride + lr], sf);
            if (NAIVE) pr = bf16_round(pr);
            lg[g * stride + lr] = pr;
        }
        __syncthreads();
        stamp(p, 9);

        // ---- 10. y = P.V over the V ring, cluster-wide reduction = X3
        float acc[G][8];
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[g][i] = 0.f;
        const int rg = tid / LPR;
        for (int j = 0; j < nlc_live; ++j) {
            const int stage = j % NS;
            mbar_wait(&vbar[stage], (j / NS) & 1);
            const int rows_c = chunk_rows_valid(j);
            const uint8_t* vbase = vring + stage * chunk_bytes;
            for (int r = rg; r < rows_c; r += NRG) {
                const int lr = (j << chs) + r;
                uint4 vv = *reinterpret_cast<const uint4*>(vbase + r * kRowBytes + li * 16);
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
            if (j + NS < nlc_live) {  // ring slot is refilled: everyone must be done with it
                __syncthreads();
                if (tid == 0) {
                    fence_proxy_async();
                    const int jn = j + NS;
                    const uint32_t bytes = chunk_rows_valid(jn) * kRowBytes;
                    mbar_expect_tx(&vbar[stage], bytes);
                    tma_load_1d(vring + stage * chunk_bytes, vc + ((size_t)((jn << csh) + rank) << chs) * D, bytes, &vbar[stage]);
                }
            }
        }
        stamp(p, 10);
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float v = acc[g][i];
#pragma unroll
                for (int d = 16; d >= LPR; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
                acc[g][i] = v;
            }
        if (lane < LPR) {
#pragma unroll
            for (int g = 0; g < G; ++g) {
                float4* dst = reinterpret_cast<float4*>(&red[warp * G * D + g * D + lane * 8]);
                dst[0] = make_float4(acc[g][0], acc[g][1], acc[g][2], acc[g][3]);
                dst[1] = make_float4(acc[g][4], acc[g][5], acc[g][6], acc[g][7]);
            }
        }
        __syncthreads();
        const int SL = (G * D) >> csh;  // outputs reduced by each CTA of the cluster (a multiple of 4)
        for (int o4 = tid; o4 < G * D / 4; o4 += kThreads) {
            const int o = o4 * 4;
            float4 s = *reinterpret_cast<const float4*>(&red[o]);
#pragma unroll
            for (int w = 1; w < kWarps; ++w) {
                const float4 t = *reinterpret_cast<const float4*>(&red[w * G * D + o]);
                s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
            }
            const int dest = o / SL, off = o - dest * SL;
            st_async_v4f32(map_to_rank(smem_u32(&yx[rank * SL + off]), dest), s, map_to_rank(smem_u32(&xbar[3]), dest));
        }
        stamp(p, 11);
    }

    // ======== off the critical path: global state writes, overlapped with the X3 barrier ========
    // ---- 7. owner CTA: in-place slot overwrite (cache.py:390-401 / 460-490) + policy state
    if (i_own && warp == kWarps - 1) {
        if (lane < LPR) {
            const uint16_t* ks = reinterpret_cast<const uint16_t*>(p.a.k_new) + (size_t)h * D + lane * 8;
            const uint16_t* vs = reinterpret_cast<const uint16_t*>(p.a.v_new) + (size_t)h * D + lane * 8;
            uint16_t* kd = reinterpret_cast<uint16_t*>(st.k_cache) + ((size_t)h * L + widx) * D + lane * 8;
            uint16_t* vd = reinterpret_cast<uint16_t*>(st.v_cache) + ((size_t)h * L + widx) * D + lane * 8;
            *reinterpret_cast<uint4*>(kd) = *reinterpret_cast<const uint4*>(ks);
            *reinterpret_cast<uint4*>(vd) = *reinterpret_cast<const uint4*>(vs);
        }
        float knorm = 0.f;
        if constexpr (POLICY == CC_POLICY_L2) {
            // key_norm[idx] = ||k_new||_2 (cache.py:592-593), canonical tree, correctly rounded sqrt
            float kf[8];
            const uint4 kv = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.a.k_new) + (size_t)h * D + li * 8);
            unpack8(kv, kf);
            float acc = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) acc = __fmaf_rn(kf[i], kf[i], acc);
            knorm = __fsqrt_rn(group_butterfly<LPR>(acc));
        }
        if (lane == 0) {
            const int32_t pos_old = posp[widx];
            if (!hybrid || hyb_mode == kHybAppend) st.mask[(size_t)h * L + widx] = 1;
            if (p.a.fill_idx && POLICY != CC_POLICY_EXTERNAL) p.a.fill_idx[h] = widx;
            if (POLICY == CC_POLICY_HEAVY_HITTER && !do_score) {
                if (hh_ring) st.attn_history_sum[(size_t)h * L + widx] = 0.0;   // the ring row is cleared below
                else reinterpret_cast<double*>(st.attn_history_num)[(size_t)h * L + widx] = 0.0;  // cache.py:753-763
                st.attn_history_denom[(size_t)h * L + widx] = 0;
            }
            if (hybrid && is_punc && st.punc_mask) st.punc_mask[(size_t)h * L + widx] = 1;  // cache.py:1011-1016
            if (needs_epoch) {  // head-constant pos, key_norm, counters are read by every cluster: write after all have read
                const int32_t target = (gen + 1) * H;
                while (ld_acquire_gpu(&st.sync_words[0]) - target < 0) {}
            }
            if constexpr (POLICY == CC_POLICY_L2)
                reinterpret_cast<__nv_bfloat16*>(st.key_norm)[(size_t)h * L + widx] = __float2bfloat16_rn(knorm);
            if (st.pos_heads != 1 || h == 0) st.pos[(size_t)pos_row * L + widx] = ipos;
            if constexpr (hybrid) {
                if (hyb_mode == kHybAppend) st.cache_cts[h] = hyb_cts + 1;
                if (h == 0 && is_punc && st.num_punc) st.num_punc[0] += 1;
            } else if (h == 0) {
                st.cache_cts[0] += (pos_old == -1) ? 1 : 0;  // cache.py:330, 356-360
            }
        }
    }
    if (hh_ring && i_own) {
        // the evicted slot's history restarts (cache.py:753-763): clear its ring row, except the column this very
        // step writes when the score update is fused into the same launch
        const int W = st.history_window;
        const int keep = (do_score && NAIVE) ? (int)(hyb_counter % W) : -1;
        if (st.attn_thresholding) {
            uint8_t* rowp = reinterpret_cast<uint8_t*>(st.attn_history_num) + ((size_t)h * L + widx) * W;
            for (int w = tid; w < W; w += kThreads) if (w != keep) rowp[w] = 0;
        } else {
            uint16_t* rowp = reinterpret_cast<uint16_t*>(st.attn_history_num) + ((size_t)h * L + widx) * W;
            for (int w = tid; w < W; w += kThreads) if (w != keep) rowp[w] = 0;
        }
    }
    if (needs_epoch && h == 0 && rank == 0 && tid == 32) {
        // generation bump (and the shared step counter): only after every cluster has arrived
        const int32_t g0 = ld_acquire_gpu(&st.sync_words[1]);
        const int32_t target = (g0 + 1) * H;
        while (ld_acquire_gpu(&st.sync_words[0]) - target < 0) {}
        if ((hybrid || hh_ring) && do_score && NAIVE) st.attn_counter[0] = hyb_counter + 1;
        st_release_gpu(&st.sync_words[1], g0 + 1);
    }

    // ---- 9b. GQA pooling (model.py:413-418) and the score update (cache.py:716-723)
    {
        const bool hh_score = do_score && NAIVE && POLICY == CC_POLICY_HEAVY_HITTER && !hh_ring;
        const bool hhr_score = do_score && NAIVE && hh_ring;
        const bool ring_score = do_score && NAIVE && hybrid;
        const int W = st.history_window;
        const int wslot = (ring_score || hhr_score) ? (int)(hyb_counter % W) : 0;
        // thresholded history: indicator p >= 1 / cache_cts with cache_cts AFTER this step's insertion (cache.py:330,
        // 708).  While the cache is filling the inserted slot is an unfilled one (score 0.0, first index): + 1.
        const float thr = hhr_score ? __fdiv_rn(1.0f, (float)(hh_cts + ((do_evict && has_slot && hh_cts < L) ? 1 : 0))) : 0.f;
        __nv_bfloat16* probs = reinterpret_cast<__nv_bfloat16*>(p.a.probs);
        if (NAIVE && (hh_score || hhr_score || ring_score || (do_attend && probs))) {
            for (int lr = tid; lr < local_rows; lr += kThreads) {
                const int s = slot_of(lr);
                if (s >= L) continue;
                float pooled = 0.f;
                if (do_attend) {
                    if (lr < live_rows) {
                        float acc = lg[lr];
#pragma unroll
                        for (int g = 1; g < G; ++g) acc = __fadd_rn(acc, lg[g * stride + lr]);
                        pooled = bf16_round(__fdiv_rn(acc, (float)G));  // torch mean: fp32 sum / G, then bf16
                    }
                    if (probs) probs[(size_t)h * L + s] = __float2bfloat16_rn(pooled);
                } else {
                    pooled = __bfloat162float(probs[(size_t)h * L + s]);
                }
                if (hh_score) {
                    double* num = reinterpret_cast<double*>(st.attn_history_num) + (size_t)h * L;
                    int32_t* den = st.attn_history_denom + (size_t)h * L;
                    const bool fresh = has_slot && s == widx;  // evicted slot's history restarts (cache.py:753-763)
                    num[s] = (fresh ? 0.0 : num[s]) + (double)pooled;
                    den[s] = (fresh ? 0 : den[s]) + 1;
                } else if (hhr_score) {
                    // ring[s][counter % W] = pooled (or its indicator), denominators + 1; the evicted slot starts
                    // from a cleared ring (cache.py:716-723, 753-763).  hsum = exact running sum of the ring.
                    int32_t* den = st.attn_history_denom + (size_t)h * L;
                    double* hs = st.attn_history_sum + (size_t)h * L + s;
                    const bool fresh = has_slot && s == widx;
                    den[s] = (fresh ? 0 : den[s]) + 1;
                    if (st.attn_thresholding) {
                        uint8_t* cell = reinterpret_cast<uint8_t*>(st.attn_history_num) + ((size_t)h * L + s) * W + wslot;
                        const float v = pooled >= thr ? 1.f : 0.f;
                        const float old = fresh ? 0.f : (float)(*cell != 0);
                        *hs = ((fresh ? 0.0 : *hs) - (double)old) + (double)v;
                        *cell = (uint8_t)(v != 0.f);
                    } else {
                        __nv_bfloat16* cell = reinterpret_cast<__nv_bfloat16*>(st.attn_history_num) + ((size_t)h * L + s) * W + wslot;
                        const float old = fresh ? 0.f : __bfloat162float(*cell);
                        *hs = ((fresh ? 0.0 : *hs) - (double)old) + (double)pooled;
                        *cell = __float2bfloat16_rn(pooled);
                    }
                } else if (ring_score) {
                    // ring[s][counter % W] = pooled; the evicted slot keeps its ring (the reference's reset is a
                    // silent no-op, cache.py:993-996).  Only live slots can hold non-zero history.
                    int32_t* den = st.attn_history_denom + (size_t)h * L;
                    den[s] += 1;
                    if (s < live_limit || pooled != 0.f) {
                        __nv_bfloat16* cell = reinterpret_cast<__nv_bfloat16*>(st.attn_history_num) + ((size_t)h * L + s) * W + wslot;
                        const float old = __bfloat162float(*cell);
                        if (old != pooled) {
                            double* hs = st.attn_history_sum + (size_t)h * L + s;
                            *hs = (*hs - (double)old) + (double)pooled;
                            *cell = __float2bfloat16_rn(pooled);
                        }
                    }
                }
            }
        }
        if (hh_score && h == 0 && rank == 0 && tid == 0) st.attn_counter[0] += 1;
    }

    if (do_attend) {
        mbar_wait(&xbar[3], 0);
        stamp(p, 12);
        const int SL = (G * D) >> csh;
        __nv_bfloat16* y = reinterpret_cast<__nv_bfloat16*>(p.a.y) + (size_t)h * G * D;
        for (int t = tid; t < SL; t += kThreads) {
            float s = yx[t];
            for (int r = 1; r < C; ++r) s += yx[r * SL + t];
            y[rank * SL + t] = __float2bfloat16_rn(s);
        }
    }
    stamp(p, 13);
}

// --------------------------------------------------------------------------- host side

using KernelFn = void (*)(const DecodeParams);

template <int D, int G>
static KernelFn pick_policy(int policy, bool naive) {
    switch (policy) {
    case CC_POLICY_EXTERNAL: return naive ? (KernelFn)cc_decode_kernel<D, G, CC_POLICY_EXTERNAL, true>   // debug_* full cache
                                          : (KernelFn)cc_decode_kernel<D, G, CC_POLICY_EXTERNAL, false>;
    case CC_POLICY_RECENT_GLOBAL: return naive ? nullptr : (KernelFn)cc_decode_kernel<D, G, CC_POLICY_RECENT_GLOBAL, false>;
    case CC_POLICY_L2: return naive ? nullptr : (KernelFn)cc_decode_kernel<D, G, CC_POLICY_L2, false>;
    case CC_POLICY_HEAVY_HITTER: return naive ? (KernelFn)cc_decode_kernel<D, G, CC_POLICY_HEAVY_HITTER, true> : nullptr;
    case CC_POLICY_HYBRID: return naive ? (KernelFn)cc_decode_kernel<D, G, CC_POLICY_HYBRID, true>
                                        : (KernelFn)cc_decode_kernel<D, G, CC_POLICY_HYBRID, false>;
    default: return nullptr;
    }
}
template <int D>
static KernelFn pick_group(int G, int policy, bool naive) {
    switch (G) {
    case 1: return pick_policy<D, 1>(policy, naive);
    case 2: return pick_policy<D, 2>(policy, naive);
    case 4: return pick_policy<D, 4>(policy, naive);
    case 6: return pick_policy<D, 6>(policy, naive);
    case 7: return pick_policy<D, 7>(policy, naive);
    case 8: return pick_policy<D, 8>(policy, naive);
    default: return nullptr;
    }
}
static KernelFn pick_kernel(int D, int G, int policy, bool naive) {
    if (D == 128) return pick_group<128>(G, policy, naive);
    if (D == 64) return pick_group<64>(G, policy, naive);
    return nullptr;
}

struct LaunchPlan {
    KernelFn fn;
    int cluster, ch_shift, max_local_chunks, n_stages, lg_stride;
    uint32_t smem;
};

static int g_num_sms = 0;
static int g_smem_optin = 0;
static unsigned long long* g_trace = nullptr;  // diagnostics only (cc_debug_set_trace)
void set_trace_buffer(unsigned long long* buf) { g_trace = buf; }

static int g_limits_dev = -1;
static int device_limits() {
    int dev = 0;
    cudaGetDevice(&dev);
    if (g_num_sms && g_limits_dev == dev) return CC_OK;
    g_num_sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return set_error(CC_ERR_NO_DEVICE, "no CUDA device");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return set_error(CC_ERR_NO_DEVICE, "cudaGetDeviceProperties failed");
    if (prop.major != 10) return set_error(CC_ERR_NO_DEVICE, "this library is built for sm_100a (Blackwell B200) only");
    g_num_sms = prop.multiProcessorCount;
    g_smem_optin = (int)prop.sharedMemPerBlockOptin;
    g_limits_dev = dev;
    return CC_OK;
}

// Plans are cached per (shape, policy class): the occupancy query is not free.
struct PlanKey { int D, G, naive, H, L, policy, dev; };
struct PlanEntry { PlanKey k; LaunchPlan p; };
static PlanEntry g_plans[128];
static int g_nplans = 0;

static int make_plan(const CcCacheState* st, bool needs_epoch, LaunchPlan* out) {
    const int D = st->head_dim, G = st->q_per_kv, H = st->n_kv_heads, L = st->max_cache_length;
    const bool naive = st->attn_mode == CC_ATTN_NAIVE_BF16;
    int dev = 0;
    cudaGetDevice(&dev);   // plans (occupancy, shared-memory limits) belong to a device
    for (int i = 0; i < g_nplans; ++i) {
        const PlanKey& k = g_plans[i].k;
        if (k.D == D && k.G == G && k.naive == (int)naive && k.H == H && k.L == L && k.policy == st->policy && k.dev == dev) {
            *out = g_plans[i].p;
            return CC_OK;
        }
    }
    int rc = device_limits();
    if (rc) return rc;
    KernelFn fn = pick_kernel(D, G, st->policy, naive);
    if (!fn) return set_error(CC_ERR_UNSUPPORTED_SHAPE, "unsupported head_dim (64/128), q_per_kv (1,2,4,6,7,8) or policy/attention-mode pair");
    cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    // the cap is per function and plans for other shapes share it: raise it once to the device
    // limit instead of to this plan's size (a later, smaller plan must not lower it)
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem_optin) != cudaSuccess) {
        cudaGetLastError();
        return set_error(CC_ERR_CUDA, "cannot raise the dynamic shared memory limit");
    }
    // chunk size: the largest of 64/32/16 rows that still gives every CTA of a 16-cluster a chunk
    int chs = 6;
    while (chs > 4 && ((L + (1 << chs) - 1) >> chs) < kMaxCluster) --chs;
    const int n_chunks = (L + (1 << chs) - 1) >> chs;
    int c_max = kMaxCluster;
    if (const char* e = getenv("CC_B200_MAX_CLUSTER")) c_max = atoi(e) > 0 ? atoi(e) : c_max;  // experiments only
    for (int C = c_max; C >= 1; C >>= 1) {
        if (C > 1 && (H * C > g_num_sms || C > n_chunks)) continue;
        const int mlc = (n_chunks + C - 1) / C;
        const int stride = ((mlc << chs) + 31) / 32 * 32;
        for (int ns = mlc < 4 ? mlc : 4; ns >= 1; --ns) {
            const SmemLayout lay = make_layout(D, G, 1 << chs, mlc << chs, ns, stride);
            if ((int)lay.total > g_smem_optin) continue;
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(H * C);
            cfg.blockDim = dim3(kThreads);
            cfg.dynamicSmemBytes = lay.total;
            cudaLaunchAttribute attr;
            attr.id = cudaLaunchAttributeClusterDimension;
            attr.val.clusterDim.x = C; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
            cfg.attrs = &attr; cfg.numAttrs = 1;
            int max_clusters = 0;
            if (cudaOccupancyMaxActiveClusters(&max_clusters, fn, &cfg) != cudaSuccess) { cudaGetLastError(); continue; }
            // every cluster of the layer must be co-resident: the epoch barrier needs it for correctness,
            // and a second wave of clusters would double the step time in any case
            if (getenv("CC_B200_DEBUG_PLAN"))
                fprintf(stderr, "[cc plan] D=%d G=%d policy=%d L=%d: C=%d stages=%d smem=%u max_active_clusters=%d\n", D, G,
                        st->policy, L, C, ns, lay.total, max_clusters);
            // (the epoch barrier spins on the other heads' clusters: without co-residency it would never complete,
            // whatever the cluster size)
            if (max_clusters < H && (C > 1 || needs_epoch)) continue;
            LaunchPlan p = {fn, C, chs, mlc, ns, stride, lay.total};
            if (g_nplans < 128) g_plans[g_nplans++] = {{D, G, (int)naive, H, L, st->policy, dev}, p};
            *out = p;
            return CC_OK;
        }
    }
    return set_error(CC_ERR_UNSUPPORTED_SHAPE, "no cluster size / shared-memory plan fits this cache shape");
}

static int validate(const CcCacheState* st, const CcDecodeArgs* a) {
    if (!st || !a) return set_error(CC_ERR_INVALID_ARG, "null state or args");
    if (st->n_kv_heads < 1 || st->q_per_kv < 1 || st->max_cache_length < 1) return set_error(CC_ERR_INVALID_ARG, "bad shape");
    if (st->policy < CC_POLICY_EXTERNAL || st->policy > CC_POLICY_HYBRID) return set_error(CC_ERR_INVALID_ARG, "bad policy");
    if (!st->k_cache || !st->v_cache || !st->pos || !st->mask || !st->cache_cts) return set_error(CC_ERR_INVALID_ARG, "null cache buffer");
    if (st->pos_heads != 1 && st->pos_heads != st->n_kv_heads) return set_error(CC_ERR_INVALID_ARG, "pos_heads must be 1 or n_kv_heads");
    const int phases = a->phases;
    if (phases < 1 || phases > CC_PHASE_ALL) return set_error(CC_ERR_INVALID_ARG, "bad phases mask");
    if (!a->input_pos) return set_error(CC_ERR_INVALID_ARG, "null input_pos");
    if ((phases & CC_PHASE_EVICT_FILL) && (!a->k_new || !a->v_new)) return set_error(CC_ERR_INVALID_ARG, "evict/fill needs k_new and v_new");
    if ((phases & CC_PHASE_ATTEND) && (!a->q || !a->y)) return set_error(CC_ERR_INVALID_ARG, "attend needs q and y");
    if (st->policy == CC_POLICY_EXTERNAL && (phases & CC_PHASE_EVICT_FILL) && !a->fill_idx)
        return set_error(CC_ERR_INVALID_ARG, "external policy needs fill_idx");
    const bool naive = st->attn_mode == CC_ATTN_NAIVE_BF16;
    if (naive && (phases & CC_PHASE_SCORE) && !(phases & CC_PHASE_ATTEND) && !a->probs)
        return set_error(CC_ERR_INVALID_ARG, "score-only phase needs probs");
    if (st->policy == CC_POLICY_HEAVY_HITTER) {
        if (st->history_window < 1) return set_error(CC_ERR_INVALID_ARG, "history_window must be >= 1");
        if (!st->attn_history_num || !st->attn_history_denom || !st->attn_counter) return set_error(CC_ERR_INVALID_ARG, "null heavy_hitter state");
        if (st->history_window > 1 && !st->attn_history_sum)
            return set_error(CC_ERR_INVALID_ARG, "heavy_hitter with a history window needs the running ring sums (attn_history_sum)");
        if (st->attn_thresholding && st->history_window == 1)
            return set_error(CC_ERR_INVALID_ARG, "attn_thresholding needs history_window_size > 1 (the reference cannot add into a bool history)");
        if (!naive) return set_error(CC_ERR_INVALID_ARG, "heavy_hitter needs the naive (probability-returning) attention mode");
    }
    if (st->policy == CC_POLICY_HYBRID) {
        if (st->cts_heads != st->n_kv_heads) return set_error(CC_ERR_INVALID_ARG, "hybrid needs per-head cache_cts");
        if (st->pos_heads != st->n_kv_heads) return set_error(CC_ERR_INVALID_ARG, "hybrid needs per-head pos");
        if (!st->hyb_strategy_bits || !st->hyb_recent_window || !st->hyb_heavy_budget) return set_error(CC_ERR_INVALID_ARG, "null hybrid strategy tables");
        if (naive && (st->history_window < 2 || !st->attn_history_num || !st->attn_history_denom || !st->attn_counter || !st->attn_history_sum))
            return set_error(CC_ERR_INVALID_ARG, "hybrid with heavy hitters needs the ring history, its denominators, counter and running sums");
        if (!st->sync_words) return set_error(CC_ERR_INVALID_ARG, "hybrid needs sync_words");
    }
    if (st->policy == CC_POLICY_L2 && !st->key_norm) return set_error(CC_ERR_INVALID_ARG, "null key_norm");
    if (a->n_prefetch < 0 || a->n_prefetch > 3) return set_error(CC_ERR_INVALID_ARG, "n_prefetch must be 0..3");
    for (int i = 0; i < a->n_prefetch; ++i)
        if (!a->prefetch[i].ptr || ((uintptr_t)a->prefetch[i].ptr & 15) || (a->prefetch[i].bytes & 15))
            return set_error(CC_ERR_INVALID_ARG, "prefetch ranges must be non-null and 16-byte aligned/sized");
    if ((st->policy == CC_POLICY_RECENT_GLOBAL || st->policy == CC_POLICY_L2) && !st->sync_words)
        return set_error(CC_ERR_INVALID_ARG, "this policy needs sync_words");
    if (st->policy == CC_POLICY_RECENT_GLOBAL && st->global_tokens >= st->max_cache_length)
        return set_error(CC_ERR_INVALID_ARG, "global_tokens must be < max_cache_length");
    return CC_OK;
}

static bool plan_needs_epoch(const CcCacheState* st) {
    return st->policy == CC_POLICY_RECENT_GLOBAL || st->policy == CC_POLICY_L2 || st->policy == CC_POLICY_HYBRID ||
           (st->policy == CC_POLICY_HEAVY_HITTER && st->history_window > 1);
}

int decode_step_impl(const CcCacheState* st, const CcDecodeArgs* a, cudaStream_t stream) {
    int rc = validate(st, a);
    if (rc) return rc;
    if (a->flags & CC_FLAG_VIA_TOKEN_KERNEL) return decode_step_via_token_impl(st, a, stream);
    LaunchPlan plan;
    rc = make_plan(st, plan_needs_epoch(st), &plan);
    if (rc) return rc;
    DecodeParams p;
    p.st = *st;
    p.a = *a;
    p.ch_shift = plan.ch_shift;
    p.max_local_chunks = plan.max_local_chunks;
    p.n_stages = plan.n_stages;
    p.lg_stride = plan.lg_stride;
    p.scale = (float)(1.0 / sqrt((double)st->head_dim));
    p.lay = make_layout(st->head_dim, st->q_per_kv, 1 << plan.ch_shift, plan.max_local_chunks << plan.ch_shift,
                        plan.n_stages, plan.lg_stride);
    p.c_shift = 0;
    while ((1 << p.c_shift) < plan.cluster) ++p.c_shift;
    p.trace = g_trace;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(st->n_kv_heads * plan.cluster);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = plan.smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = plan.cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl_enabled() && (a->flags & CC_FLAG_OVERLAP_PREDECESSOR)) ? 2 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, plan.fn, p);
    if (e != cudaSuccess) return set_cuda_error(e, "cc_decode_step launch");
    return CC_OK;
}

int decode_cluster_size_impl(const CcCacheState* st) {
    if (!st) return set_error(CC_ERR_INVALID_ARG, "null state");
    LaunchPlan plan;
    int rc = make_plan(st, plan_needs_epoch(st), &plan);
    return rc ? rc : plan.cluster;
}

}  // namespace cc
