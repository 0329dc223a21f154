// prefill_attn.cu -- causal prompt attention on the 5th-gen tensor cores (tcgen05 + TMEM), with the
// reductions of the GQA-pooled probability map that the prompt compressors and the score-keeping caches
// consume (SnapKV observation-window mean, heavy-hitter column sums, FastGen window bands) produced in
// the same pass, so the [Hq, S, S] map the reference materialises (attention_utils.py:36-54: 4 GiB per
// layer at S = 8192) never exists.
//
// Replaces, for the prefill branch of Attention.forward (model.py:399-427):
//   scaled_dot_product_attention(q, k, v, mask, return_attn=True)      attention_utils.py:36-54
//   attn.view(B, Hkv, G, S, S).mean(2)                                 model.py:413-418
//   attn[:, :, -16:, :].mean(2)  /  attn.sum(2)                        prompt_compression.py:175, 192; cache.py:704
//   band sums of build_masks' windows                                  cache.py:1066-1173
//
// One CTA = one KV head x one block of QB = 128/G query positions, all G query heads of the group
// stacked into the M = 128 rows of the MMA (row = g * QB + i), so a K/V tile is read once per group.
// Two sweeps over the key tiles (BN = 128 keys):
//   sweep 1  S = Q K^T (tcgen05.mma, fp32 in TMEM) -> row max and sum of exp      (softmax statistics)
//   sweep 2  S again -> P = bf16(exp(l - max) / sum) exactly as the reference rounds it -> shared memory
//            -> Y += P V (tcgen05.mma); the same P tile is read column-wise for the pooled reductions.
// The reference's rounding points are kept (NAIVE: bf16 after q.k, after *scale, after softmax, after the
// GQA mean; MATH: fp32 logits, one rounding of P for the tensor core).
// Operands are staged by TMA (cp.async.bulk.tensor, 128-byte swizzle) into a double-buffered ring; the S
// accumulator is double-buffered in TMEM so the next tile's Q K^T runs under the current tile's softmax.
// Per-tile partial column sums go to a scratch buffer and are added in a fixed order by a second small
// kernel: results are bit-reproducible run to run (no floating-point atomics).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdint.h>

#include "common.cuh"
#include "internal.h"

namespace cc {
namespace {

constexpr int BM = 128;                 // MMA rows of a query tile (G * QB, padded)
constexpr int BN = 128;                 // keys per tile visit of a CTA ...
constexpr int BNH = 64;                 // ... split between the two key-half groups of 8 warps
constexpr int ATOM_BYTES = 128 * 128;   // one 128-byte-swizzled sub-tile: 128 rows x 64 bf16
constexpr int NTHREADS = 512;          // 16 warps: warp w owns TMEM lane quarter w % 4 and column quarter w / 4
constexpr int MAX_WINDOWS = 4;
constexpr int TMEM_COLS = 512;          // group g: S[b] at 128 g + 64 b, Y at 256 + 128 g
constexpr float LOG2E = 1.4426950408889634f;

struct PrefillParams {
    __nv_bfloat16* y;
    long long y_stride_s, y_stride_h;   // elements
    float* part_col;                    // [Hkv][n_qtiles][S_pad]
    float* part_obs;                    // [Hkv][n_obs_tiles][S_pad]
    float* part_band;                   // [n_windows][Hkv][n_qtiles][S_pad]
    int S, S_pad, Hkv, G, QB, n_qtiles;
    int obs_start;                      // first query row of the observation window (S = none)
    int first_obs_tile, n_obs_tiles, n_windows;
    int windows[MAX_WINDOWS];
    float scale;
    int stats;
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ void tma_load_3d(uint32_t smem_dst, const CUtensorMap* tm, int c0, int c1, int c2,
                                            uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        :: "r"(smem_dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void mbar_init_u32(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx_u32(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_u32(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor, 128-byte swizzle (layout type 2), descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: bf16 x bf16 -> fp32, A K-major, B K-major or MN-major
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((b_mn_major ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns of this warp's TMEM lane quarter -> registers (and wait)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        "tcgen05.wait::ld.sync.aligned;\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ float ex2(float x) {
    float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
}
// two fp32 -> packed bf16x2 (round to nearest even), and back
__device__ __forceinline__ uint32_t cvt_bf16x2(float lo, float hi) {
    uint32_t d; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo)); return d;
}
// fp32 value of bf16(x): the high half of the packed pair {x, 0}
__device__ __forceinline__ float round_bf16(float x) {
    uint32_t d; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(x), "f"(0.0f)); return __uint_as_float(d);
}
__device__ __forceinline__ float bf16x2_lo(uint32_t d) { return __uint_as_float(d << 16); }
__device__ __forceinline__ float bf16x2_hi(uint32_t d) { return __uint_as_float(d & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&p);
}

// Sum v[0..7] over the 32 lanes of a warp; every lane returns the total of column
// 4 * bit4(lane) + 2 * bit3(lane) + bit2(lane)  (transpose-reduce: 4 + 2 + 1 + 1 + 1 shuffles).
__device__ __forceinline__ float warp_reduce8(const float (&v)[8], int lane) {
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
    float w[4], x[2];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float keep = b4 ? v[c + 4] : v[c], send = b4 ? v[c] : v[c + 4];
        w[c] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const float keep = b3 ? w[c + 2] : w[c], send = b3 ? w[c] : w[c + 2];
        x[c] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    float y = (b2 ? x[1] : x[0]) + __shfl_xor_sync(0xffffffffu, b2 ? x[0] : x[1], 4);
    y += __shfl_xor_sync(0xffffffffu, y, 2);
    y += __shfl_xor_sync(0xffffffffu, y, 1);
    return y;
}
__device__ __forceinline__ int warp_reduce8_column(int lane) { return ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1); }

// Reductions of the GQA-pooled probabilities of one P tile (128 rows = (head g, query i) x 64 keys, bf16, the
// swizzled A-operand layout in shared memory).  Warp `ck` (0..7) owns the 8 key columns of 16-byte chunk ck; lane l reads
// rows l, l + 32, l + 64, l + 96 (one LDS.128 each), pools over the heads (model.py:413-418), and the warp adds the
// pooled values over the query rows.  QBL = log2(query rows per tile); row = g * QB + i.
template <int QBL>
__device__ __forceinline__ void tile_stats(const uint8_t* sP_ptr, float (*s_comb)[BNH], const PrefillParams& p, int ck,
                                           int lane, int i0, int jbase, bool want_obs) {
    constexpr int QB = 1 << QBL;
    constexpr int NI = QB >= 32 ? QB / 32 : 1;   // query rows held by one lane
    constexpr int NK = 4 / NI;                   // heads per query row held by one lane
    const uint8_t* base = sP_ptr + lane * 128 + ((ck ^ (lane & 7)) << 4);
    float col[8], obs[8], band[MAX_WINDOWS][8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        col[c] = 0.f; obs[c] = 0.f;
#pragma unroll
        for (int w = 0; w < MAX_WINDOWS; ++w) band[w][c] = 0.f;
    }
    const float fG = (float)p.G, invG = 1.0f / (float)p.G;
    const bool g_pow2 = (p.G & (p.G - 1)) == 0;
#pragma unroll
    for (int ii = 0; ii < NI; ++ii) {
        float sum[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) sum[c] = 0.f;
#pragma unroll
        for (int kk = 0; kk < NK; ++kk) {
            const int k = ii + NI * kk;                                  // row = lane + 32 k (same swizzle phase)
            const int g = QB >= 32 ? kk : 2 * kk + (lane >> 4);          // QB = 16: rows l + 32 k hold heads 2k, 2k + 1
            const uint4 u = *reinterpret_cast<const uint4*>(base + k * 32 * 128);
            if (g < p.G) {
                sum[0] += bf16x2_lo(u.x); sum[1] += bf16x2_hi(u.x); sum[2] += bf16x2_lo(u.y); sum[3] += bf16x2_hi(u.y);
                sum[4] += bf16x2_lo(u.z); sum[5] += bf16x2_hi(u.z); sum[6] += bf16x2_lo(u.w); sum[7] += bf16x2_hi(u.w);
            }
        }
        if (QB < 32) {   // the other half of the heads sits 16 lanes away
#pragma unroll
            for (int c = 0; c < 8; ++c) sum[c] += __shfl_xor_sync(0xffffffffu, sum[c], 16);
        }
        const int i = QB >= 32 ? lane + 32 * ii : (lane & (QB - 1));
        const int q_i = i0 + i;
        const bool counted = (q_i < p.S) && (QB >= 32 || lane < QB);   // QB = 16: lanes 16..31 hold duplicates
        float pooled[8];
#pragma unroll
        for (int c = 0; c < 8; c += 2) {
            const float a = g_pow2 ? sum[c] * invG : __fdiv_rn(sum[c], fG);
            const float b = g_pow2 ? sum[c + 1] * invG : __fdiv_rn(sum[c + 1], fG);
            const uint32_t d = cvt_bf16x2(a, b);
            pooled[c] = counted ? bf16x2_lo(d) : 0.f;
            pooled[c + 1] = counted ? bf16x2_hi(d) : 0.f;
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) col[c] += pooled[c];
        if (want_obs && q_i >= p.obs_start) {
#pragma unroll
            for (int c = 0; c < 8; ++c) obs[c] += pooled[c];
        }
        if (p.n_windows) {
#pragma unroll
            for (int w = 0; w < MAX_WINDOWS; ++w) {
                if (w < p.n_windows) {
                    const int lim = q_i - jbase - p.windows[w];   // column c is inside the band iff c > lim
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        if (c > lim) band[w][c] += pooled[c];
                }
            }
        }
    }
    const int oc = ck * 8 + warp_reduce8_column(lane);
    const bool writer = (lane & 3) == 0;
    float r = warp_reduce8(col, lane);
    if (writer) s_comb[0][oc] = r;
    if (want_obs) {
        r = warp_reduce8(obs, lane);
        if (writer) s_comb[1][oc] = r;
    }
#pragma unroll
    for (int w = 0; w < MAX_WINDOWS; ++w) {
        if (w < p.n_windows) {
            r = warp_reduce8(band[w], lane);
            if (writer) s_comb[2 + w][oc] = r;
        }
    }
}

// ------------------------------------------------------------------ the kernel
// grid = n_qtiles * Hkv CTAs (longest query tiles first), 512 threads = two independent KEY-HALF GROUPS of 8 warps.
// Group g walks the 64-key blocks [128 t + 64 g, +64) of the keys this query block can see, with its own K/V ring,
// S accumulators, P tile, Y accumulator and barriers, so the two groups drift out of phase and one group's
// MUFU-bound softmax overlaps the other's rounding / store / reduction phases.  In a group, thread
// (row = t % 128, ch = t / 128) owns keys [32 ch, 32 ch + 32) of its row of every S block (TMEM lane quarter row / 32).
// The partial accumulators Y_0 + Y_1 are added in the epilogue (fixed order: bit-reproducible).
__device__ __forceinline__ void group_sync(int grp) { asm volatile("bar.sync %0, 256;" :: "r"(grp + 1) : "memory"); }

template <int D, bool NAIVE>
__global__ void __launch_bounds__(NTHREADS, 1)
prefill_attn_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                    const __grid_constant__ CUtensorMap tm_v, const PrefillParams p) {
    constexpr int DH = D / 64;                  // 64-column (128-byte) sub-tiles along the head dimension
    constexpr int Q_BYTES = DH * ATOM_BYTES;    // Q tile: 128 rows
    constexpr int HALF_ATOM = BNH * 128;        // 64 rows x 128 bytes
    constexpr int KV_BYTES = DH * HALF_ATOM;    // one K or V block of a group: 64 keys x D
    constexpr int GROUP_BYTES = 4 * KV_BYTES + ATOM_BYTES;   // K[2], V[2], P (128 rows x 64 keys)
    constexpr int NSTAT = 2 + MAX_WINDOWS;      // col, obs, bands
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[16];  // 0 q; group g at 1 + 7 g: k[2], v[2], s[2], pv
    __shared__ uint32_t s_tmem_base;
    __shared__ float s_rowstat[4][BM][2];
    __shared__ float s_comb[2][NSTAT][BNH];     // per-block column sums of each group, flushed one block later

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int grp = tid >> 8, gt = tid & 255, gwarp = gt >> 5;
    const int row = gt & 127, ch = gt >> 7;
    const int h = blockIdx.x % p.Hkv;
    const int qt = p.n_qtiles - 1 - blockIdx.x / p.Hkv;
    const int i0 = qt * p.QB;
    const int lim = min(i0 + p.QB, p.S);                        // keys [0, lim) are visible to this query block
    const int T = max(0, (lim - BNH * grp + BN - 1) / BN);      // 64-key blocks of this group
    const int T_other = max(0, (lim - BNH * (1 - grp) + BN - 1) / BN);

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sQ = smem_base;
    const uint32_t sG = sQ + Q_BYTES + grp * GROUP_BYTES;
    const uint32_t sK = sG, sV = sG + 2 * KV_BYTES, sP = sG + 4 * KV_BYTES;
    uint8_t* const sP_ptr = smem_raw + (sP - smem_u32(smem_raw));
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t BAR_Q = bar0, bG = bar0 + 8u * (1 + 7 * grp);
    auto BAR_K = [&](int b) { return bG + 8u * b; };
    auto BAR_V = [&](int b) { return bG + 8u * (2 + b); };
    auto BAR_S = [&](int b) { return bG + 8u * (4 + b); };
    const uint32_t BAR_PV = bG + 8u * 6;

    if (tid == 0) {
        for (int i = 0; i < 15; ++i) mbar_init_u32(bar0 + 8u * i, 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(&s_tmem_base)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (p.G * p.QB < BM) {   // rows the Q box does not cover: zero them once (they only feed unused S rows)
        uint4* q4 = reinterpret_cast<uint4*>(smem_raw + (sQ - smem_u32(smem_raw)));
        for (int i = tid; i < Q_BYTES / 16; i += NTHREADS) q4[i] = make_uint4(0, 0, 0, 0);
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem_base;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tS = tmem + grp * 128, tY = tmem + 256 + grp * 128;

    constexpr uint32_t IDESC_S = make_idesc(BM, BNH, false);
    constexpr uint32_t IDESC_PV = make_idesc(BM, D, true);

    auto issue_k = [&](int blk, int buf) {
        mbar_expect_tx_u32(BAR_K(buf), KV_BYTES);
#pragma unroll
        for (int dh = 0; dh < DH; ++dh)
            tma_load_3d(sK + buf * KV_BYTES + dh * HALF_ATOM, &tm_k, dh * 64, blk * BN + grp * BNH, h, BAR_K(buf));
    };
    auto issue_v = [&](int blk, int buf) {
        mbar_expect_tx_u32(BAR_V(buf), KV_BYTES);
#pragma unroll
        for (int dh = 0; dh < DH; ++dh)
            tma_load_3d(sV + buf * KV_BYTES + dh * HALF_ATOM, &tm_v, dh * 64, blk * BN + grp * BNH, h, BAR_V(buf));
    };
    // S[buf] = Q K_buf^T : [128 x 64], D/16 MMAs of K = 16
    auto mma_s = [&](int buf) {
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {
            const uint32_t koff = (k & 3) * 32;
            umma_bf16(tS + buf * BNH, smem_desc(sQ + (k >> 2) * ATOM_BYTES + koff, 16, 1024),
                      smem_desc(sK + buf * KV_BYTES + (k >> 2) * HALF_ATOM + koff, 16, 1024), IDESC_S, k > 0);
        }
        umma_commit(BAR_S(buf));
    };
    // Y (+)= P V_buf : 64 keys = 4 MMAs; V is [key][d] in shared memory = MN-major B operand
    auto mma_pv = [&](int buf, bool first) {
#pragma unroll
        for (int k = 0; k < BNH / 16; ++k)
            umma_bf16(tY, smem_desc(sP + k * 32, 16, 1024), smem_desc(sV + buf * KV_BYTES + k * 2048, HALF_ATOM, 1024),
                      IDESC_PV, !(first && k == 0));
        umma_commit(BAR_PV);
    };

    const int NT = 2 * T;   // block visits of this group: sweep 1 then sweep 2
    if (tid == 0) {
        mbar_expect_tx_u32(BAR_Q, (uint32_t)(DH * p.G * p.QB * 128));
#pragma unroll
        for (int dh = 0; dh < DH; ++dh) tma_load_3d(sQ + dh * ATOM_BYTES, &tm_q, dh * 64, i0, h * p.G, BAR_Q);
    }
    if (gt == 0 && T > 0) {
        issue_k(0, 0);
        issue_k(1 % T, 1);
        issue_v(0, 0);
        if (T > 1) issue_v(1, 1);
        mbar_wait_u32(BAR_Q, 0);
        mbar_wait_u32(BAR_K(0), 0);
        tc_fence_after();
        mma_s(0);
    }
    __syncwarp();

    // this thread's row
    const int qi = i0 + row % p.QB;
    float m_run = -INFINITY, s_run = 0.f;
    float m_fin = 0.f, inv_s = 0.f;
    const float scale = p.scale;

    // pooled-map reductions: warp = 16-byte chunk of key columns (tile_stats); the sums of block t are written to the
    // scratch during block t + 1 (after its first barrier) so no extra barrier is needed
    const bool want_obs = qt >= p.first_obs_tile;
    auto flush_stats = [&](int t) {   // group threads < 64: one key column of block t
        const int j = t * BN + grp * BNH + gt;
        const size_t slab = (size_t)p.S_pad;
        p.part_col[((size_t)h * p.n_qtiles + qt) * slab + j] = s_comb[grp][0][gt];
        if (want_obs) p.part_obs[((size_t)h * p.n_obs_tiles + (qt - p.first_obs_tile)) * slab + j] = s_comb[grp][1][gt];
        for (int w = 0; w < p.n_windows; ++w)
            p.part_band[(((size_t)w * p.Hkv + h) * p.n_qtiles + qt) * slab + j] = s_comb[grp][2 + w][gt];
    };

    // one block visit up to the logits: l[c] = logit of key (block t, 32 ch + c), -inf past the diagonal
    auto visit = [&](int n, int t, float (&l)[32]) {
        const int sb = n & 1;
        if (gt == 0 && n + 1 < NT) {   // next block's logits under this block's softmax
            mbar_wait_u32(BAR_K((n + 1) & 1), ((n + 1) >> 1) & 1);
            tc_fence_after();
            mma_s((n + 1) & 1);
        }
        __syncwarp();
        mbar_wait_u32(BAR_S(sb), (n >> 1) & 1);
        tc_fence_after();
        uint32_t v0[32];
        tmem_ld32(tS + lane_base + sb * BNH + ch * 32, v0);
        if (gt == 0 && n + 2 < NT) issue_k((n + 2) % T, sb);   // K[sb] is free: its MMA has completed
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            const float x = __uint_as_float(v0[c]);
            l[c] = NAIVE ? round_bf16(round_bf16(x) * scale) : x * scale;
        }
        if (t * BN + grp * BNH + BNH - 1 > i0) {   // the block reaches past the diagonal for some row
            const int kbase = t * BN + grp * BNH + ch * 32;
#pragma unroll
            for (int c = 0; c < 32; ++c)
                if (kbase + c > qi) l[c] = -INFINITY;
        }
    };

    // ---- sweep 1: softmax statistics
    for (int n = 0; n < T; ++n) {
        float l[32];
        visit(n, n, l);
        float tmax = l[0];
#pragma unroll
        for (int c = 1; c < 32; ++c) tmax = fmaxf(tmax, l[c]);
        const float m_new = fmaxf(m_run, tmax);
        if (m_new > -INFINITY) {
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int c = 0; c < 32; c += 2) {
                a0 += ex2((l[c] - m_new) * LOG2E);
                a1 += ex2((l[c + 1] - m_new) * LOG2E);
            }
            s_run = s_run * ex2((m_run - m_new) * LOG2E) + (a0 + a1);
            m_run = m_new;
        }
        tc_fence_before();
        group_sync(grp);   // every warp of the group has read S[n & 1] before the MMA two visits ahead overwrites it
    }
    // merge the four key quarters of every row (two per group)
    s_rowstat[grp * 2 + ch][row][0] = m_run;
    s_rowstat[grp * 2 + ch][row][1] = s_run;
    __syncthreads();
    {
        float mq[4], sq[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) { mq[c] = s_rowstat[c][row][0]; sq[c] = s_rowstat[c][row][1]; }
        m_fin = fmaxf(fmaxf(mq[0], mq[1]), fmaxf(mq[2], mq[3]));   // finite: key 0 is visible to every query
        float tot = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) tot += sq[c] * ex2((mq[c] - m_fin) * LOG2E);
        inv_s = 1.0f / tot;
    }

    // ---- sweep 2: probabilities, P V, pooled reductions
    for (int t = 0; t < T; ++t) {
        float l[32];
        visit(T + t, t, l);
        uint32_t pk[16];
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
            const float p0 = ex2((l[c] - m_fin) * LOG2E) * inv_s;
            const float p1 = ex2((l[c + 1] - m_fin) * LOG2E) * inv_s;
            pk[c >> 1] = cvt_bf16x2(p0, p1);
        }
        if (t > 0) mbar_wait_u32(BAR_PV, (t - 1) & 1);   // P V of the previous block done: sP and V[(t-1)&1] are free
        group_sync(grp);                                  // the group finished reading the previous P block
        if (gt == 0 && t >= 1 && t + 1 < T) issue_v(t + 1, (t + 1) & 1);
        if (p.stats && t > 0 && gt < BNH) flush_stats(t - 1);
        {
            uint8_t* prow = sP_ptr + row * 128;
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                const int c8 = ch * 4 + c4;
                *reinterpret_cast<uint4*>(prow + ((c8 ^ (row & 7)) << 4)) =
                    make_uint4(pk[4 * c4], pk[4 * c4 + 1], pk[4 * c4 + 2], pk[4 * c4 + 3]);
            }
        }
        fence_proxy_async();
        tc_fence_before();
        group_sync(grp);
        if (gt == 0) {
            tc_fence_after();
            mbar_wait_u32(BAR_V(t & 1), (t >> 1) & 1);
            tc_fence_after();
            mma_pv(t & 1, t == 0);
        }
        __syncwarp();
        if (p.stats) {
            const int jbase = t * BN + grp * BNH + gwarp * 8;
            switch (p.QB) {
                case 128: tile_stats<7>(sP_ptr, s_comb[grp], p, gwarp, lane, i0, jbase, want_obs); break;
                case 64:  tile_stats<6>(sP_ptr, s_comb[grp], p, gwarp, lane, i0, jbase, want_obs); break;
                case 32:  tile_stats<5>(sP_ptr, s_comb[grp], p, gwarp, lane, i0, jbase, want_obs); break;
                default:  tile_stats<4>(sP_ptr, s_comb[grp], p, gwarp, lane, i0, jbase, want_obs); break;
            }
        }
    }

    // ---- epilogue: Y_0 + Y_1 (fp32 in TMEM) -> bf16 -> global
    if (T > 0) mbar_wait_u32(BAR_PV, (T - 1) & 1);
    tc_fence_after();
    __syncthreads();   // both groups' accumulators and last column sums are complete
    tc_fence_after();
    if (p.stats && T > 0 && gt < BNH) flush_stats(T - 1);
    {
        constexpr int CPT = D / 4;   // columns per thread: 32 (D = 128) or 16 (D = 64)
        const int cq = tid >> 7, orow = tid & 127;
        const int og = orow / p.QB, oqi = i0 + orow % p.QB;
        const bool row_ok = (og < p.G) && (oqi < p.S);
        const bool two = (grp == 0 ? T_other : T) > 0;   // group 1 saw at least one block
        uint32_t y0[32], y1[32];
        const uint32_t ycol = (CPT == 32) ? cq * 32 : (cq >> 1) * 32;   // D = 64: load the enclosing 32 columns
        tmem_ld32(tmem + 256 + lane_base + ycol, y0);
        if (two) tmem_ld32(tmem + 256 + 128 + lane_base + ycol, y1);
        __nv_bfloat16* yrow = p.y + (long long)(h * p.G + og) * p.y_stride_h + (long long)oqi * p.y_stride_s + cq * CPT;
        if (row_ok) {
            float f[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) f[c] = __uint_as_float(y0[c]) + (two ? __uint_as_float(y1[c]) : 0.f);
#pragma unroll
            for (int c = 0; c < CPT; c += 8) {
                float e[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) e[u] = (CPT == 32) ? f[c + u] : ((cq & 1) ? f[16 + c + u] : f[c + u]);
                *reinterpret_cast<uint4*>(yrow + c) = make_uint4(pack_bf16(e[0], e[1]), pack_bf16(e[2], e[3]),
                                                                  pack_bf16(e[4], e[5]), pack_bf16(e[6], e[7]));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(TMEM_COLS) : "memory");
    }
}

// Fixed-order sum of the per-tile partials: one thread per (head, column).
__global__ void __launch_bounds__(256)
prefill_reduce_kernel(const PrefillParams p, float* __restrict__ col_sum, float* __restrict__ obs_sum,
                      float* __restrict__ band_sum) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int h = blockIdx.y;
    if (j >= p.S) return;
    const size_t slab = (size_t)p.S_pad;
    const int qt0 = j / p.QB;
    float acc = 0.f;
    for (int qt = qt0; qt < p.n_qtiles; ++qt) {
        acc += p.part_col[((size_t)h * p.n_qtiles + qt) * slab + j];
    }
    col_sum[(size_t)h * p.S + j] = acc;
    if (obs_sum) {
        float o = 0.f;
        for (int qt = max(qt0, p.first_obs_tile); qt < p.n_qtiles; ++qt) {
            o += p.part_obs[((size_t)h * p.n_obs_tiles + (qt - p.first_obs_tile)) * slab + j];
        }
        obs_sum[(size_t)h * p.S + j] = o;
    }
    for (int w = 0; w < p.n_windows; ++w) {
        const int qt1 = min(p.n_qtiles - 1, (j + p.windows[w] - 1) / p.QB);   // later query tiles are outside the band
        float b = 0.f;
        for (int qt = qt0; qt <= qt1; ++qt) {
            b += p.part_band[(((size_t)w * p.Hkv + h) * p.n_qtiles + qt) * slab + j];
        }
        band_sum[((size_t)w * p.Hkv + h) * p.S + j] = b;
    }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
    }
    return fn;
}

// bf16 tensor [heads][S][D] with element strides (head, row); box = (64 columns, box_rows, box_heads), 128B swizzle
int make_map(CUtensorMap* tm, const void* base, int D, int S, int heads, long long stride_s, long long stride_h,
             int box_rows, int box_heads) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return set_error(CC_ERR_CUDA, "cc_prefill_attention: cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[3] = {(cuuint64_t)D, (cuuint64_t)S, (cuuint64_t)heads};
    cuuint64_t strides[2] = {(cuuint64_t)stride_s * 2, (cuuint64_t)stride_h * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)box_rows, (cuuint32_t)box_heads};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char msg[160];
        snprintf(msg, sizeof(msg), "cc_prefill_attention: cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
        return set_error(CC_ERR_CUDA, msg);
    }
    return CC_OK;
}

int query_block(int G) {   // largest power of two <= 128 / G
    int qb = 1;
    while (qb * 2 * G <= BM) qb *= 2;
    return qb;
}

struct Plan {
    int QB, n_qtiles, S_pad, first_obs_tile, n_obs_tiles;
    size_t col_floats, obs_floats, band_floats;
};

Plan make_plan(int S, int Hkv, int G, int obs_len, int n_windows) {
    Plan pl;
    pl.QB = query_block(G);
    pl.n_qtiles = (S + pl.QB - 1) / pl.QB;
    pl.S_pad = (S + BN - 1) / BN * BN;
    const int obs = obs_len > S ? S : obs_len;
    pl.first_obs_tile = obs > 0 ? (S - obs) / pl.QB : pl.n_qtiles;
    pl.n_obs_tiles = pl.n_qtiles - pl.first_obs_tile;
    pl.col_floats = (size_t)Hkv * pl.n_qtiles * pl.S_pad;
    pl.obs_floats = (size_t)Hkv * pl.n_obs_tiles * pl.S_pad;
    pl.band_floats = (size_t)n_windows * pl.col_floats;
    return pl;
}

template <int D, bool NAIVE>
int launch(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const PrefillParams& p, int grid,
           cudaStream_t stream) {
    const int smem = 1024 + (D / 64) * ATOM_BYTES + 2 * (4 * (D / 64) * BNH * 128 + ATOM_BYTES);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(prefill_attn_kernel<D, NAIVE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return set_cuda_error(e, "cc_prefill_attention: shared memory attribute");
        configured = true;
    }
    prefill_attn_kernel<D, NAIVE><<<grid, NTHREADS, smem, stream>>>(tq, tk, tv, p);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? CC_OK : set_cuda_error(e, "cc_prefill_attention launch");
}

}  // namespace

unsigned long long prefill_attention_workspace_impl(int S, int Hkv, int G, int obs_len, int n_windows) {
    if (S <= 0 || Hkv <= 0 || G <= 0 || G > 8 || n_windows < 0 || n_windows > MAX_WINDOWS) return 0;
    const Plan pl = make_plan(S, Hkv, G, obs_len, n_windows);
    return (uint64_t)(pl.col_floats + pl.obs_floats + pl.band_floats) * sizeof(float) + 256;
}

int prefill_attention_impl(const CcPrefillAttnArgs* a, cudaStream_t stream) {
    if (!a || !a->q || !a->k || !a->v || !a->y) return set_error(CC_ERR_INVALID_ARG, "cc_prefill_attention: null argument");
    const int S = a->seq_len, Hkv = a->n_kv_heads, G = a->q_per_kv, D = a->head_dim;
    if (S <= 0 || Hkv <= 0) return set_error(CC_ERR_INVALID_ARG, "cc_prefill_attention: empty problem");
    if ((D != 64 && D != 128) || G < 1 || G > 8)
        return set_error(CC_ERR_UNSUPPORTED_SHAPE, "cc_prefill_attention: head_dim must be 64 or 128 and q_per_kv <= 8");
    if (a->n_windows < 0 || a->n_windows > MAX_WINDOWS)
        return set_error(CC_ERR_INVALID_ARG, "cc_prefill_attention: at most 4 band windows");
    if (a->attn_mode != CC_ATTN_NAIVE_BF16 && a->attn_mode != CC_ATTN_MATH_FP32)
        return set_error(CC_ERR_INVALID_ARG, "cc_prefill_attention: bad attn_mode");
    const long long strides[] = {a->q_stride_s, a->q_stride_h, a->k_stride_s, a->k_stride_h, a->v_stride_s,
                                 a->v_stride_h, a->y_stride_s, a->y_stride_h};
    for (long long s : strides)
        if (s <= 0 || (s % 8) != 0) return set_error(CC_ERR_INVALID_ARG, "cc_prefill_attention: strides must be positive multiples of 8 elements");
    const uintptr_t ptrs[] = {(uintptr_t)a->q, (uintptr_t)a->k, (uintptr_t)a->v, (uintptr_t)a->y};
    for (uintptr_t q : ptrs)
        if (q % 16) return set_error(CC_ERR_INVALID_ARG, "cc_prefill_attention: tensors must be 16-byte aligned");
    const bool stats = a->col_sum != nullptr;
    if (!stats && (a->obs_sum || a->band_sum || a->n_windows))
        return set_error(CC_ERR_INVALID_ARG, "cc_prefill_attention: obs_sum / band_sum need col_sum");
    if (a->n_windows && !a->band_sum) return set_error(CC_ERR_INVALID_ARG, "cc_prefill_attention: band_sum missing");

    const int obs_len = (stats && a->obs_sum) ? (a->obs_len > S ? S : a->obs_len) : 0;
    const Plan pl = make_plan(S, Hkv, G, obs_len, stats ? a->n_windows : 0);
    PrefillParams p;
    p.y = (__nv_bfloat16*)a->y; p.y_stride_s = a->y_stride_s; p.y_stride_h = a->y_stride_h;
    p.S = S; p.S_pad = pl.S_pad; p.Hkv = Hkv; p.G = G; p.QB = pl.QB; p.n_qtiles = pl.n_qtiles;
    p.obs_start = obs_len > 0 ? S - obs_len : S;
    p.first_obs_tile = pl.first_obs_tile; p.n_obs_tiles = pl.n_obs_tiles;
    p.n_windows = stats ? a->n_windows : 0;
    for (int w = 0; w < MAX_WINDOWS; ++w) p.windows[w] = w < p.n_windows ? a->windows[w] : 0;
    p.scale = (float)(1.0 / sqrt((double)D));
    p.stats = stats ? 1 : 0;
    p.part_col = p.part_obs = p.part_band = nullptr;
    if (stats) {
        const uint64_t need = prefill_attention_workspace_impl(S, Hkv, G, obs_len, p.n_windows);
        if (!a->workspace || a->workspace_bytes < need)
            return set_error(CC_ERR_INVALID_ARG, "cc_prefill_attention: workspace too small (see cc_prefill_attention_workspace)");
        float* w = (float*)(((uintptr_t)a->workspace + 255) & ~(uintptr_t)255);
        p.part_col = w;
        p.part_obs = w + pl.col_floats;
        p.part_band = p.part_obs + pl.obs_floats;
    }
    CUtensorMap tq, tk, tv;
    int rc;
    if ((rc = make_map(&tq, a->q, D, S, Hkv * G, a->q_stride_s, a->q_stride_h, pl.QB, G)) != CC_OK) return rc;
    if ((rc = make_map(&tk, a->k, D, S, Hkv, a->k_stride_s, a->k_stride_h, BNH, 1)) != CC_OK) return rc;
    if ((rc = make_map(&tv, a->v, D, S, Hkv, a->v_stride_s, a->v_stride_h, BNH, 1)) != CC_OK) return rc;
    const int grid = pl.n_qtiles * Hkv;
    const bool naive = a->attn_mode == CC_ATTN_NAIVE_BF16;
    if (D == 128) rc = naive ? launch<128, true>(tq, tk, tv, p, grid, stream) : launch<128, false>(tq, tk, tv, p, grid, stream);
    else          rc = naive ? launch<64, true>(tq, tk, tv, p, grid, stream) : launch<64, false>(tq, tk, tv, p, grid, stream);
    if (rc != CC_OK) return rc;
    if (stats) {
        dim3 rgrid((S + 255) / 256, Hkv);
        prefill_reduce_kernel<<<rgrid, 256, 0, stream>>>(p, a->col_sum, obs_len > 0 ? a->obs_sum : nullptr, a->band_sum);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return set_cuda_error(e, "cc_prefill_attention reduce launch");
    }
    return CC_OK;
}

}  // namespace cc
