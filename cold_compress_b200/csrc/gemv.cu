// gemv.cu -- the rest of the decode layer around the cache kernel:
// one-token matrix-vector products over bf16 weights with the elementwise work fused in, so a
// decoded token is 5 launches per layer instead of ~60 ATen kernels:
//
//   rmsnorm_qkv_rope : attention_norm -> wqkv (+bias) -> RoPE           (model.py:376-387, 446-457, 507-519)
//   gemv_residual    : wo / w2 projection (+ residual add)              (model.py:431, 317, 326, 443)
//   rmsnorm_ffn_gate : ffn_norm -> silu(w1 x) * w3 x                    (model.py:442-443)
//   rmsnorm_gemv     : final norm -> output head                        (model.py:290-291)
//
// All are HBM-bound streams of the weight matrix (15 GB per token for Llama-3-8B): each warp owns
// output rows, every lane keeps 4-8 independent 16-byte loads in flight (ld.global.nc, no L1
// allocation), fp32 accumulation, one shuffle reduction per row.  The bf16 rounding points are the
// reference's (every nn.Linear output, the normalised activations, RoPE outputs and the residual sums
// are bf16 tensors there).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "common.cuh"
#include "internal.h"

namespace cc {

constexpr int kGemvThreads = 256;
constexpr int kGemvWarps = kGemvThreads / 32;

__device__ __forceinline__ uint4 ldg_stream(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// ---------------------------------------------------------------- fused tensor-parallel exchange
constexpr int kMaxTp = 8;
struct TpDev {
    const uint64_t* peer_buf;   // [world] base addresses of the exchange buffers (u64 [n_ops][world][dim / 2])
    const uint64_t* peer_flag;  // unused (flag-in-data protocol)
    int32_t* cta_done;
    const int32_t* epoch;
    __nv_bfloat16* recv_out;
    int world, rank, dim, n_ops, send_op, recv_op;
};

__device__ __forceinline__ int32_t ld_acquire_sys(const int32_t* p) {
    int32_t v; asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_release_sys(int32_t* p, int32_t v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

// Exchange protocol: flag-in-data ("LL") words.  The exchange buffer of a rank is u64 [n_ops][world][dim / 2]; word
// [op][r][i] = {two bf16 partial sums (rows 2i, 2i+1) of rank r, epoch}.  8-byte stores are delivered atomically
// over NVLink, so a word whose upper half equals the current epoch carries valid data: no fences, no flags, no
// "last CTA" -- the producer lane that owns a row pair stores it straight into every rank's buffer the moment it is
// computed, and the exchange costs one NVLink write latency.
__device__ __forceinline__ void tp_push_pair(const TpDev& tp, int r0, float a, float b, int32_t ep) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    const uint32_t data = *reinterpret_cast<const uint32_t*>(&v);
    const size_t word = (size_t)(tp.send_op * tp.world + tp.rank) * (tp.dim / 2) + (r0 >> 1);
    for (int p = 0; p < tp.world; ++p)
        asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};"
                     :: "l"(reinterpret_cast<uint64_t*>(tp.peer_buf[p]) + word), "r"(data), "r"((uint32_t)ep) : "memory");
}
// Consumer side: 8 elements of x per thread and pass = 4 words from each of the W ranks.  All 4 W loads are issued
// before any flag is looked at (one memory round trip when the data is there; re-polled as a whole otherwise).
// Consumer side: 8 elements of x per thread and pass = 4 words from each rank; the 8 loads of two ranks are issued
// before any flag is looked at (one memory round trip when the data is there; re-polled as a whole otherwise).
__device__ __forceinline__ void tp_collect(const TpDev& tp, const uint64_t* recv, uint32_t ep, const __nv_bfloat16* x,
                                           __nv_bfloat16* xs, int K) {
    const int tid = threadIdx.x;
    for (int i = tid * 8; i < K; i += kGemvThreads * 8) {
        float acc[8], f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        // two ranks (8 loads) per round: the prefetched weight registers of the caller stay live across this loop
        for (int r0 = 0; r0 < tp.world; r0 += 2) {   // rank order: the same bits on every rank (world is even)
            uint32_t d[2][4];
            uint32_t bad;
            do {
                bad = 0;
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const uint64_t* w = recv + (size_t)(r0 + r) * (tp.dim / 2) + (i >> 1);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint32_t fl;
                        asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(d[r][j]), "=r"(fl) : "l"(w + j) : "memory");
                        bad |= fl ^ ep;
                    }
                }
            } while (bad);
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                unpack8(make_uint4(d[r][0], d[r][1], d[r][2], d[r][3]), f);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] += f[j];
            }
        }
        unpack8(*reinterpret_cast<const uint4*>(x + i), f);
        __nv_bfloat16 o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = __float2bfloat16_rn(__fadd_rn(f[j], bf16_round(acc[j])));
        *reinterpret_cast<uint4*>(xs + i) = *reinterpret_cast<uint4*>(o);
        if (blockIdx.x == 0 && tp.recv_out) *reinterpret_cast<uint4*>(tp.recv_out + i) = *reinterpret_cast<uint4*>(o);
    }
}

// x (bf16 [K]) -> shared memory as bf16, optionally RMS-normalised the way the reference's RMSNorm
// does it: bf16(bf16(float(x) * rsqrt(mean(x^2) + eps)) * weight)   (model.py:452-457)
__device__ void stage_input(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ norm_w,
                            float eps, int K, __nv_bfloat16* xs, float* scratch, const TpDev& tp) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tp.recv_op >= 0) {
        // consumer of a fused all-reduce: collect every rank's partial vector of this op (polling the flag-in-data
        // words), then x = bf16(residual + bf16(sum of the partials))   (tp.py:156-160 + model.py:317 / 326)
        const uint32_t ep = (uint32_t)*reinterpret_cast<const volatile int32_t*>(tp.epoch);
        const uint64_t* recv = reinterpret_cast<const uint64_t*>(tp.peer_buf[tp.rank]) +
                               (size_t)tp.recv_op * tp.world * (tp.dim / 2);
        tp_collect(tp, recv, ep, x, xs, K);
    } else {
        for (int i = tid * 8; i < K; i += kGemvThreads * 8)
            *reinterpret_cast<uint4*>(xs + i) = *reinterpret_cast<const uint4*>(x + i);
    }
    __syncthreads();
    if (norm_w == nullptr) return;
    // RMSNorm in place on the staged vector
    float ss = 0.f;
    for (int i = tid * 8; i < K; i += kGemvThreads * 8) {
        float f[8];
        unpack8(*reinterpret_cast<const uint4*>(xs + i), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) ss = fmaf(f[j], f[j], ss);
    }
    ss = warp_sum(ss);
    if (lane == 0) scratch[warp] = ss;
    __syncthreads();
    float tot = 0.f;
    for (int w = 0; w < kGemvWarps; ++w) tot += scratch[w];
    const float rstd = rsqrtf(tot / (float)K + eps);
    for (int i = tid * 8; i < K; i += kGemvThreads * 8) {
        float f[8], wv[8];
        unpack8(*reinterpret_cast<const uint4*>(xs + i), f);
        unpack8(*reinterpret_cast<const uint4*>(norm_w + i), wv);
        __nv_bfloat16 o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = __float2bfloat16_rn(__fmul_rn(bf16_round(__fmul_rn(f[j], rstd)), wv[j]));
        *reinterpret_cast<uint4*>(xs + i) = *reinterpret_cast<uint4*>(o);
    }
    __syncthreads();
}

// A warp's share of R weight rows over the K-range [k0, k1) (a multiple of 256 long), in two steps so
// that the weight stream starts before anything else the kernel has to do:
//   prefetch(): issues the first NV K-steps of every row (NV*R independent 16-byte loads per lane);
//   finish():   consumes them against the staged input and streams the rest.
template <int R, int NV>
struct WarpDots {
    uint4 w[NV][R];
    int k, k1;
    __device__ __forceinline__ void prefetch(const __nv_bfloat16* const (&rows)[R], int k0, int k1_) {
        const int lane = threadIdx.x & 31;
        k = k0 + lane * 8;
        k1 = k1_;
#pragma unroll
        for (int u = 0; u < NV; ++u)
#pragma unroll
            for (int r = 0; r < R; ++r)
                w[u][r] = (k + u * 256 < k1) ? ldg_stream(rows[r] + k + u * 256) : make_uint4(0, 0, 0, 0);
    }
    __device__ __forceinline__ void finish(const __nv_bfloat16* const (&rows)[R], const __nv_bfloat16* xs, float (&out)[R]) {
        float acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = 0.f;
        // rolling refill: the moment slot u has been consumed its registers are re-armed with the loads of the next
        // round, so (NV - 1) / NV of the warp's loads stay in flight instead of draining to zero once per round
        while (true) {
#pragma unroll
            for (int u = 0; u < NV; ++u) {
                if (k + u * 256 < k1) {
                    float x[8];
                    unpack8(*reinterpret_cast<const uint4*>(xs + k + u * 256), x);
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        float f[8];
                        unpack8(w[u][r], f);
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[r] = fmaf(f[j], x[j], acc[r]);
                    }
                }
                const int kn = k + (NV + u) * 256;
                if (kn < k1) {
#pragma unroll
                    for (int r = 0; r < R; ++r) w[u][r] = ldg_stream(rows[r] + kn);
                }
            }
            k += NV * 256;
            if (k >= k1) break;
        }
#pragma unroll
        for (int r = 0; r < R; ++r) out[r] = warp_sum(acc[r]);
    }
};

// KS warps of a CTA share a row group and split K between them; their partial sums are combined in
// shared memory in a fixed order (deterministic).  Returns true in the lane that owns the epilogue,
// with the complete sums in out[].
template <int R, int KS>
__device__ __forceinline__ bool combine_split(float (*part)[R], float (&out)[R]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if constexpr (KS == 1) return lane == 0;
    const int ks = warp % KS;
    if (lane == 0) {
#pragma unroll
        for (int r = 0; r < R; ++r) part[warp][r] = out[r];
    }
    __syncthreads();
    if (ks == 0 && lane == 0) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float s = part[warp][r];
#pragma unroll
            for (int j = 1; j < KS; ++j) s += part[warp + j][r];
            out[r] = s;
        }
    }
    return ks == 0 && lane == 0;
}

// Skeleton shared by the three kernels: every warp owns row group `base + warp / KS` and the K-slice
// `warp % KS`.  First pass: weights are requested BEFORE the dependency wait and the input staging.
#define CC_GEMV_PROLOGUE(R_, NV_, ROWS_OF)                                                          \
    constexpr int GPB = kGemvWarps / KS;                                                            \
    const int warp = threadIdx.x >> 5;                                                              \
    const int klen = p.K / KS, k0 = (warp % KS) * klen;                                             \
    WarpDots<R_, NV_> wd;                                                                           \
    int base = blockIdx.x * GPB;                                                                    \
    {                                                                                               \
        const int pr = min(base + warp / KS, n_pairs - 1);                                          \
        const int r0 = 2 * pr;                                                                      \
        const __nv_bfloat16* rows[R_] = ROWS_OF;                                                    \
        wd.prefetch(rows, k0, k0 + klen);                                                           \
    }                                                                                               \
    pdl_launch_dependents();                                                                        \
    pdl_wait();                                                                                     \
    stage_input(p.x, p.norm_w, p.eps, p.K, xs, scratch, p.tp);

// ---------------------------------------------------------------- attention_norm -> wqkv -> RoPE
struct QkvParams {
    const __nv_bfloat16 *x, *norm_w, *w, *bias, *freqs;  // freqs: [max_pos, D/2, 2] (cos, sin)
    const int32_t* input_pos;
    __nv_bfloat16 *q, *k, *v;
    float eps;
    int K, n_q_rows, n_kv_rows, head_dim;
    TpDev tp;
};

template <int KS>
__global__ void __launch_bounds__(kGemvThreads, 3) rmsnorm_qkv_rope_kernel(const __grid_constant__ QkvParams p) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(smem_raw);
    __shared__ float scratch[kGemvWarps];
    __shared__ float part[kGemvWarps][2];
    const int n_pairs = (p.n_q_rows + 2 * p.n_kv_rows) / 2;
#define CC_ROWS {p.w + (size_t)r0 * p.K, p.w + (size_t)(r0 + 1) * p.K}
    CC_GEMV_PROLOGUE(2, 4, CC_ROWS)
    const int pos = p.input_pos[0];
    const int half = p.head_dim / 2;
    // a row group = the two rows of a rotary pair; loop bound is CTA-uniform (barriers inside)
    for (; base < n_pairs; base += gridDim.x * GPB) {
        const int pr = min(base + warp / KS, n_pairs - 1);
        const bool active = base + warp / KS < n_pairs;
        const int r0 = 2 * pr;
        const __nv_bfloat16* rows[2] = CC_ROWS;
        if (base != (int)blockIdx.x * GPB) wd.prefetch(rows, k0, k0 + klen);
        float d[2];
        wd.finish(rows, xs, d);
        const bool owner = combine_split<2, KS>(part, d);
        if (owner && active) {
            float a = d[0], b = d[1];
            if (p.bias) { a += __bfloat162float(p.bias[r0]); b += __bfloat162float(p.bias[r0 + 1]); }
            a = bf16_round(a);  // nn.Linear output is a bf16 tensor
            b = bf16_round(b);
            if (r0 < p.n_q_rows + p.n_kv_rows) {  // q and k rows rotate (model.py:507-519), v rows do not
                const int i = (r0 % p.head_dim) / 2;
                const float c = __bfloat162float(p.freqs[((size_t)pos * half + i) * 2]);
                const float s = __bfloat162float(p.freqs[((size_t)pos * half + i) * 2 + 1]);
                const float ra = __fsub_rn(__fmul_rn(a, c), __fmul_rn(b, s));
                const float rb = __fadd_rn(__fmul_rn(b, c), __fmul_rn(a, s));
                a = ra; b = rb;
            }
            __nv_bfloat16* dst = r0 < p.n_q_rows ? p.q + r0
                               : r0 < p.n_q_rows + p.n_kv_rows ? p.k + (r0 - p.n_q_rows)
                                                               : p.v + (r0 - p.n_q_rows - p.n_kv_rows);
            *reinterpret_cast<__nv_bfloat162*>(dst) = __floats2bfloat162_rn(a, b);
        }
        if (KS > 1) __syncthreads();  // `part` is reused by the next pass
    }
#undef CC_ROWS
}

// ---------------------------------------------------------------- y = [residual +] W x
struct GemvParams {
    const __nv_bfloat16 *x, *norm_w, *w, *residual;
    __nv_bfloat16* out;
    float eps;
    int K, N;
    TpDev tp;
};

template <int KS>
__global__ void __launch_bounds__(kGemvThreads, 3) gemv_residual_kernel(const __grid_constant__ GemvParams p) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(smem_raw);
    __shared__ float scratch[kGemvWarps];
    __shared__ float part[kGemvWarps][2];
    const int n_pairs = p.N / 2;
#define CC_ROWS {p.w + (size_t)r0 * p.K, p.w + (size_t)(r0 + 1) * p.K}
    CC_GEMV_PROLOGUE(2, 4, CC_ROWS)
    const int32_t send_ep = p.tp.send_op >= 0 ? *reinterpret_cast<const volatile int32_t*>(p.tp.epoch) : 0;
    for (; base < n_pairs; base += gridDim.x * GPB) {
        const int pr = min(base + warp / KS, n_pairs - 1);
        const bool active = base + warp / KS < n_pairs;
        const int r0 = 2 * pr;
        const __nv_bfloat16* rows[2] = CC_ROWS;
        if (base != (int)blockIdx.x * GPB) wd.prefetch(rows, k0, k0 + klen);
        float d[2];
        wd.finish(rows, xs, d);
        const bool owner = combine_split<2, KS>(part, d);
        if (owner && active) {
            float a = bf16_round(d[0]), b = bf16_round(d[1]);
            if (p.residual) {  // h = x + attention(...) / out = h + feed_forward(...): bf16 + bf16 -> bf16
                a = __fadd_rn(__bfloat162float(p.residual[r0]), a);
                b = __fadd_rn(__bfloat162float(p.residual[r0 + 1]), b);
            }
            if (p.tp.send_op >= 0) tp_push_pair(p.tp, r0, a, b, send_ep);
            else *reinterpret_cast<__nv_bfloat162*>(p.out + r0) = __floats2bfloat162_rn(a, b);
        }
        if (KS > 1) __syncthreads();
    }
#undef CC_ROWS
}

// ---------------------------------------------------------------- ffn_norm -> silu(w1 x) * (w3 x)
struct GateParams {
    const __nv_bfloat16 *x, *norm_w, *w1, *w3;
    __nv_bfloat16* out;
    float eps;
    int K, N;
    TpDev tp;
};

template <int KS>
__global__ void __launch_bounds__(kGemvThreads, 3) rmsnorm_ffn_gate_kernel(const __grid_constant__ GateParams p) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(smem_raw);
    __shared__ float scratch[kGemvWarps];
    __shared__ float part[kGemvWarps][4];
    const int n_pairs = p.N / 2;
#define CC_ROWS {p.w1 + (size_t)r0 * p.K, p.w1 + (size_t)(r0 + 1) * p.K, p.w3 + (size_t)r0 * p.K, p.w3 + (size_t)(r0 + 1) * p.K}
    CC_GEMV_PROLOGUE(4, 2, CC_ROWS)
    for (; base < n_pairs; base += gridDim.x * GPB) {
        const int pr = min(base + warp / KS, n_pairs - 1);
        const bool active = base + warp / KS < n_pairs;
        const int r0 = 2 * pr;
        const __nv_bfloat16* rows[4] = CC_ROWS;
        if (base != (int)blockIdx.x * GPB) wd.prefetch(rows, k0, k0 + klen);
        float d[4];
        wd.finish(rows, xs, d);
        const bool owner = combine_split<4, KS>(part, d);
        if (owner && active) {
            float o[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float g = bf16_round(d[j]);                       // w1(x), bf16 tensor
                const float sg = bf16_round(g / (1.0f + expf(-g)));     // F.silu in fp32, bf16 out
                o[j] = __fmul_rn(sg, bf16_round(d[2 + j]));             // * w3(x)
            }
            *reinterpret_cast<__nv_bfloat162*>(p.out + r0) = __floats2bfloat162_rn(o[0], o[1]);
        }
        if (KS > 1) __syncthreads();
    }
#undef CC_ROWS
}

// --------------------------------------------------------------------------- host launchers

static int fill_tp(const CcTpArgs* tp, int K, TpDev* out, const char* what) {
    TpDev d = {};
    d.send_op = d.recv_op = -1;
    if (tp && tp->link && (tp->send_op >= 0 || tp->recv_op >= 0)) {
        const CcTpLink* l = tp->link;
        if (l->world < 2 || l->world > kMaxTp || (l->world & 1) || l->rank < 0 || l->rank >= l->world || !l->peer_buf ||
            !l->epoch || (l->dim & 1) || tp->send_op >= l->n_ops || tp->recv_op >= l->n_ops)
            return set_error(CC_ERR_INVALID_ARG, what);
        if (tp->recv_op >= 0 && l->dim != K) return set_error(CC_ERR_INVALID_ARG, "tensor-parallel receive: dim must equal K");
        d.peer_buf = l->peer_buf; d.peer_flag = l->peer_flag; d.cta_done = l->cta_done; d.epoch = l->epoch;
        d.recv_out = (__nv_bfloat16*)tp->recv_out;
        d.world = l->world; d.rank = l->rank; d.dim = l->dim; d.n_ops = l->n_ops; d.send_op = tp->send_op; d.recv_op = tp->recv_op;
    }
    *out = d;
    return CC_OK;
}

static int g_sms = 0;
static int sm_count() {
    if (!g_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_sms <= 0) g_sms = 148;
    }
    return g_sms;
}

// K-split: the smallest KS in {1,2,4,8} (with K/KS a multiple of 256) that yields at least 4 CTAs per SM,
// otherwise the largest valid one.  Short-and-wide products (w2, wo) would otherwise leave most SMs
// with too few loads in flight.
static int pick_ksplit(int n_pairs, int K) {
    const int want = sm_count() * 4;
    int best = 1;
    for (int ks = 1; ks <= kGemvWarps; ks *= 2) {
        if (K % (ks * 256)) break;
        best = ks;
        if ((long long)n_pairs * ks / kGemvWarps >= want) break;
    }
    return best;
}

template <typename P, typename K1, typename K2, typename K4, typename K8>
static int launch_split(K1 k1, K2 k2, K4 k4, K8 k8, const P& p, int n_pairs, int Kdim, cudaStream_t stream, const char* what) {
    const int ks = pick_ksplit(n_pairs, Kdim);
    const size_t smem = (size_t)Kdim * 2;
    const int gpb = kGemvWarps / ks;
    int grid = (n_pairs + gpb - 1) / gpb;
    const int cap = sm_count() * 8;
    if (grid > cap) grid = cap;
    // (sharing the input staging between the CTAs of a cluster over DSMEM was measured 5-8 us per launch SLOWER:
    // gang scheduling + two cluster barriers, profiles/r1_notes.md)
    auto go = [&](auto kernel) -> int {
        static thread_local const void* raised[16] = {};
        if (smem > 48 * 1024) {  // raise the per-function cap once, to the device limit
            bool done = false;
            for (int i = 0; i < 16; ++i) done |= raised[i] == (const void*)kernel;
            if (!done) {
                int dev = 0, optin = 0;
                cudaGetDevice(&dev);
                cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
                cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
                if (e != cudaSuccess) return set_cuda_error(e, what);
                for (int i = 0; i < 16; ++i)
                    if (!raised[i]) { raised[i] = (const void*)kernel; break; }
            }
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(kGemvThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr.val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = &attr;
        cfg.numAttrs = pdl_enabled() ? 1 : 0;
        cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, p);
        return e == cudaSuccess ? CC_OK : set_cuda_error(e, what);
    };
    switch (ks) {
    case 1: return go(k1);
    case 2: return go(k2);
    case 4: return go(k4);
    default: return go(k8);
    }
}

int rmsnorm_qkv_rope_impl(const void* x, const void* norm_w, float eps, const void* wqkv, const void* bias,
                          const void* freqs, const int32_t* input_pos, void* q, void* k, void* v, int dim,
                          int n_q_heads, int n_kv_heads, int head_dim, const CcTpArgs* tp, cudaStream_t stream) {
    if (!x || !wqkv || !freqs || !input_pos || !q || !k || !v) return set_error(CC_ERR_INVALID_ARG, "cc_rmsnorm_qkv_rope: null argument");
    if (dim % 256 || head_dim % 2) return set_error(CC_ERR_UNSUPPORTED_SHAPE, "cc_rmsnorm_qkv_rope: dim must be a multiple of 256");
    QkvParams p = {(const __nv_bfloat16*)x, (const __nv_bfloat16*)norm_w, (const __nv_bfloat16*)wqkv,
                   (const __nv_bfloat16*)bias, (const __nv_bfloat16*)freqs, input_pos,
                   (__nv_bfloat16*)q, (__nv_bfloat16*)k, (__nv_bfloat16*)v, eps, dim,
                   n_q_heads * head_dim, n_kv_heads * head_dim, head_dim, {}};
    if (int rc = fill_tp(tp, dim, &p.tp, "cc_rmsnorm_qkv_rope: bad tensor-parallel link")) return rc;
    if (p.tp.send_op >= 0) return set_error(CC_ERR_INVALID_ARG, "cc_rmsnorm_qkv_rope cannot be an all-reduce producer");
    return launch_split(rmsnorm_qkv_rope_kernel<1>, rmsnorm_qkv_rope_kernel<2>, rmsnorm_qkv_rope_kernel<4>, rmsnorm_qkv_rope_kernel<8>,
                        p, (p.n_q_rows + 2 * p.n_kv_rows) / 2, dim, stream, "cc_rmsnorm_qkv_rope");
}

int gemv_residual_impl(const void* x, const void* norm_w, float eps, const void* w, const void* residual, void* out,
                       int K, int N, const CcTpArgs* tp, cudaStream_t stream) {
    if (!x || !w || !out) return set_error(CC_ERR_INVALID_ARG, "cc_gemv_residual: null argument");
    if (K % 256 || N % 2) return set_error(CC_ERR_UNSUPPORTED_SHAPE, "cc_gemv_residual: K must be a multiple of 256, N even");
    GemvParams p = {(const __nv_bfloat16*)x, (const __nv_bfloat16*)norm_w, (const __nv_bfloat16*)w,
                    (const __nv_bfloat16*)residual, (__nv_bfloat16*)out, eps, K, N, {}};
    if (int rc = fill_tp(tp, K, &p.tp, "cc_gemv_residual: bad tensor-parallel link")) return rc;
    if (p.tp.send_op >= 0 && (p.tp.dim != N || residual)) return set_error(CC_ERR_INVALID_ARG, "cc_gemv_residual: a producer writes dim rows and takes no residual");
    return launch_split(gemv_residual_kernel<1>, gemv_residual_kernel<2>, gemv_residual_kernel<4>, gemv_residual_kernel<8>,
                        p, N / 2, K, stream, "cc_gemv_residual");
}

int rmsnorm_ffn_gate_impl(const void* x, const void* norm_w, float eps, const void* w1, const void* w3, void* out,
                          int K, int N, const CcTpArgs* tp, cudaStream_t stream) {
    if (!x || !w1 || !w3 || !out) return set_error(CC_ERR_INVALID_ARG, "cc_rmsnorm_ffn_gate: null argument");
    if (K % 256 || N % 2) return set_error(CC_ERR_UNSUPPORTED_SHAPE, "cc_rmsnorm_ffn_gate: K must be a multiple of 256, N even");
    GateParams p = {(const __nv_bfloat16*)x, (const __nv_bfloat16*)norm_w, (const __nv_bfloat16*)w1,
                    (const __nv_bfloat16*)w3, (__nv_bfloat16*)out, eps, K, N, {}};
    if (int rc = fill_tp(tp, K, &p.tp, "cc_rmsnorm_ffn_gate: bad tensor-parallel link")) return rc;
    if (p.tp.send_op >= 0) return set_error(CC_ERR_INVALID_ARG, "cc_rmsnorm_ffn_gate cannot be an all-reduce producer");
    return launch_split(rmsnorm_ffn_gate_kernel<1>, rmsnorm_ffn_gate_kernel<2>, rmsnorm_ffn_gate_kernel<4>, rmsnorm_ffn_gate_kernel<8>,
                        p, N / 2, K, stream, "cc_rmsnorm_ffn_gate");
}

}  // namespace cc
