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

// x (bf16 [K]) -> shared memory as bf16, optionally RMS-normalised the way the reference's RMSNorm
// does it: bf16(bf16(float(x) * rsqrt(mean(x^2) + eps)) * weight)   (model.py:452-457)
__device__ void stage_input(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ norm_w,
                            float eps, int K, __nv_bfloat16* xs, float* scratch) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (norm_w == nullptr) {
        for (int i = tid * 8; i < K; i += kGemvThreads * 8)
            *reinterpret_cast<uint4*>(xs + i) = *reinterpret_cast<const uint4*>(x + i);
        __syncthreads();
        return;
    }
    float ss = 0.f;
    for (int i = tid * 8; i < K; i += kGemvThreads * 8) {
        float f[8];
        unpack8(*reinterpret_cast<const uint4*>(x + i), f);
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
        unpack8(*reinterpret_cast<const uint4*>(x + i), f);
        unpack8(*reinterpret_cast<const uint4*>(norm_w + i), wv);
        __nv_bfloat16 o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = __float2bfloat16_rn(__fmul_rn(bf16_round(__fmul_rn(f[j], rstd)), wv[j]));
        *reinterpret_cast<uint4*>(xs + i) = *reinterpret_cast<uint4*>(o);
    }
    __syncthreads();
}

// dot products of R weight rows (row pointers) with the staged input; all lanes get the sums
template <int R>
__device__ __forceinline__ void warp_dots(const __nv_bfloat16* const (&rows)[R], const __nv_bfloat16* xs, int K,
                                          float (&out)[R]) {
    const int lane = threadIdx.x & 31;
    float acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.f;
    int k = lane * 8;
    // two K-steps per iteration: 2*R independent 16-byte loads in flight per lane
    for (; k + 256 < K; k += 512) {
        uint4 w0[R], w1[R];
#pragma unroll
        for (int r = 0; r < R; ++r) { w0[r] = ldg_stream(rows[r] + k); w1[r] = ldg_stream(rows[r] + k + 256); }
        float x0[8], x1[8];
        unpack8(*reinterpret_cast<const uint4*>(xs + k), x0);
        unpack8(*reinterpret_cast<const uint4*>(xs + k + 256), x1);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float f[8];
            unpack8(w0[r], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[r] = fmaf(f[j], x0[j], acc[r]);
            unpack8(w1[r], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[r] = fmaf(f[j], x1[j], acc[r]);
        }
    }
    for (; k < K; k += 256) {
        float x0[8];
        unpack8(*reinterpret_cast<const uint4*>(xs + k), x0);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float f[8];
            unpack8(ldg_stream(rows[r] + k), f);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[r] = fmaf(f[j], x0[j], acc[r]);
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) out[r] = warp_sum(acc[r]);
}

// ---------------------------------------------------------------- attention_norm -> wqkv -> RoPE
struct QkvParams {
    const __nv_bfloat16 *x, *norm_w, *w, *bias, *freqs;  // freqs: [max_pos, D/2, 2] (cos, sin)
    const int32_t* input_pos;
    __nv_bfloat16 *q, *k, *v;
    float eps;
    int K, n_q_rows, n_kv_rows, head_dim;
};

__global__ void __launch_bounds__(kGemvThreads) rmsnorm_qkv_rope_kernel(const __grid_constant__ QkvParams p) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(smem_raw);
    __shared__ float scratch[kGemvWarps];
    stage_input(p.x, p.norm_w, p.eps, p.K, xs, scratch);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_rows = p.n_q_rows + 2 * p.n_kv_rows;
    const int pos = p.input_pos[0];
    const int half = p.head_dim / 2;
    // a warp owns the two rows of a rotary pair
    for (int pr = blockIdx.x * kGemvWarps + warp; pr < n_rows / 2; pr += gridDim.x * kGemvWarps) {
        const int r0 = 2 * pr;
        const __nv_bfloat16* rows[2] = {p.w + (size_t)r0 * p.K, p.w + (size_t)(r0 + 1) * p.K};
        float d[2];
        warp_dots<2>(rows, xs, p.K, d);
        if (lane == 0) {
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
    }
}

// ---------------------------------------------------------------- y = [residual +] W x
struct GemvParams {
    const __nv_bfloat16 *x, *norm_w, *w, *residual;
    __nv_bfloat16* out;
    float eps;
    int K, N;
};

__global__ void __launch_bounds__(kGemvThreads) gemv_residual_kernel(const __grid_constant__ GemvParams p) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(smem_raw);
    __shared__ float scratch[kGemvWarps];
    stage_input(p.x, p.norm_w, p.eps, p.K, xs, scratch);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int pr = blockIdx.x * kGemvWarps + warp; pr < p.N / 2; pr += gridDim.x * kGemvWarps) {
        const int r0 = 2 * pr;
        const __nv_bfloat16* rows[2] = {p.w + (size_t)r0 * p.K, p.w + (size_t)(r0 + 1) * p.K};
        float d[2];
        warp_dots<2>(rows, xs, p.K, d);
        if (lane == 0) {
            float a = bf16_round(d[0]), b = bf16_round(d[1]);
            if (p.residual) {  // h = x + attention(...) / out = h + feed_forward(...): bf16 + bf16 -> bf16
                a = __fadd_rn(__bfloat162float(p.residual[r0]), a);
                b = __fadd_rn(__bfloat162float(p.residual[r0 + 1]), b);
            }
            *reinterpret_cast<__nv_bfloat162*>(p.out + r0) = __floats2bfloat162_rn(a, b);
        }
    }
}

// ---------------------------------------------------------------- ffn_norm -> silu(w1 x) * (w3 x)
struct GateParams {
    const __nv_bfloat16 *x, *norm_w, *w1, *w3;
    __nv_bfloat16* out;
    float eps;
    int K, N;
};

__global__ void __launch_bounds__(kGemvThreads) rmsnorm_ffn_gate_kernel(const __grid_constant__ GateParams p) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(smem_raw);
    __shared__ float scratch[kGemvWarps];
    stage_input(p.x, p.norm_w, p.eps, p.K, xs, scratch);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int pr = blockIdx.x * kGemvWarps + warp; pr < p.N / 2; pr += gridDim.x * kGemvWarps) {
        const int r0 = 2 * pr;
        const __nv_bfloat16* rows[4] = {p.w1 + (size_t)r0 * p.K, p.w1 + (size_t)(r0 + 1) * p.K,
                                        p.w3 + (size_t)r0 * p.K, p.w3 + (size_t)(r0 + 1) * p.K};
        float d[4];
        warp_dots<4>(rows, xs, p.K, d);
        if (lane == 0) {
            float o[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float g = bf16_round(d[j]);                       // w1(x), bf16 tensor
                const float sg = bf16_round(g / (1.0f + expf(-g)));     // F.silu in fp32, bf16 out
                o[j] = __fmul_rn(sg, bf16_round(d[2 + j]));             // * w3(x)
            }
            *reinterpret_cast<__nv_bfloat162*>(p.out + r0) = __floats2bfloat162_rn(o[0], o[1]);
        }
    }
}

// --------------------------------------------------------------------------- host launchers

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

static int grid_for(int n_pairs) {
    const int want = (n_pairs + kGemvWarps - 1) / kGemvWarps;
    const int cap = sm_count() * 4;
    return want < cap ? want : cap;
}

template <typename K, typename P>
static int launch(K kernel, const P& p, int n_pairs, int Kdim, cudaStream_t stream, const char* what) {
    const size_t smem = (size_t)Kdim * 2;
    static thread_local const void* raised[8] = {};
    if (smem > 48 * 1024) {  // raise the per-function cap once, to the device limit
        bool done = false;
        for (int i = 0; i < 8; ++i) done |= raised[i] == (const void*)kernel;
        if (!done) {
            int dev = 0, optin = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
            cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
            if (e != cudaSuccess) return set_cuda_error(e, what);
            for (int i = 0; i < 8; ++i)
                if (!raised[i]) { raised[i] = (const void*)kernel; break; }
        }
    }
    kernel<<<grid_for(n_pairs), kGemvThreads, smem, stream>>>(p);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? CC_OK : set_cuda_error(e, what);
}

int rmsnorm_qkv_rope_impl(const void* x, const void* norm_w, float eps, const void* wqkv, const void* bias,
                          const void* freqs, const int32_t* input_pos, void* q, void* k, void* v, int dim,
                          int n_q_heads, int n_kv_heads, int head_dim, cudaStream_t stream) {
    if (!x || !wqkv || !freqs || !input_pos || !q || !k || !v) return set_error(CC_ERR_INVALID_ARG, "cc_rmsnorm_qkv_rope: null argument");
    if (dim % 256 || head_dim % 2) return set_error(CC_ERR_UNSUPPORTED_SHAPE, "cc_rmsnorm_qkv_rope: dim must be a multiple of 256");
    QkvParams p = {(const __nv_bfloat16*)x, (const __nv_bfloat16*)norm_w, (const __nv_bfloat16*)wqkv,
                   (const __nv_bfloat16*)bias, (const __nv_bfloat16*)freqs, input_pos,
                   (__nv_bfloat16*)q, (__nv_bfloat16*)k, (__nv_bfloat16*)v, eps, dim,
                   n_q_heads * head_dim, n_kv_heads * head_dim, head_dim};
    return launch(rmsnorm_qkv_rope_kernel, p, (p.n_q_rows + 2 * p.n_kv_rows) / 2, dim, stream, "cc_rmsnorm_qkv_rope");
}

int gemv_residual_impl(const void* x, const void* norm_w, float eps, const void* w, const void* residual, void* out,
                       int K, int N, cudaStream_t stream) {
    if (!x || !w || !out) return set_error(CC_ERR_INVALID_ARG, "cc_gemv_residual: null argument");
    if (K % 256 || N % 2) return set_error(CC_ERR_UNSUPPORTED_SHAPE, "cc_gemv_residual: K must be a multiple of 256, N even");
    GemvParams p = {(const __nv_bfloat16*)x, (const __nv_bfloat16*)norm_w, (const __nv_bfloat16*)w,
                    (const __nv_bfloat16*)residual, (__nv_bfloat16*)out, eps, K, N};
    return launch(gemv_residual_kernel, p, N / 2, K, stream, "cc_gemv_residual");
}

int rmsnorm_ffn_gate_impl(const void* x, const void* norm_w, float eps, const void* w1, const void* w3, void* out,
                          int K, int N, cudaStream_t stream) {
    if (!x || !w1 || !w3 || !out) return set_error(CC_ERR_INVALID_ARG, "cc_rmsnorm_ffn_gate: null argument");
    if (K % 256 || N % 2) return set_error(CC_ERR_UNSUPPORTED_SHAPE, "cc_rmsnorm_ffn_gate: K must be a multiple of 256, N even");
    GateParams p = {(const __nv_bfloat16*)x, (const __nv_bfloat16*)norm_w, (const __nv_bfloat16*)w1,
                    (const __nv_bfloat16*)w3, (__nv_bfloat16*)out, eps, K, N};
    return launch(rmsnorm_ffn_gate_kernel, p, N / 2, K, stream, "cc_rmsnorm_ffn_gate");
}

}  // namespace cc
