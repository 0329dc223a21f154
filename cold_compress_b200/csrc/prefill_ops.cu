// prefill_ops.cu -- the HBM-bound helpers on the prefill side of the cache:
//   * key L2 norms of cached keys (KVCacheL2.update_state, cache.py:607-612)
//   * prompt-compression gather into the cache (prompt_compression.py:82-88 +
//     KVCache._prefill_update / _fill_contiguous, cache.py:381-401)
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "common.cuh"
#include "internal.h"

namespace cc {

// one row per (D/8)-lane group; canonical dot tree (same bits as the decode kernel's norm)
template <int D>
__global__ void __launch_bounds__(256) key_norms_kernel(const uint16_t* __restrict__ k, __nv_bfloat16* __restrict__ out, int n_rows) {
    constexpr int LPR = D / 8;
    constexpr int RPB = 256 / LPR;
    const int li = threadIdx.x % LPR, sub = threadIdx.x / LPR;
    for (int r0 = blockIdx.x * RPB; r0 < n_rows; r0 += gridDim.x * RPB) {
        const int r = r0 + sub;
        float kf[8];
        uint4 kv = make_uint4(0, 0, 0, 0);
        if (r < n_rows) kv = *reinterpret_cast<const uint4*>(k + (size_t)r * D + li * 8);
        unpack8(kv, kf);
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc = __fmaf_rn(kf[i], kf[i], acc);
#pragma unroll
        for (int d = LPR / 2; d >= 1; d >>= 1) acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, d));
        if (r < n_rows && li == 0) out[r] = __float2bfloat16_rn(__fsqrt_rn(acc));
    }
}

int key_norms_impl(const void* k, void* key_norm, int n_rows, int head_dim, cudaStream_t stream) {
    if (!k || !key_norm || n_rows < 0) return set_error(CC_ERR_INVALID_ARG, "cc_key_norms: bad argument");
    if (n_rows == 0) return CC_OK;
    const int rpb = 256 / (head_dim / 8);
    int blocks = (n_rows + rpb - 1) / rpb;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (head_dim == 128)
        key_norms_kernel<128><<<blocks, 256, 0, stream>>>((const uint16_t*)k, (__nv_bfloat16*)key_norm, n_rows);
    else if (head_dim == 64)
        key_norms_kernel<64><<<blocks, 256, 0, stream>>>((const uint16_t*)k, (__nv_bfloat16*)key_norm, n_rows);
    else
        return set_error(CC_ERR_UNSUPPORTED_SHAPE, "cc_key_norms: head_dim must be 64 or 128");
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? CC_OK : set_cuda_error(e, "cc_key_norms launch");
}

// One 16-byte vector per thread: cache[h, j, :] = prompt[h, keep[h or 0, j], :], K and V together;
// pos / mask written by the first lane of each row.
__global__ void __launch_bounds__(256) prefill_gather_kernel(
    uint4* __restrict__ kc, uint4* __restrict__ vc, int32_t* __restrict__ pos, uint8_t* __restrict__ mask,
    const uint4* __restrict__ kp, const uint4* __restrict__ vp, const int64_t* __restrict__ keep,
    int H, int L, int vec_per_row, int keep_heads, int n_keep, int S, int pos_heads) {
    const long long total = (long long)H * n_keep * vec_per_row;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i % vec_per_row);
        const long long row = i / vec_per_row;
        const int j = (int)(row % n_keep), h = (int)(row / n_keep);
        const int64_t src = keep[(size_t)(keep_heads == 1 ? 0 : h) * n_keep + j];
        const size_t so = ((size_t)h * S + src) * vec_per_row + v;
        const size_t d = ((size_t)h * L + j) * vec_per_row + v;
        kc[d] = kp[so];
        vc[d] = vp[so];
        if (v == 0) {
            mask[(size_t)h * L + j] = 1;
            if (pos_heads != 1 || h == 0) pos[(size_t)(pos_heads == 1 ? 0 : h) * L + j] = (int32_t)src;
        }
    }
}

int prefill_gather_impl(const CcCacheState* st, const void* k_prompt, const void* v_prompt,
                        const int64_t* keep_idx, int keep_heads, int n_keep, int prompt_len, cudaStream_t stream) {
    if (!st || !k_prompt || !v_prompt || !keep_idx) return set_error(CC_ERR_INVALID_ARG, "cc_prefill_gather: null argument");
    if (n_keep < 0 || n_keep > st->max_cache_length || n_keep > prompt_len)
        return set_error(CC_ERR_INVALID_ARG, "cc_prefill_gather: n_keep out of range");
    if (keep_heads != 1 && keep_heads != st->n_kv_heads) return set_error(CC_ERR_INVALID_ARG, "cc_prefill_gather: keep_heads");
    if (st->head_dim % 8) return set_error(CC_ERR_UNSUPPORTED_SHAPE, "cc_prefill_gather: head_dim % 8");
    if (n_keep == 0) return CC_OK;
    const int vpr = st->head_dim / 8;
    const long long total = (long long)st->n_kv_heads * n_keep * vpr;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    prefill_gather_kernel<<<blocks, 256, 0, stream>>>(
        (uint4*)st->k_cache, (uint4*)st->v_cache, st->pos, st->mask, (const uint4*)k_prompt, (const uint4*)v_prompt,
        keep_idx, st->n_kv_heads, st->max_cache_length, vpr, keep_heads, n_keep, prompt_len, st->pos_heads);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? CC_OK : set_cuda_error(e, "cc_prefill_gather launch");
}

// ---------------------------------------------------------------------------------------------------------------
// Quantised KV cache (--cache_bits 2 / 4 / 8).  Between two steps the reference holds the cache as integer codes
// with one (scale, zero point) pair per SLOT -- the group is that slot across all heads and channels
// (quantize_tensor(axis=2) on [1, H, L, D], cache.py:183, quantization_utils.py:4-46) -- and every update_kv
// de-quantises the whole cache, updates it and re-quantises everything with fresh min / max (cache.py:323-338).
// What the next step sees is dequantize(quantize(cache)); this kernel applies that round trip in place to the bf16
// working cache (one CTA per slot, K and V in one launch) and emits the codes and statistics the reference would
// hold.  Every arithmetic step is the reference's bf16 tensor op: fp32 inside, one rounding per op.
__global__ void __launch_bounds__(256) requantize_kernel(uint16_t* __restrict__ kc, uint16_t* __restrict__ vc, int H, int L, int D,
                                                         int n_bit, uint8_t* __restrict__ k_codes, uint8_t* __restrict__ v_codes,
                                                         __nv_bfloat16* __restrict__ k_scales, __nv_bfloat16* __restrict__ k_zeros,
                                                         __nv_bfloat16* __restrict__ v_scales, __nv_bfloat16* __restrict__ v_zeros) {
    const int slot = blockIdx.x, which = blockIdx.y;
    uint16_t* cache = which ? vc : kc;
    uint8_t* codes = which ? v_codes : k_codes;
    __nv_bfloat16* scales = which ? v_scales : k_scales;
    __nv_bfloat16* zeros = which ? v_zeros : k_zeros;
    const int vpr = D / 8, nvec = H * vpr;   // 16-byte vectors of this slot's group
    __shared__ float s_red[2][8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float mn = CUDART_INF_F, mx = -CUDART_INF_F;
    for (int i = tid; i < nvec; i += 256) {
        const int h = i / vpr, c = i - h * vpr;
        float f[8];
        unpack8(*reinterpret_cast<const uint4*>(cache + ((size_t)h * L + slot) * D + c * 8), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) { mn = fminf(mn, f[j]); mx = fmaxf(mx, f[j]); }
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, d));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    }
    if (lane == 0) { s_red[0][warp] = mn; s_red[1][warp] = mx; }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < 8; ++w) { mn = fminf(mn, s_red[0][w]); mx = fmaxf(mx, s_red[1][w]); }
    // scales = (max - min).clamp(min=1e-6) / max_int ; zeros = min + scales * 2^(n-1)   (bf16 ops)
    const float max_int = (float)((1 << n_bit) - 1), half = (float)(1 << (n_bit - 1));
    const float sc = bf16_round(__fdiv_rn(fmaxf(bf16_round(__fsub_rn(mx, mn)), bf16_round(1e-6f)), max_int));
    const float zp = bf16_round(__fadd_rn(mn, bf16_round(__fmul_rn(sc, half))));
    if (tid == 0) {
        if (scales) scales[slot] = __float2bfloat16_rn(sc);
        if (zeros) zeros[slot] = __float2bfloat16_rn(zp);
    }
    for (int i = tid; i < nvec; i += 256) {
        const int h = i / vpr, c = i - h * vpr;
        const size_t off = ((size_t)h * L + slot) * D + c * 8;
        float f[8];
        unpack8(*reinterpret_cast<const uint4*>(cache + off), f);
        uint32_t code[8];
        __nv_bfloat16 o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            // x.sub(min).div(scales).round().clamp_(0, max_int)
            float q = rintf(bf16_round(__fdiv_rn(bf16_round(__fsub_rn(f[j], mn)), sc)));
            q = fminf(fmaxf(q, 0.f), max_int);
            code[j] = (uint32_t)q;
            // (code - 2^(n-1)).mul(scales).add(zeros)
            o[j] = __float2bfloat16_rn(__fadd_rn(bf16_round(__fmul_rn(q - half, sc)), zp));
        }
        *reinterpret_cast<uint4*>(cache + off) = *reinterpret_cast<uint4*>(o);
        if (codes) {   // packed like quantization_utils.pack_low_bit_tensor: consecutive elements, first in the low bits
            if (n_bit == 8) {
                *reinterpret_cast<uint2*>(codes + off) =
                    make_uint2(code[0] | code[1] << 8 | code[2] << 16 | code[3] << 24, code[4] | code[5] << 8 | code[6] << 16 | code[7] << 24);
            } else if (n_bit == 4) {
                *reinterpret_cast<uint32_t*>(codes + off / 2) = code[0] | code[1] << 4 | code[2] << 8 | code[3] << 12 |
                                                                code[4] << 16 | code[5] << 20 | code[6] << 24 | code[7] << 28;
            } else {
                *reinterpret_cast<uint16_t*>(codes + off / 4) = (uint16_t)(code[0] | code[1] << 2 | code[2] << 4 | code[3] << 6 |
                                                                           code[4] << 8 | code[5] << 10 | code[6] << 12 | code[7] << 14);
            }
        }
    }
}

int requantize_cache_impl(void* k_cache, void* v_cache, int H, int L, int D, int n_bit, uint8_t* k_codes, uint8_t* v_codes,
                          void* k_scales, void* k_zeros, void* v_scales, void* v_zeros, cudaStream_t stream) {
    if (!k_cache || !v_cache || H < 1 || L < 1) return set_error(CC_ERR_INVALID_ARG, "cc_requantize_cache: bad argument");
    if (n_bit != 2 && n_bit != 4 && n_bit != 8) return set_error(CC_ERR_INVALID_ARG, "cc_requantize_cache: n_bit must be 2, 4 or 8");
    if (D % 8) return set_error(CC_ERR_UNSUPPORTED_SHAPE, "cc_requantize_cache: head_dim % 8");
    requantize_kernel<<<dim3(L, 2), 256, 0, stream>>>((uint16_t*)k_cache, (uint16_t*)v_cache, H, L, D, n_bit, k_codes, v_codes,
                                                     (__nv_bfloat16*)k_scales, (__nv_bfloat16*)k_zeros,
                                                     (__nv_bfloat16*)v_scales, (__nv_bfloat16*)v_zeros);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? CC_OK : set_cuda_error(e, "cc_requantize_cache launch");
}

}  // namespace cc
