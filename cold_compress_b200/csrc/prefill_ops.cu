// prefill_ops.cu -- the HBM-bound helpers on the prefill side of the cache:
//   * key L2 norms of cached keys (KVCacheL2.update_state, cache.py:607-612)
//   * prompt-compression gather into the cache (prompt_compression.py:82-88 +
//     KVCache._prefill_update / _fill_contiguous, cache.py:381-401)
#include <cuda_runtime.h>
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

}  // namespace cc
