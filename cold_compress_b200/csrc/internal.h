// internal.h -- declarations shared by the translation units behind the C ABI.
#pragma once
#include <cuda_runtime.h>

#include "../../include/cc_b200.h"

namespace cc {
int set_error(int code, const char* msg);                 // records msg for cc_last_error(), returns code
int set_cuda_error(cudaError_t e, const char* where);     // CC_ERR_CUDA with the runtime's message
int decode_step_impl(const CcCacheState* st, const CcDecodeArgs* a, cudaStream_t stream);
int decode_cluster_size_impl(const CcCacheState* st);
int key_norms_impl(const void* k, void* key_norm, int n_rows, int head_dim, cudaStream_t stream);
int prefill_gather_impl(const CcCacheState* st, const void* k_prompt, const void* v_prompt,
                        const int64_t* keep_idx, int keep_heads, int n_keep, int prompt_len,
                        cudaStream_t stream);
}  // namespace cc
