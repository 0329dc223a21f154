// internal.h -- declarations shared by the translation units behind the C ABI.
#pragma once
#include <cuda_runtime.h>

#include "../../include/cc_b200.h"

namespace cc {
bool pdl_enabled();                                       // programmatic dependent launch (CC_B200_PDL=0 turns it off)
int set_error(int code, const char* msg);                 // records msg for cc_last_error(), returns code
int set_cuda_error(cudaError_t e, const char* where);     // CC_ERR_CUDA with the runtime's message
int decode_step_impl(const CcCacheState* st, const CcDecodeArgs* a, cudaStream_t stream);
int decode_cluster_size_impl(const CcCacheState* st);
void set_trace_buffer(unsigned long long* buf);
int key_norms_impl(const void* k, void* key_norm, int n_rows, int head_dim, cudaStream_t stream);
int prefill_gather_impl(const CcCacheState* st, const void* k_prompt, const void* v_prompt,
                        const int64_t* keep_idx, int keep_heads, int n_keep, int prompt_len,
                        cudaStream_t stream);
int prefill_attention_impl(const CcPrefillAttnArgs* a, cudaStream_t stream);
unsigned long long prefill_attention_workspace_impl(int S, int Hkv, int G, int obs_len, int n_windows);
int requantize_cache_impl(void* k_cache, void* v_cache, int H, int L, int D, int n_bit, uint8_t* k_codes, uint8_t* v_codes,
                          void* k_scales, void* k_zeros, void* v_scales, void* v_zeros, cudaStream_t stream);
int rmsnorm_qkv_rope_impl(const void* x, const void* norm_w, float eps, const void* wqkv, const void* bias,
                          const void* freqs, const int32_t* input_pos, void* q, void* k, void* v, int dim,
                          int n_q_heads, int n_kv_heads, int head_dim, const CcTpArgs* tp, cudaStream_t stream);
int gemv_residual_impl(const void* x, const void* norm_w, float eps, const void* w, const void* residual, void* out,
                       int K, int N, const CcTpArgs* tp, cudaStream_t stream);
int rmsnorm_ffn_gate_impl(const void* x, const void* norm_w, float eps, const void* w1, const void* w3, void* out,
                          int K, int N, const CcTpArgs* tp, cudaStream_t stream);
// whole-token persistent kernel (decode_token.cu)
unsigned long long token_scratch_bytes_impl(const CcTokenModel* m);
int token_plan_create_impl(const CcTokenModel* m, void** plan);
int token_plan_destroy_impl(void* plan);
int decode_token_impl(void* plan, const int32_t* token_in, const int32_t* input_pos, int32_t* token_out, void* logits_out,
                      void* probs_out, cudaStream_t stream);
int token_buffer_impl(void* plan, int layer, int which, uint64_t* off, int32_t* words);
int token_set_trace_impl(void* plan, void* buf);
int token_info_impl(void* plan, int32_t* n_ctas, int32_t* n_ops, int32_t* ring_stages, int32_t* smem);
int decode_step_via_token_impl(const CcCacheState* st, const CcDecodeArgs* a, cudaStream_t stream);
}  // namespace cc
