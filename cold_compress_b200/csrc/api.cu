// api.cu -- the extern "C" surface declared in include/cc_b200.h.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include "internal.h"

#define CC_STR2(x) #x
#define CC_STR(x) CC_STR2(x)

#include <stdlib.h>

namespace cc {
bool pdl_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("CC_B200_PDL");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on == 1;
}
static thread_local char g_err[512] = "";
int set_error(int code, const char* msg) {
    strncpy(g_err, msg, sizeof(g_err) - 1);
    g_err[sizeof(g_err) - 1] = 0;
    return code;
}
int set_cuda_error(cudaError_t e, const char* where) {
    snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
    return CC_ERR_CUDA;
}
}  // namespace cc

extern "C" {

int cc_abi_version(void) { return CC_ABI_VERSION; }

const char* cc_build_info(void) {
    return "cold_compress_b200 sm_100a nvcc " CC_STR(__CUDACC_VER_MAJOR__) "." CC_STR(__CUDACC_VER_MINOR__);
}

const char* cc_last_error(void) { return cc::g_err; }

int cc_sync_words(void) { return 4; }

int cc_decode_step(const CcCacheState* state, const CcDecodeArgs* args, void* stream) {
    return cc::decode_step_impl(state, args, (cudaStream_t)stream);
}

int cc_decode_cluster_size(const CcCacheState* state) { return cc::decode_cluster_size_impl(state); }

void cc_debug_set_trace(void* device_buffer) { cc::set_trace_buffer((unsigned long long*)device_buffer); }

int cc_key_norms(const void* k, void* key_norm, int32_t n_rows, int32_t head_dim, void* stream) {
    return cc::key_norms_impl(k, key_norm, n_rows, head_dim, (cudaStream_t)stream);
}

int cc_prefill_gather(const CcCacheState* state, const void* k_prompt, const void* v_prompt,
                      const int64_t* keep_idx, int32_t keep_heads, int32_t n_keep, int32_t prompt_len,
                      void* stream) {
    return cc::prefill_gather_impl(state, k_prompt, v_prompt, keep_idx, keep_heads, n_keep, prompt_len,
                                   (cudaStream_t)stream);
}

int cc_requantize_cache(void* k_cache, void* v_cache, int32_t n_kv_heads, int32_t max_cache_length, int32_t head_dim, int32_t n_bit,
                        uint8_t* k_codes, uint8_t* v_codes, void* k_scales, void* k_zero_points, void* v_scales,
                        void* v_zero_points, void* stream) {
    return cc::requantize_cache_impl(k_cache, v_cache, n_kv_heads, max_cache_length, head_dim, n_bit, k_codes, v_codes, k_scales,
                                     k_zero_points, v_scales, v_zero_points, (cudaStream_t)stream);
}

uint64_t cc_prefill_attention_workspace(int32_t seq_len, int32_t n_kv_heads, int32_t q_per_kv, int32_t obs_len,
                                        int32_t n_windows) {
    return cc::prefill_attention_workspace_impl(seq_len, n_kv_heads, q_per_kv, obs_len, n_windows);
}

int cc_prefill_attention(const CcPrefillAttnArgs* args, void* stream) {
    return cc::prefill_attention_impl(args, (cudaStream_t)stream);
}

int cc_rmsnorm_qkv_rope(const void* x, const void* norm_w, float eps, const void* wqkv, const void* bias,
                        const void* freqs_cis, const int32_t* input_pos, void* q, void* k, void* v,
                        int32_t dim, int32_t n_q_heads, int32_t n_kv_heads, int32_t head_dim,
                        const CcTpArgs* tp, void* stream) {
    return cc::rmsnorm_qkv_rope_impl(x, norm_w, eps, wqkv, bias, freqs_cis, input_pos, q, k, v, dim, n_q_heads,
                                     n_kv_heads, head_dim, tp, (cudaStream_t)stream);
}

int cc_gemv_residual(const void* x, const void* norm_w, float eps, const void* w, const void* residual,
                     void* out, int32_t K, int32_t N, const CcTpArgs* tp, void* stream) {
    return cc::gemv_residual_impl(x, norm_w, eps, w, residual, out, K, N, tp, (cudaStream_t)stream);
}

int cc_rmsnorm_ffn_gate(const void* x, const void* norm_w, float eps, const void* w1, const void* w3,
                        void* out, int32_t K, int32_t N, const CcTpArgs* tp, void* stream) {
    return cc::rmsnorm_ffn_gate_impl(x, norm_w, eps, w1, w3, out, K, N, tp, (cudaStream_t)stream);
}

uint64_t cc_token_scratch_bytes(const CcTokenModel* model) { return cc::token_scratch_bytes_impl(model); }
int cc_token_plan_create(const CcTokenModel* model, void** plan) { return cc::token_plan_create_impl(model, plan); }
int cc_token_plan_destroy(void* plan) { return cc::token_plan_destroy_impl(plan); }
int cc_decode_token(void* plan, const int32_t* token_in, const int32_t* input_pos, int32_t* token_out, void* logits_out,
                    void* probs_out, void* stream) {
    return cc::decode_token_impl(plan, token_in, input_pos, token_out, logits_out, probs_out, (cudaStream_t)stream);
}
int cc_token_buffer(void* plan, int32_t layer, int32_t which, uint64_t* offset_bytes, int32_t* n_words) {
    return cc::token_buffer_impl(plan, layer, which, offset_bytes, n_words);
}
int cc_token_set_trace(void* plan, void* device_buffer) { return cc::token_set_trace_impl(plan, device_buffer); }
int cc_token_info(void* plan, int32_t* n_ctas, int32_t* n_stages_program, int32_t* ring_stages, int32_t* smem_bytes) {
    return cc::token_info_impl(plan, n_ctas, n_stages_program, ring_stages, smem_bytes);
}

}  // extern "C"
