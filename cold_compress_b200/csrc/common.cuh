// common.cuh -- device helpers for the sm_100a compressed-KV kernels:
// bf16 rounding, the bit-reproducible exp / dot definitions shared with the test oracle's
// restatement, mbarrier + TMA bulk copies (UBLKCP), cluster barriers and DSMEM stores.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cc {

// ------------------------------------------------------------------ bf16
__device__ __forceinline__ float bf16_round(float x) {
    return __bfloat162float(__float2bfloat16_rn(x));
}
__device__ __forceinline__ float bf16_bits_to_float(uint32_t lo16) {
    return __uint_as_float(lo16 << 16);
}
// unpack 8 bf16 (one 16-byte vector) into fp32
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
    f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
    f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
    f[4] = __uint_as_float(v.z << 16); f[5] = __uint_as_float(v.z & 0xffff0000u);
    f[6] = __uint_as_float(v.w << 16); f[7] = __uint_as_float(v.w & 0xffff0000u);
}

// ------------------------------------------------ exp(x) for x <= 0, bit-reproducible
// n = rint(x*log2e); r = x - n*ln2 in two fma steps; degree-6 polynomial; exact 2^n scaling in
// two steps so results below 2^-126 round once.  Only *_rn intrinsics: ptxas cannot contract
// or reorder them, so every IEEE machine evaluating the same sequence gets the same bits.
__device__ __forceinline__ float det_expf(float x) {
    if (!(x > -104.0f)) return 0.0f;
    float t = __fmul_rn(x, 1.44269504088896341f);
    float n = rintf(t);
    float r = __fmaf_rn(n, -0.693359375f, x);
    r = __fmaf_rn(n, 2.12194440e-4f, r);
    float p = 1.9875691500e-4f;
    p = __fmaf_rn(p, r, 1.3981999507e-3f);
    p = __fmaf_rn(p, r, 8.3334519073e-3f);
    p = __fmaf_rn(p, r, 4.1665795894e-2f);
    p = __fmaf_rn(p, r, 1.6666665459e-1f);
    p = __fmaf_rn(p, r, 5.0000001201e-1f);
    float r2 = __fmul_rn(r, r);
    p = __fmaf_rn(p, r2, r);
    p = __fadd_rn(p, 1.0f);
    int ni = (int)n;
    int n1 = ni < -100 ? -100 : ni;
    int n2 = ni - n1;
    float s1 = __uint_as_float((uint32_t)(n1 + 127) << 23);
    float s2 = __uint_as_float((uint32_t)(n2 + 127) << 23);
    return __fmul_rn(__fmul_rn(p, s1), s2);
}

// order-preserving map float -> uint32 (for packed (score, index) arg-min keys)
__device__ __forceinline__ uint32_t float_order_key(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// ------------------------------------------------------------------ shared / cluster
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
    uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r;
}
__device__ __forceinline__ void cluster_arrive_relaxed() {
    asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_arrive_release() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait_acquire() {
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of `local_smem_addr` in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t map_to_rank(uint32_t local_smem_addr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" :: "r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_u64(uint32_t addr, uint64_t v) {
    asm volatile("st.shared::cluster.u64 [%0], %1;" :: "r"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_v4f32(uint32_t addr, float4 v) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Asynchronous store into a peer CTA's shared memory that also signals the peer's mbarrier with the
// number of bytes written (st.async ... mbarrier::complete_tx): data and "it has arrived" in one
// message, without a cluster-wide barrier or a CTA-wide release fence on the sender.
__device__ __forceinline__ void st_async_u32(uint32_t remote_addr, uint32_t v, uint32_t remote_mbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.u32 [%0], %1, [%2];"
                 :: "r"(remote_addr), "r"(v), "r"(remote_mbar) : "memory");
}
__device__ __forceinline__ void st_async_u64(uint32_t remote_addr, uint64_t v, uint32_t remote_mbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.u64 [%0], %1, [%2];"
                 :: "r"(remote_addr), "l"(v), "r"(remote_mbar) : "memory");
}
__device__ __forceinline__ void st_async_v4f32(uint32_t remote_addr, float4 v, uint32_t remote_mbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];"
                 :: "r"(remote_addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(remote_mbar) : "memory");
}

// ------------------------------------------------------------------ mbarrier + TMA bulk copy
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk copy global -> this CTA's shared memory through the TMA engine (SASS: UBLKCP).
// dst and src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        :: "r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// The same with an L2 eviction-priority hint (createpolicy): a stream that is read exactly once (weights) should
// not displace what the kernel re-reads (its own instructions, stage descriptors, exchange words).
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void tma_load_1d_hint(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        :: "r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
// fire-and-forget request to bring [ptr, ptr+bytes) into L2 (bytes a multiple of 16)
__device__ __forceinline__ void l2_prefetch_bulk(const void* ptr, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(ptr), "r"(bytes) : "memory");
}
// order earlier generic-proxy accesses to shared memory before later async-proxy (TMA) writes
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ------------------------------------------------------------------ programmatic dependent launch
// A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while its
// predecessor in the stream is still running.  Everything before pdl_wait() must only touch memory
// no in-flight predecessor writes (weights, the K/V cache, policy state of this layer);
// pdl_wait() returns when the predecessor grid has completed and its writes are visible.
// Without the launch attribute both are no-ops.
__device__ __forceinline__ void pdl_launch_dependents() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

// ------------------------------------------------------------------ global memory, gpu scope
__device__ __forceinline__ int32_t ld_acquire_gpu(const int32_t* p) {
    int32_t v; asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void red_release_gpu_add(int32_t* p, int32_t v) {
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release_gpu(int32_t* p, int32_t v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

// ------------------------------------------------------------------ warp-level reductions shared by the decode kernels
template <int LPR>
__device__ __forceinline__ float group_butterfly(float v) {
#pragma unroll
    for (int d = LPR / 2; d >= 1; d >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, d));
    return v;
}

// Reduce G per-lane partial sums over the LPR lanes of a row group with the SAME addition tree as
// group_butterfly (pairs at lane distance LPR/2, ..., 1) but exchanging halves of the value set while
// more than one value is held: 5 shuffles instead of 16 for G = 4.  On return lane `li` holds the
// full sum of query head `*g_out` (replicated over the lanes that share it); fp add is commutative,
// so the bits equal the plain butterfly's.
template <int LPR, int G>
__device__ __forceinline__ float group_reduce_multi(const float (&v)[G], int li, int* g_out) {
    constexpr int P0 = G <= 1 ? 1 : G <= 2 ? 2 : G <= 4 ? 4 : 8;
    static_assert(P0 <= LPR, "more query heads than lanes per row");
    float w[P0];
#pragma unroll
    for (int i = 0; i < P0; ++i) w[i] = i < G ? v[i] : 0.f;
    int g = 0;
    int held = P0;
#pragma unroll
    for (int d = LPR / 2; d >= 1; d >>= 1) {
        if (held > 1) {
            const int half = held / 2;
            const bool upper = (li & d) != 0;  // upper lanes keep the upper half of the values
#pragma unroll
            for (int i = 0; i < P0 / 2; ++i) {
                if (i < half) {
                    const float send = upper ? w[i] : w[i + half];
                    const float keep = upper ? w[i + half] : w[i];
                    w[i] = __fadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, d));
                }
            }
            if (upper) g += half;
            held = half;
        } else {
            w[0] = __fadd_rn(w[0], __shfl_xor_sync(0xffffffffu, w[0], d));
        }
    }
    *g_out = g;
    return w[0];
}

__device__ __forceinline__ uint64_t warp_min_u64(uint64_t k) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        uint64_t o = __shfl_xor_sync(0xffffffffu, k, d);
        k = o < k ? o : k;
    }
    return k;
}

}  // namespace cc
