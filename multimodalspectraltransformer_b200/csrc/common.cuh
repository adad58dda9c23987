// Shared helpers for the MMT B200 engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cmath>
#include <string>

namespace mmt {

constexpr int D = 128;            // d_model (config_V8 hidden_size); the kernels are specialised for it
constexpr int VOCAB_MAX = 64;     // out_size 43 <= 64 (two values per lane in the sampler)
constexpr int PAGE_TOKENS = 16;   // tokens per self-attention KV page
constexpr int CP_SMAX = 200;      // row stride of the ragged-encoder index maps (>= longest modality sequence, 193)

#define MMT_NEG_INF (-INFINITY)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-serialization attribute may
// start while its predecessor in the stream is still running.  pdl_launch_dependents() lets the successor be
// scheduled; pdl_wait() blocks until the predecessor grid has completed and its writes are visible.  Everything a
// kernel does before pdl_wait() must only touch data that is constant during the decode loop (weights, indices).
// Both are no-ops when the kernel was launched without the attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// one warp, lane owns 4 consecutive columns of a 128-wide row: LayerNorm(v) * gamma + beta
__device__ __forceinline__ float4 ln_row(float4 v, const float* gamma, const float* beta, float eps, int lane) {
    const float mean = warp_sum(v.x + v.y + v.z + v.w) * (1.0f / D);
    const float dx = v.x - mean, dy = v.y - mean, dz = v.z - mean, dw = v.w - mean;
    const float var = warp_sum(dx * dx + dy * dy + dz * dz + dw * dw) * (1.0f / D);
    const float rstd = rsqrtf(var + eps);
    const float4 ga = *reinterpret_cast<const float4*>(gamma + lane * 4);
    const float4 be = *reinterpret_cast<const float4*>(beta + lane * 4);
    return make_float4(dx * rstd * ga.x + be.x, dy * rstd * ga.y + be.y, dz * rstd * ga.z + be.z, dw * rstd * ga.w + be.w);
}

// One key / value row of a head (8 elements) in the KV caches: fp32 in the fp32 check mode, bf16 in
// the tensor-core mode (halves the HBM stream that bounds the decode step).  `Raw` is what one lane
// loads (kept un-converted while the loads are in flight), `unpack` widens it to fp32.
template <typename T> struct KvRow;
template <> struct KvRow<float> {
    struct Raw { float4 a, b; };
    static __device__ __forceinline__ Raw ld(const float* p) {
        Raw r; r.a = reinterpret_cast<const float4*>(p)[0]; r.b = reinterpret_cast<const float4*>(p)[1]; return r;
    }
    static __device__ __forceinline__ void unpack(const Raw& r, float (&v)[8]) {
        v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w; v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w;
    }
    static __device__ __forceinline__ void st(float* p, float x) { *p = x; }
    static __device__ __forceinline__ Raw pack(const float* x) {      // the row as the cache will hold it
        Raw r; r.a = make_float4(x[0], x[1], x[2], x[3]); r.b = make_float4(x[4], x[5], x[6], x[7]); return r;
    }
};
template <> struct KvRow<__nv_bfloat16> {
    typedef uint4 Raw;
    static __device__ __forceinline__ Raw ld(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }
    static __device__ __forceinline__ void unpack(const Raw& r, float (&v)[8]) {
        v[0] = __uint_as_float(r.x << 16); v[1] = __uint_as_float(r.x & 0xffff0000u);
        v[2] = __uint_as_float(r.y << 16); v[3] = __uint_as_float(r.y & 0xffff0000u);
        v[4] = __uint_as_float(r.z << 16); v[5] = __uint_as_float(r.z & 0xffff0000u);
        v[6] = __uint_as_float(r.w << 16); v[7] = __uint_as_float(r.w & 0xffff0000u);
    }
    static __device__ __forceinline__ void st(__nv_bfloat16* p, float x) { *p = __float2bfloat16_rn(x); }
    static __device__ __forceinline__ Raw pack(const float* x) {      // the row as the cache will hold it (same rounding as st)
        Raw r;
        __nv_bfloat162 a = __floats2bfloat162_rn(x[0], x[1]), b = __floats2bfloat162_rn(x[2], x[3]);
        __nv_bfloat162 c = __floats2bfloat162_rn(x[4], x[5]), d = __floats2bfloat162_rn(x[6], x[7]);
        r.x = *reinterpret_cast<uint32_t*>(&a); r.y = *reinterpret_cast<uint32_t*>(&b);
        r.z = *reinterpret_cast<uint32_t*>(&c); r.w = *reinterpret_cast<uint32_t*>(&d);
        return r;
    }
};

__host__ __device__ __forceinline__ int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace mmt
