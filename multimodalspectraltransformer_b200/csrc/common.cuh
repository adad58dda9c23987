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

__host__ __device__ __forceinline__ int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace mmt
