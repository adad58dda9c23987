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

__host__ __device__ __forceinline__ int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace mmt
