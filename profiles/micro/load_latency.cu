// Micro-benchmark: latency of one dependent 16-byte-per-lane global load round under the decode_attn launch shape
// (128 CTAs x 1024 threads), data resident in L2 vs streamed from DRAM, few vs all warps loading.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o load_latency load_latency.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(1024, 1) chase(const uint4* buf, uint64_t n16, int rounds, int active_warps, long long* out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp >= active_warps) return;
    uint64_t idx = ((uint64_t)blockIdx.x * 32 + warp) * 7919u % (n16 / 32);
    uint32_t acc = 0;
    const long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
        const uint4 v = buf[idx * 32 + lane];           // 512 contiguous bytes per warp
        acc += v.x;
        const uint32_t next = __shfl_sync(0xffffffffu, v.y, 0);
        idx = (idx * 1103515245u + 12345u + next) % (n16 / 32);   // depends on the loaded value
    }
    const long long t1 = clock64();
    if (lane == 0 && warp == 5 % active_warps && (blockIdx.x == 64 % gridDim.x)) { out[0] = (t1 - t0) / rounds; out[1] = acc; }
}

int main() {
    long long* d_out; cudaMalloc(&d_out, 16);
    for (int big = 0; big < 2; ++big) {
        const uint64_t bytes = big ? (2ull << 30) : (8ull << 20);
        uint4* buf; cudaMalloc(&buf, bytes); cudaMemset(buf, 0, bytes);
        for (int aw : {1, 8, 32}) for (int grid : {16, 128}) {
            for (int rep = 0; rep < 2; ++rep) chase<<<grid, 1024>>>(buf, bytes / 16, 16, aw, d_out);
            cudaDeviceSynchronize();
            long long h[2]; cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
            printf("%s buffer, grid %3d, %2d warps/CTA loading: %lld cycles per dependent round\n", big ? "2 GiB (DRAM)" : "8 MiB (L2)  ", grid, aw, h[0]);
        }
        cudaFree(buf);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
