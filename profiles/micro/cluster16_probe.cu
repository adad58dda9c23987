// Micro-probe for the cluster decode kernel: can 8 clusters of 16 CTAs x 1024 threads x ~214 KB of
// shared memory be co-resident on a B200, and what do the two cluster-wide exchanges of the FFN phase cost?
//   (a) cudaOccupancyMaxActiveClusters for that shape
//   (b) cycles of: cluster.sync alone | all-gather of 2 x 256 B rows to 16 CTAs + sync | scatter of a 32 x 128 fp32 tile
//       (2 rows to each of 16 owners) + sync
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cluster16_probe cluster16_probe.cu ; run: ./cluster16_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int THREADS = 1024;
constexpr int SMEM = 214 * 1024;

__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cta_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t map_remote(const void* p, uint32_t rank) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(p), r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_remote_v4(uint32_t addr, uint4 v) {
    asm volatile("st.shared::cluster.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_remote_v2(uint32_t addr, float2 v) {
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1,%2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}

template <int CL>
__global__ void __launch_bounds__(THREADS, 1) probe(long long* out, int iters) {
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned char* xg = smem;                 // [32][256 B]
    float* recv = reinterpret_cast<float*>(smem + 8192);   // [16][2][128]
    const uint32_t c = cta_rank();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    cluster_sync();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) cluster_sync();
    long long t1 = clock64();
    for (int i = 0; i < iters; ++i) {         // all-gather: warp w < 16 writes this CTA's two rows into CTA w
        if (warp < CL) {  // (rows of the CTAs beyond the cluster size do not exist)
            const int r = lane >> 4, ch = lane & 15;
            st_remote_v4(map_remote(xg + (2 * c + r) * 256 + ch * 16, warp), make_uint4(i, c, r, ch));
        }
        cluster_sync();
    }
    long long t2 = clock64();
    for (int i = 0; i < iters; ++i) {         // scatter: warp (mt, nt) holds rows g, g+8 of m-tile mt, cols nt*8 + 2t
        const int mt = warp >> 4, nt = warp & 15, g = lane >> 2, t = lane & 3;
        for (int hh = 0; hh < 2; ++hh) {
            const int R = mt * 16 + g + hh * 8;
            if (R < 2 * CL) st_remote_v2(map_remote(recv + ((c * 2) + (R & 1)) * 128 + nt * 8 + 2 * t, R >> 1), make_float2((float)i, (float)R));
        }
        cluster_sync();
    }
    long long t3 = clock64();
    if (threadIdx.x == 0) {
        long long* o = out + (size_t)blockIdx.x * 4;
        o[0] = (t1 - t0) / iters; o[1] = (t2 - t1) / iters; o[2] = (t3 - t2) / iters; o[3] = recv[5] > 1e30f;
    }
}

template <int CL>
void run(int ncl) {
    cudaFuncSetAttribute(probe<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    cudaFuncSetAttribute(probe<CL>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ncl * CL); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = SMEM;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int maxc = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&maxc, probe<CL>, &cfg);
    printf("cluster size %d x %d clusters: cudaOccupancyMaxActiveClusters -> %d (%s)\n", CL, ncl, maxc, cudaGetErrorString(e));
    long long* d; cudaMalloc(&d, ncl * CL * 4 * sizeof(long long));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    int iters = 200;
    cudaLaunchKernelEx(&cfg, probe<CL>, d, iters);
    cudaEventRecord(a); cudaLaunchKernelEx(&cfg, probe<CL>, d, iters); cudaEventRecord(b);
    e = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    static long long h[4 * 1024]; cudaMemcpy(h, d, ncl * CL * 4 * sizeof(long long), cudaMemcpyDeviceToHost);
    printf("  launch %s, %.1f us total; CTA 0: cluster.sync %lld cyc | all-gather+sync %lld | scatter+sync %lld ; last CTA: %lld %lld %lld\n",
           cudaGetErrorString(e), ms * 1e3, h[0], h[1], h[2], h[(ncl * CL - 1) * 4], h[(ncl * CL - 1) * 4 + 1], h[(ncl * CL - 1) * 4 + 2]);
    cudaFree(d);
}

int main() {
    run<16>(7); run<16>(8); run<8>(16); run<8>(18); run<4>(32); run<2>(64);
    return 0;
}
