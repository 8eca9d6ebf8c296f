// Ceiling probe for gathers that HIT in the SM (what a locality-preserving row order can reach): random row gathers
// out of a per-block region small enough to stay in L1, or out of shared memory.
//   * l1 mode: W bytes per lane (4 / 8 / 16), G lanes per row; rows of G*W bytes, 8 independent loads in flight
//   * smem mode: the region is staged in shared memory and gathered with LDS.128
// Also prints %nsmid and the set of %smid values the grid saw (the CSR kernel keys its home supertickets by %smid).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/l1_gather_probe tools/l1_gather_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <set>

template <typename V, int G, int UNROLL>
__global__ void __launch_bounds__(256, 4) l1_probe(const V *B, unsigned region_rows, int iters, float *sink) {
    const unsigned lane = threadIdx.x & 31, sub = lane / G, l = lane % G;
    unsigned state = ((blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5)) * (32 / G) + sub) * 2654435761u + 777u;
    const V *base = B + (size_t)blockIdx.x * region_rows * G;
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        V v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            state = state * 1664525u + 1013904223u;
            const unsigned row = (unsigned)(((unsigned long long)state * region_rows) >> 32);
            v[u] = __ldg(base + (size_t)row * G + l);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) acc += *reinterpret_cast<float *>(&v[u]);
    }
    if (acc == 123.456f) *sink = acc;
}

template <int G, int UNROLL>
__global__ void __launch_bounds__(1024, 1) smem_probe(const float4 *B, unsigned region_rows, int iters, float *sink) {
    extern __shared__ float4 tile[];
    const unsigned lane = threadIdx.x & 31, sub = lane / G, l = lane % G;
    for (unsigned i = threadIdx.x; i < region_rows * G; i += blockDim.x) tile[i] = B[(size_t)blockIdx.x * region_rows * G + i];
    __syncthreads();
    unsigned state = ((blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5)) * (32 / G) + sub) * 2654435761u + 777u;
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        float4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            state = state * 1664525u + 1013904223u;
            const unsigned row = (unsigned)(((unsigned long long)state * region_rows) >> 32);
            v[u] = tile[row * G + l];
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) acc += v[u].x;
    }
    if (acc == 123.456f) *sink = acc;
}

__global__ void smid_probe(unsigned *out) {
    unsigned smid, nsmid;
    asm("mov.u32 %0, %%smid;" : "=r"(smid));
    asm("mov.u32 %0, %%nsmid;" : "=r"(nsmid));
    if (threadIdx.x == 0) { out[2 * blockIdx.x] = smid; out[2 * blockIdx.x + 1] = nsmid; }
}

template <typename V, int G> static void run_l1(const V *B, int region_kb, float *sink, int sms) {
    int bps = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, l1_probe<V, G, 8>, 256, 0);
    const int blocks = bps * sms, iters = 2000;
    const unsigned region_rows = (unsigned)(region_kb * 1024 / (sizeof(V) * G));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    l1_probe<V, G, 8><<<blocks, 256>>>(B, region_rows, iters, sink);
    cudaEventRecord(e0);
    l1_probe<V, G, 8><<<blocks, 256>>>(B, region_rows, iters, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double gathered = (double)blocks * 256 * iters * 8 * sizeof(V);
    printf("L1   %2zu B/lane, row %4zu B (G=%2d), %3d KB/block x %d blocks/SM: %6.2f TB/s = %5.1f B/clk/SM @1.965 GHz (%.3f ms)\n",
           sizeof(V), sizeof(V) * G, G, region_kb, bps, gathered / ms / 1e9, gathered / ms / 1e9 * 1e12 / sms / 1.965e9, ms);
}

template <int G> static void run_smem(const float4 *B, int region_kb, float *sink, int sms) {
    const int iters = 2000;
    const unsigned region_rows = (unsigned)(region_kb * 1024 / (16 * G));
    cudaFuncSetAttribute(smem_probe<G, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, region_kb * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    smem_probe<G, 8><<<sms, 1024, region_kb * 1024>>>(B, region_rows, iters, sink);
    cudaEventRecord(e0);
    smem_probe<G, 8><<<sms, 1024, region_kb * 1024>>>(B, region_rows, iters, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double gathered = (double)sms * 1024 * iters * 8 * 16;
    printf("SMEM 16 B/lane, row %4d B (G=%2d), %3d KB/block x 1 block/SM:  %6.2f TB/s = %5.1f B/clk/SM @1.965 GHz (%.3f ms)\n",
           16 * G, G, region_kb, gathered / ms / 1e9, gathered / ms / 1e9 * 1e12 / sms / 1.965e9, ms);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    const size_t bytes = (size_t)256 << 20;
    void *B;
    float *sink;
    cudaMalloc(&B, bytes);
    cudaMalloc(&sink, 4);
    cudaMemset(B, 0, bytes);
    {
        unsigned *d, h[2 * 1024];
        cudaMalloc(&d, sizeof h);
        smid_probe<<<1024, 32>>>(d);
        cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
        std::set<unsigned> ids;
        unsigned mx = 0;
        for (int i = 0; i < 1024; ++i) { ids.insert(h[2 * i]); mx = h[2 * i] > mx ? h[2 * i] : mx; }
        printf("multiProcessorCount %d, %%nsmid %u, distinct %%smid seen %zu, max %%smid %u\n", sms, h[1], ids.size(), mx);
    }
    for (int kb : {8, 32}) {
        run_l1<float4, 4>((const float4 *)B, kb, sink, sms);
        run_l1<float4, 8>((const float4 *)B, kb, sink, sms);
        run_l1<float4, 16>((const float4 *)B, kb, sink, sms);
        run_l1<float4, 32>((const float4 *)B, kb, sink, sms);
        run_l1<float2, 16>((const float2 *)B, kb, sink, sms);
        run_l1<float, 32>((const float *)B, kb, sink, sms);
    }
    run_smem<4>((const float4 *)B, 128, sink, sms);
    run_smem<8>((const float4 *)B, 128, sink, sms);
    run_smem<32>((const float4 *)B, 128, sink, sms);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
