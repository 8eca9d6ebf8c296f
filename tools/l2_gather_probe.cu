// Ceiling probe for the access pattern that bounds the SpMM on dense-ish graphs: random ROW gathers out of an
// L2-resident matrix, 16 bytes per lane, G lanes per row, UNROLL independent loads in flight, no index stream and
// (almost) no arithmetic.  Prints the sustained gather rate per row size - the practical L2 -> SM ceiling the
// CSR kernel's "gather TB/s" column is to be read against.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/l2_gather_probe tools/l2_gather_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

template <int G, int UNROLL>
__global__ void __launch_bounds__(256, 4) gather_probe(const float4 *B, unsigned n_rows, unsigned row_words, int iters,
                                                       float *sink) {
    const unsigned lane = threadIdx.x & 31, sub = lane / G, l = lane % G;
    unsigned state = (blockIdx.x * blockDim.x + threadIdx.x) / G * 2654435761u + 12345u + sub;
    // all G lanes of a group must draw the same row: seed per group
    state = ((blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5)) * (32 / G) + sub) * 2654435761u + 777u;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int it = 0; it < iters; ++it) {
        float4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            state = state * 1664525u + 1013904223u;
            const unsigned row = (unsigned)(((unsigned long long)state * n_rows) >> 32);
            v[u] = __ldg(B + (size_t)row * row_words + l);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
    if (acc.x + acc.y + acc.z + acc.w == 123.456f) *sink = acc.x;
}

template <int G> static void run(const float4 *B, size_t bytes, float *sink) {
    const unsigned row_words = G, n_rows = (unsigned)(bytes / (16 * G));
    int bps = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, gather_probe<G, 8>, 256, 0);
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int blocks = bps * p.multiProcessorCount, iters = 400;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    gather_probe<G, 8><<<blocks, 256>>>(B, n_rows, row_words, iters, sink);   // warm-up: pulls the matrix into L2
    gather_probe<G, 8><<<blocks, 256>>>(B, n_rows, row_words, iters, sink);
    cudaEventRecord(e0);
    gather_probe<G, 8><<<blocks, 256>>>(B, n_rows, row_words, iters, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double gathered = (double)blocks * 256 * iters * 8 * 16;
    printf("row %4d B (G=%2d): %6.2f TB/s  (%d blocks x 256 threads, %d resident blocks/SM, %.3f ms, matrix %.0f MB)\n",
           16 * G, G, gathered / ms / 1e9, blocks, bps, ms, bytes / 1e6);
}

int main(int argc, char **argv) {
    const size_t bytes = (argc > 1 ? atoll(argv[1]) : 48) << 20;     // default 48 MB: resident in L2
    float4 *B;
    float *sink;
    cudaMalloc(&B, bytes);
    cudaMalloc(&sink, 4);
    cudaMemset(B, 0, bytes);
    run<4>(B, bytes, sink);
    run<8>(B, bytes, sink);
    run<16>(B, bytes, sink);
    run<32>(B, bytes, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
