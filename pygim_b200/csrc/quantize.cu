// symmetric_quantize (models/quantize.py:20-38) on the device: an absmax reduction and one fused
// divide-round-convert pass - the producer side of the quantised aggregation
// (x_q feeds adj_t.mul, pyg_gcn_conv.py:130-137).  torch runs this as abs, max, mul, div, div, round, to:
// seven launches and five N*H-sized temporaries.
#include "../../include/pygim_b200.h"

#include <cuda_runtime.h>
#include <stdint.h>

namespace pygim {

// |x| >= 0, so the float order equals the order of the bit patterns taken as unsigned ints
__global__ void absmax_kernel(const float *x, long long rows, long long cols, long long ldx, unsigned int *out_bits) {
    const long long n = rows * cols;
    float m = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / cols, c = i - r * cols;
        const float v = fabsf(x[r * ldx + c]);
        m = (v > m || v != v) ? v : m;            // NaN propagates like torch.max
    }
    for (int off = 16; off > 0; off >>= 1) {
        const float o = __shfl_xor_sync(0xffffffffu, m, off);
        m = (o > m || o != o) ? o : m;
    }
    if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_uint(m));
}

// scale = absmax * 2 / 2^k: a multiplication by a power of two, exact in float32 like torch's two steps
__global__ void make_scale_kernel(const unsigned int *absmax_bits, float *scale, float pow2) {
    *scale = __fdiv_rn(__fmul_rn(__uint_as_float(*absmax_bits), 2.0f), pow2);
}

template <typename Q>
__global__ void quantize_kernel(const float *x, long long rows, long long cols, long long ldx, Q *xq, long long ldq,
                                const float *scale) {
    const float s = *scale;
    const long long n = rows * cols;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / cols, c = i - r * cols;
        const float q = rintf(__fdiv_rn(x[r * ldx + c], s));       // torch.round = round half to even
        xq[r * ldq + c] = (Q)q;
    }
}

}  // namespace pygim

int pygim_fail_invalid(const char *msg);   // backend_pim.cu: records the message for pygim_last_error()

using namespace pygim;

extern "C" PYGIM_API int pygim_quantize(const float *x, int64_t rows, int64_t cols, int64_t ldx, int dtype, void *xq,
                                        int64_t ldq, float *scale, void *stream) {
    if (!x || !xq || !scale || rows < 0 || cols < 0) return pygim_fail_invalid("pygim_quantize: null buffer or negative size");
    float pow2;
    switch (dtype) {
        case PYGIM_INT8: pow2 = 32.f; break;            // 2^5
        case PYGIM_INT16: pow2 = 1024.f; break;         // 2^10
        case PYGIM_INT32: case PYGIM_FLT32: pow2 = 1048576.f; break;   // 2^20
        default: return pygim_fail_invalid("pygim_quantize: INT8 / INT16 / INT32 / FLT32 only");
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // the scale slot doubles as the absmax accumulator (bit pattern), then becomes the scale
    unsigned int *bits = reinterpret_cast<unsigned int *>(scale);
    if (cudaMemsetAsync(bits, 0, sizeof(unsigned int), st) != cudaSuccess) return pygim_fail_invalid("pygim_quantize: memset failed");
    const long long n = (long long)rows * cols;
    if (n > 0) {
        const int blocks = (int)((n + 1023) / 1024 < 1184 ? (n + 1023) / 1024 : 1184);
        absmax_kernel<<<blocks, 256, 0, st>>>(x, rows, cols, ldx, bits);
        make_scale_kernel<<<1, 1, 0, st>>>(bits, scale, pow2);
        switch (dtype) {
            case PYGIM_INT8: quantize_kernel<int8_t><<<blocks, 256, 0, st>>>(x, rows, cols, ldx, static_cast<int8_t *>(xq), ldq, scale); break;
            case PYGIM_INT16: quantize_kernel<int16_t><<<blocks, 256, 0, st>>>(x, rows, cols, ldx, static_cast<int16_t *>(xq), ldq, scale); break;
            case PYGIM_INT32: quantize_kernel<int32_t><<<blocks, 256, 0, st>>>(x, rows, cols, ldx, static_cast<int32_t *>(xq), ldq, scale); break;
            default: quantize_kernel<float><<<blocks, 256, 0, st>>>(x, rows, cols, ldx, static_cast<float *>(xq), ldq, scale); break;
        }
    } else {
        make_scale_kernel<<<1, 1, 0, st>>>(bits, scale, pow2);
    }
    if (cudaGetLastError() != cudaSuccess) return pygim_fail_invalid("pygim_quantize: launch failed");
    return PYGIM_OK;
}
