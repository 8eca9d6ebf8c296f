// Per-dtype arithmetic and 16-byte vector helpers shared by the CSR and COO kernels.
//
// The reference accumulates in val_dt (spmm_default/dpu_kernels/spmm_mul_csr_dpu.c:72-75,110-114),
// so integer results are defined modulo 2^bits.  We accumulate 8/16/32-bit integers in uint32 and
// 64-bit integers in uint64 and truncate on store: (a + b*c) mod 2^n is a ring homomorphism, so
// wide-accumulate-then-truncate equals the reference's narrow wraparound bit for bit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

namespace pygim {

template <typename T> struct Arith;
template <> struct Arith<int8_t>  { using Acc = uint32_t; using Shfl = int; };
template <> struct Arith<int16_t> { using Acc = uint32_t; using Shfl = int; };
template <> struct Arith<int32_t> { using Acc = uint32_t; using Shfl = int; };
template <> struct Arith<int64_t> { using Acc = unsigned long long; using Shfl = long long; };
template <> struct Arith<float>   { using Acc = float; using Shfl = float; };
template <> struct Arith<double>  { using Acc = double; using Shfl = double; };

// E elements of T moved as one naturally aligned word of sizeof(T)*E bytes (1..16).
template <typename T, int E> struct alignas(sizeof(T) * E) Pack { T e[E]; };

template <int BYTES> struct Word;
template <> struct Word<1>  { using type = unsigned char; };
template <> struct Word<2>  { using type = unsigned short; };
template <> struct Word<4>  { using type = unsigned int; };
template <> struct Word<8>  { using type = uint2; };
template <> struct Word<16> { using type = uint4; };

// read-only (non-coherent) load: the dense feature rows; default L1/L2 policy so reuse is kept
template <typename T, int E> __device__ __forceinline__ Pack<T, E> ld_dense(const T *p) {
    using W = typename Word<sizeof(T) * E>::type;
    union { W w; Pack<T, E> v; } u;
    u.w = __ldg(reinterpret_cast<const W *>(p));
    return u.v;
}
// Address of this lane's word in dense row `col`: base + col * row_stride_bytes as ONE 32x32->64-bit multiply-add
// (IMAD.WIDE.U32).  A 64-bit stride costs seven integer instructions per gather in the inner loop.
template <typename T> __device__ __forceinline__ const T *row_ptr(const T *base, int col, unsigned stride_bytes) {
    return reinterpret_cast<const T *>(reinterpret_cast<const char *>(base) +
                                       (unsigned long long)(unsigned)col * stride_bytes);
}
// plain load (read-modify-write of C in accumulate mode)
template <typename T, int E> __device__ __forceinline__ Pack<T, E> ld_plain(const T *p) {
    using W = typename Word<sizeof(T) * E>::type;
    union { W w; Pack<T, E> v; } u;
    u.w = *reinterpret_cast<const W *>(p);
    return u.v;
}
// L1-bypassing load (ld.global.cg): data another SM wrote during this launch (segment partial sums)
template <typename T, int E> __device__ __forceinline__ Pack<T, E> ld_cg(const T *p) {
    using W = typename Word<sizeof(T) * E>::type;
    union { W w; Pack<T, E> v; } u;
    u.w = __ldcg(reinterpret_cast<const W *>(p));
    return u.v;
}
// streaming store (evict-first): C rows are written once and not re-read by this kernel
template <typename T, int E> __device__ __forceinline__ void st_stream(T *p, const Pack<T, E> &v) {
    using W = typename Word<sizeof(T) * E>::type;
    union { W w; Pack<T, E> v; } u;
    u.v = v;
    __stcs(reinterpret_cast<W *>(p), u.w);
}
template <typename T, int E> __device__ __forceinline__ void st_plain(T *p, const Pack<T, E> &v) {
    using W = typename Word<sizeof(T) * E>::type;
    union { W w; Pack<T, E> v; } u;
    u.v = v;
    *reinterpret_cast<W *>(p) = u.w;
}

// streaming loads of the sparse index/value streams (read exactly once: evict-first)
__device__ __forceinline__ int ld_stream(const int *p) { return __ldcs(p); }
__device__ __forceinline__ float ld_stream(const float *p) { return __ldcs(p); }
__device__ __forceinline__ double ld_stream(const double *p) { return __ldcs(p); }
__device__ __forceinline__ long long ld_stream(const int64_t *p) {
    return __ldcs(reinterpret_cast<const long long *>(p));
}
__device__ __forceinline__ int ld_stream(const int16_t *p) { return (int)__ldcs(reinterpret_cast<const short *>(p)); }
__device__ __forceinline__ int ld_stream(const int8_t *p) {
    return (int)(signed char)__ldcs(reinterpret_cast<const signed char *>(p));
}

// acc[k] += v * b[k] in the accumulator type
template <typename T, int E>
__device__ __forceinline__ void fma_pack(typename Arith<T>::Acc (&acc)[E], const Pack<T, E> &b,
                                         typename Arith<T>::Shfl v) {
    using Acc = typename Arith<T>::Acc;
#pragma unroll
    for (int k = 0; k < E; ++k) {
        if constexpr (std::is_integral<T>::value) {
            // sign-extend to the accumulator width, multiply-add modulo 2^32 / 2^64
            acc[k] += (Acc)(typename Arith<T>::Shfl)b.e[k] * (Acc)v;
        } else {
            acc[k] = fma((Acc)b.e[k], (Acc)v, acc[k]);
        }
    }
}

// INT8, 16 elements per 16-byte word: one IDP.4A per element and no unpacking.  The multiplier byte is
// placed in lane j of the second operand (zeros elsewhere), so dp4a(word, v << 8j, acc) = acc + word.byte[j] * v
// with both bytes taken as signed - exactly the sign-extended multiply-add, modulo 2^32.
template <>
__device__ __forceinline__ void fma_pack<int8_t, 16>(uint32_t (&acc)[16], const Pack<int8_t, 16> &b, int v) {
    union { Pack<int8_t, 16> p; int w[4]; } u;
    u.p = b;
    const int vb = v & 0xff;
    const int sel[4] = {vb, vb << 8, vb << 16, vb << 24};
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = (uint32_t)__dp4a(u.w[k >> 2], sel[k & 3], (int)acc[k]);
}

template <typename T, int E>
__device__ __forceinline__ Pack<T, E> narrow(const typename Arith<T>::Acc (&acc)[E]) {
    Pack<T, E> r;
#pragma unroll
    for (int k = 0; k < E; ++k) r.e[k] = (T)acc[k];
    return r;
}

// r[k] = old[k] + acc[k] (accumulate mode: sparse parts >= 1 add into C, spmm_mul_csr.c:497-502)
template <typename T, int E>
__device__ __forceinline__ void add_old(typename Arith<T>::Acc (&acc)[E], const Pack<T, E> &old) {
    using Acc = typename Arith<T>::Acc;
#pragma unroll
    for (int k = 0; k < E; ++k) {
        if constexpr (std::is_integral<T>::value) acc[k] += (Acc)(typename Arith<T>::Shfl)old.e[k];
        else acc[k] = (Acc)old.e[k] + acc[k];
    }
}

}  // namespace pygim
