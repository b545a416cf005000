// Device-side integer packing / unpacking of the GPTQ layouts -- the shift-OR loops of
// QuantLinear.pack (quant/quant_linear.py:341-369 of the reference; its "TODO: perform packing
// on GPU", llama.py:264).  Bit-exact integer work; one thread owns one run of 32 values
// (= BITS output words), so no atomics are needed.
#include "common.cuh"
#include "kernels.h"

namespace gptq {
namespace {

// vals(i, o) = vals[i * vs_i + o * vs_o], i = index along the packed axis (length R), o = other axis (length C)
// words(wi, o) = packed[wi * ps_i + o * ps_o], wi in [0, R/32*BITS)
template <int BITS>
__global__ void pack_kernel(const int32_t* __restrict__ vals, int32_t* __restrict__ packed, int runs, int C, int64_t vs_i, int64_t vs_o, int64_t ps_i,
                            int64_t ps_o) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)runs * C) return;
    // consecutive threads walk the unit-stride axis
    int r, o;
    if (vs_o == 1) {
        o = int(t % C);
        r = int(t / C);
    } else {
        r = int(t % runs);
        o = int(t / runs);
    }
    uint32_t w[BITS];
#pragma unroll
    for (int i = 0; i < BITS; ++i) w[i] = 0u;
    constexpr uint32_t maxq = (1u << BITS) - 1u;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        // like the reference there is no clamp; we mask so that an off-grid value cannot corrupt neighbours
        const uint32_t v = (uint32_t)vals[(int64_t)(r * 32 + j) * vs_i + (int64_t)o * vs_o] & maxq;
        const int bit = BITS * j;
        const int wi = bit >> 5, sh = bit & 31;
        w[wi] |= v << sh;
        if constexpr (BITS == 3) {
            if (sh > 29) w[wi + 1] |= v >> (32 - sh);
        }
    }
#pragma unroll
    for (int i = 0; i < BITS; ++i) packed[(int64_t)(r * BITS + i) * ps_i + (int64_t)o * ps_o] = (int32_t)w[i];
}

template <int BITS>
__global__ void unpack_kernel(const int32_t* __restrict__ packed, int32_t* __restrict__ vals, int runs, int C, int64_t vs_i, int64_t vs_o, int64_t ps_i,
                              int64_t ps_o) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)runs * C) return;
    int r, o;
    if (vs_o == 1) {
        o = int(t % C);
        r = int(t / C);
    } else {
        r = int(t % runs);
        o = int(t / runs);
    }
    uint32_t w[BITS];
#pragma unroll
    for (int i = 0; i < BITS; ++i) w[i] = (uint32_t)packed[(int64_t)(r * BITS + i) * ps_i + (int64_t)o * ps_o];
#pragma unroll
    for (int j = 0; j < 32; ++j) vals[(int64_t)(r * 32 + j) * vs_i + (int64_t)o * vs_o] = extract_field<BITS>(w, j);
}

template <bool PACK>
cudaError_t launch(const int32_t* src, int32_t* dst, int R, int C, int bits, bool along_cols, cudaStream_t stream) {
    // along_cols == false: values [R, C] packed along rows  -> words [R/32*bits, C]   (qweight)
    // along_cols == true : values [C, R] packed along cols  -> words [C, R/32*bits]   (qzeros); R = packed-axis length
    const int runs = R / 32;
    const int64_t vs_i = along_cols ? 1 : C, vs_o = along_cols ? R : 1;
    const int64_t ps_i = along_cols ? 1 : C, ps_o = along_cols ? (int64_t)runs * bits : 1;
    const int64_t total = (int64_t)runs * C;
    if (total == 0) return cudaSuccess;
    const int threads = 256;
    const unsigned blocks = (unsigned)((total + threads - 1) / threads);
#define GPTQ_PACK_CASE(B)                                                                                                  \
    case B:                                                                                                                \
        if (PACK)                                                                                                          \
            pack_kernel<B><<<blocks, threads, 0, stream>>>(src, dst, runs, C, vs_i, vs_o, ps_i, ps_o);                     \
        else                                                                                                               \
            unpack_kernel<B><<<blocks, threads, 0, stream>>>(src, dst, runs, C, vs_i, vs_o, ps_i, ps_o);                   \
        break;
    switch (bits) {
        GPTQ_PACK_CASE(2)
        GPTQ_PACK_CASE(3)
        GPTQ_PACK_CASE(4)
        GPTQ_PACK_CASE(8)
        default: return cudaErrorInvalidValue;
    }
#undef GPTQ_PACK_CASE
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_pack_rows(const int32_t* vals, int32_t* packed, int R, int C, int bits, bool along_cols, cudaStream_t stream) {
    return launch<true>(vals, packed, R, C, bits, along_cols, stream);
}
cudaError_t launch_unpack_rows(const int32_t* packed, int32_t* vals, int R, int C, int bits, bool along_cols, cudaStream_t stream) {
    return launch<false>(packed, vals, R, C, bits, along_cols, stream);
}

}  // namespace gptq
