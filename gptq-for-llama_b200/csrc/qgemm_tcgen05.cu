// Batched (prefill) quantized GEMM on the 5th-generation tensor cores: out[M,N] = x[M,K] . deq(W), int4, M > 8.
//
// This path IS a dense contraction (2*M flop per 0.53 B of weight: compute-bound above M ~ 70), so unlike the decode
// matvec it belongs on tcgen05.  Replaces matmul_248_kernel for large M (quant/quant_linear.py:72-137 of the reference;
// there: mma.sync tiles chosen by an autotuner) with one static configuration:
//
//   CTA tile 128 (M) x 128 (N), K step 64, 3-stage shared-memory ring, fp32 accumulator in TMEM (128 lanes x 128 columns)
//   A (activations)  : cp.async 16 B chunks into the canonical K-major SWIZZLE_128B layout (row r at r*128 B, chunk c at c ^ (r & 7))
//   B (weights)      : each thread dequantises packed words -- one int32 = 8 consecutive k of one column = exactly one 16 B
//                      chunk of the K-major B tile -- with the reference-exact fp16 arithmetic (int4_core.cuh) and stores it
//                      swizzled; the dequantised tile never touches HBM
//   MMA              : one elected thread issues 4 x tcgen05.mma.cta_group::1.kind::f16 (M128 N128 K16) per K step from
//                      shared-memory descriptors, tcgen05.commit -> mbarrier releases the stage
//   epilogue         : tcgen05.ld 32x32b -> fp16 (+ bias) -> global
//
// Numerics: fp16 operands identical to the reference's (exact dequant), fp32 accumulation in TMEM, one fp16 rounding.
#include <cuda.h>
#include <cudaTypedefs.h>

#include "common.cuh"
#include "int4_core.cuh"
#include "kernels.h"

namespace gptq {
namespace {

using namespace int4;

constexpr int BM = 128, BN = 128, BK = 64;
constexpr int kAhead = 2;     // operands are requested kAhead K-steps before they are consumed (hides ~1.5 us of L2/DRAM latency)
constexpr int kTileBytes = BM * BK * 2;  // 16 KB (A and B tiles have the same size)
constexpr int kGemmThreads = 256;              // producer threads: A staging + B dequantisation, later the epilogue
constexpr int kGemmBlock = kGemmThreads + 64;  // + one warp issuing the tcgen05.mma stream + one warp issuing the TMA loads of A

// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (bit 4), A = B = F16 (0), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor), K-major, SWIZZLE_128B: LBO = 1 (16 B), SBO = 1024 B between
// 8-row groups, version = 1 (Blackwell), layout type = 2
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
          "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
          "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct GemmParams {
    const __half* x;
    int64_t ldx;
    const uint32_t* qw[2];  // [1]: second weight of the fused SwiGLU MLP (DUAL)
    const __half* sc[2];
    const uint32_t* qz[2];
    const __half* bias;
    __half* out;
    int64_t ldo;
    int M, K, N, groupsize;
};

// DUAL: out = silu(x.Wg) * (x.Wu) with both fp32 accumulators in TMEM (fusedmatmul_248_kernel, quant/fused_mlp.py:84-168)
// MT: M tiles of 128 rows per CTA (1 or 2).  With MT = 2 every dequantised B tile feeds two MMAs (256 rows), which halves the
// CUDA-core dequant work per tensor-core flop -- the limiter of the MT = 1 configuration.
template <bool DUAL, int MT, int kStagesG>
__global__ void __launch_bounds__(kGemmBlock, 1) qgemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const GemmParams p) {
    constexpr int NW = DUAL ? 2 : 1;
    constexpr uint32_t kTmemCols = 128 * NW * MT;  // accumulator (weight w, M tile mt) lives at column 128 * (w * MT + mt)
    constexpr int kATile = MT * kTileBytes;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ __align__(8) unsigned long long bars[2 * kStagesG + 1];  // empty[S] (MMA -> producers), acc_done, full[S] (producers -> MMA)
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.y * (BM * MT), n0 = blockIdx.x * BN;
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1024 B alignment
    uint8_t* sptr = smem_raw + (sbase - smem_u32(smem_raw));
    const uint32_t a_s = sbase, b_s = sbase + kStagesG * kATile;  // A tiles: [stage][MT x 128 rows]; B tiles: [stage][weight]
    uint8_t* b_ptr = sptr + kStagesG * kATile;
    const uint32_t bar0 = smem_u32(&bars[0]);

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int i = 0; i <= kStagesG; ++i) mbar_init(bar0 + i * 8, 1);
        for (int i = 0; i < kStagesG; ++i) mbar_init(bar0 + (kStagesG + 1 + i) * 8, kGemmThreads + 1);  // 256 B-producers + the TMA thread's expect_tx
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;

    const uint32_t full0 = bar0 + (kStagesG + 1) * 8;
    const int nkb = p.K / BK;
    if (warp == kGemmThreads / 32) {
        // ===== MMA issuer: one elected lane streams tcgen05.mma; tcgen05.commit releases each stage and finally the accumulator =====
        if (lane == 0) {
#pragma unroll 1
            for (int it = 0; it < nkb; ++it) {
                const int s = it % kStagesG;
                mbar_wait(full0 + s * 8, (it / kStagesG) & 1u);  // all 256 producers have filled this stage
                tc_fence_after();
#pragma unroll
                for (int w = 0; w < NW; ++w) {
                    const uint64_t bd = smem_desc(b_s + (s * NW + w) * kTileBytes);
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        const uint64_t ad = smem_desc(a_s + s * kATile + mt * kTileBytes);
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)  // +32 B per K=16 inside the swizzle atom
                            tc_mma(tmem + 128 * (w * MT + mt), ad + 2 * k, bd + 2 * k, kIdesc, (it > 0 || k > 0) ? 1u : 0u);
                    }
                }
                tc_commit(bar0 + s * 8);
                if (it + 1 == nkb) tc_commit(bar0 + kStagesG * 8);  // accumulator complete
            }
        }
        tc_fence_before();
        __syncthreads();  // matches the producers' barrier before the TMEM dealloc
        return;
    }

    if (warp == kGemmThreads / 32 + 1) {
        // ===== TMA producer for A: box 64 (k) x 128 (rows), SWIZZLE_128B = exactly the canonical K-major layout; rows beyond M are zero-filled =====
        if (lane == 0) {
#pragma unroll 1
            for (int it = 0; it < nkb; ++it) {
                const int s = it % kStagesG;
                if (it >= kStagesG) mbar_wait(bar0 + s * 8, ((it / kStagesG) - 1) & 1u);  // MMAs that read this stage are done
                mbar_expect_tx(full0 + s * 8, MT * kTileBytes);
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                                     a_s + s * kATile + mt * kTileBytes),
                                 "l"(&tmA), "r"(it * BK), "r"(m0 + 128 * mt), "r"(full0 + s * 8)
                                 : "memory");
            }
        }
        tc_fence_before();
        __syncthreads();
        return;
    }

    // ---- per-thread roles ---------------------------------------------------------------------------------------
    // A: 1024 chunks of 16 B per stage, 4 per thread: chunk id = tid + 256 i -> row id >> 3, k-chunk id & 7
    // B: column n = tid & 127, k-chunks c = (tid >> 7) + 2 i
    const int bn = tid & 127, bc0 = tid >> 7;
    const int col = n0 + bn;
    const int zshift = (col & 7) * 4;

    uint32_t bq[kAhead + 1][NW][4];  // packed words of K steps it .. it + kAhead (a register ring, rotated every step)
    auto load_b = [&](int it, uint32_t (&dst)[NW][4]) {
        const int kr0 = it * (BK / 8);
#pragma unroll
        for (int w = 0; w < NW; ++w)
#pragma unroll
            for (int i = 0; i < 4; ++i) dst[w][i] = __ldg(p.qw[w] + col + (size_t)(kr0 + bc0 + 2 * i) * p.N);
    };

    int cur_grp = -1;
    __half2 za[NW], zb[NW], sc2[NW];
#pragma unroll
    for (int d = 0; d < kAhead; ++d) {
        if (d < nkb) load_b(d, bq[d]);
    }
#pragma unroll 1
    for (int it = 0; it < nkb; ++it) {
        const int s = it % kStagesG;
        // ---- first: request the packed words of step it + kAhead (A arrives by TMA from the producer warp) -------
        if (it + kAhead < nkb) load_b(it + kAhead, bq[kAhead]);
        if (it >= kStagesG) mbar_wait(bar0 + s * 8, ((it / kStagesG) - 1) & 1u);  // the MMAs that read this stage have completed
        // ---- dequantise this step's packed words into the B tile --------------------------------------------------
        const int grp = (it * BK) / p.groupsize;
        if (grp != cur_grp) {
            cur_grp = grp;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const __half sv = __ldg(p.sc[w] + (size_t)grp * p.N + col);
                const uint32_t zw = __ldg(p.qz[w] + (size_t)grp * (p.N >> 3) + (col >> 3));
                const float z = (float)(((zw >> zshift) & 0xfu) + 1u);  // stored minus one, +1 unmasked (quant_linear.py:120-121)
                za[w] = __float2half2_rn(1024.f + z);
                zb[w] = __float2half2_rn(-(64.f + z));
                sc2[w] = __half2half2(sv);
            }
        }
#pragma unroll
        for (int w = 0; w < NW; ++w) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                uint32_t v[4];  // (k0,k4) (k1,k5) (k2,k6) (k3,k7)
                dequant8<0>(bq[0][w][i], za[w], zb[w], sc2[w], v);
                uint4 o;
                o.x = __byte_perm(v[0], v[1], 0x5410);  // (k0,k1)
                o.y = __byte_perm(v[2], v[3], 0x5410);  // (k2,k3)
                o.z = __byte_perm(v[0], v[1], 0x7632);  // (k4,k5)
                o.w = __byte_perm(v[2], v[3], 0x7632);  // (k6,k7)
                const int c = bc0 + 2 * i;
                *reinterpret_cast<uint4*>(b_ptr + (s * NW + w) * kTileBytes + bn * 128 + ((c ^ (bn & 7)) << 4)) = o;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core (async proxy)
        mbar_arrive(full0 + s * 8);  // no block-wide barrier in the loop: the MMA warp waits for 256 arrivals
#pragma unroll
        for (int d = 0; d < kAhead; ++d)
#pragma unroll
            for (int w = 0; w < NW; ++w)
#pragma unroll
                for (int i = 0; i < 4; ++i) bq[d][w][i] = bq[d + 1][w][i];
    }

    // ---- epilogue: TMEM -> registers -> fp16 (+bias) -> global -------------------------------------------------------
    mbar_wait(bar0 + kStagesG * 8, 0);
    tc_fence_after();
    {
        const int q = warp & 3, half = warp >> 2;  // a warp may only touch TMEM lanes [32 (warp % 4), +32)
#pragma unroll 1
        for (int jj = 0; jj < 2 * MT; ++jj) {
            const int mt = jj >> 1, j = jj & 1;
            const int row = m0 + 128 * mt + 32 * q + lane;
            const int c0 = 64 * half + 32 * j;
            uint32_t r[32];
            tc_ld32(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(128 * mt + c0), r);
            if constexpr (DUAL) {  // silu(gate) * up on the fp32 accumulators, one rounding (quant/fused_mlp.py:163-165)
                uint32_t r2[32];
                tc_ld32(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(128 * (MT + mt) + c0), r2);
#pragma unroll
                for (int e = 0; e < 32; ++e) r[e] = __float_as_uint(swiglu(__uint_as_float(r[e]), __uint_as_float(r2[e])));
            }
            if (row < p.M) {
                __half* orow = p.out + (size_t)row * p.ldo + n0 + c0;
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    uint32_t pk[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        __half h0 = __float2half_rn(__uint_as_float(r[v * 8 + 2 * e])), h1 = __float2half_rn(__uint_as_float(r[v * 8 + 2 * e + 1]));
                        if (p.bias != nullptr) {
                            h0 = __hadd(h0, p.bias[n0 + c0 + v * 8 + 2 * e]);
                            h1 = __hadd(h1, p.bias[n0 + c0 + v * 8 + 2 * e + 1]);
                        }
                        pk[e] = h2_as_u32(__halves2half2(h0, h1));
                    }
                    *reinterpret_cast<uint4*>(orow + v * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
}

inline bool al(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

}  // namespace

bool gemm_tc_supported(const QLinearArgs& a) {
    const gptq_qweight& w = a.w;
    if (w.bits != 4 || a.M <= 8) return false;
    if (a.dual && (!al(a.w2.qweight, 4))) return false;
    if (w.groupsize <= 0 || w.groupsize % BK != 0) return false;
    if (w.N % BN != 0 || w.K % BK != 0) return false;
    if (!al(a.x, 16) || a.ldx % 8 != 0 || !al(a.out, 16) || a.ldo % 8 != 0) return false;
    if (a.norm_w != nullptr || a.residual != nullptr) return false;
    return true;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (libcuda is not linked: the library must load without a driver)
static bool make_x_tensor_map(CUtensorMap* tm, const void* x, int M, int K, int64_t ldx) {
    static PFN_cuTensorMapEncodeTiled encode = []() -> PFN_cuTensorMapEncodeTiled {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
        return reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
    }();
    if (encode == nullptr) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M};   // innermost first
    const cuuint64_t strides[1] = {(cuuint64_t)ldx * 2};         // bytes between rows
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};  // 64 halves (128 B) x 128 rows
    const cuuint32_t estr[2] = {1, 1};
    return encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

cudaError_t launch_qlinear_gemm_tc(const QLinearArgs& a) {
    CUtensorMap tmA;
    if (!make_x_tensor_map(&tmA, a.x, a.M, a.w.K, a.ldx)) return cudaErrorNotSupported;
    GemmParams p{};
    p.x = reinterpret_cast<const __half*>(a.x);
    p.ldx = a.ldx;
    p.qw[0] = reinterpret_cast<const uint32_t*>(a.w.qweight);
    p.sc[0] = reinterpret_cast<const __half*>(a.w.scales);
    p.qz[0] = reinterpret_cast<const uint32_t*>(a.w.qzeros);
    if (a.dual) {
        p.qw[1] = reinterpret_cast<const uint32_t*>(a.w2.qweight);
        p.sc[1] = reinterpret_cast<const __half*>(a.w2.scales);
        p.qz[1] = reinterpret_cast<const uint32_t*>(a.w2.qzeros);
    }
    p.bias = reinterpret_cast<const __half*>(a.bias);
    p.out = reinterpret_cast<__half*>(a.out);
    p.ldo = a.ldo;
    p.M = a.M; p.K = a.w.K; p.N = a.w.N; p.groupsize = a.w.groupsize;
    const int mt = a.M > BM ? 2 : 1;
    const int stages = (a.dual && mt == 2) ? 3 : 4;  // 256-row dual tiles: 64 KB per stage -> 3 stages; everything else 4
    const size_t smem = 1024 + (size_t)(mt + (a.dual ? 2 : 1)) * stages * kTileBytes;
    const dim3 grid(a.w.N / BN, ceil_div(a.M, BM * mt));
    auto go = [&](auto kernel) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kernel<<<grid, kGemmBlock, smem, a.stream>>>(tmA, p);
        return cudaGetLastError();
    };
    if (a.dual) return mt == 2 ? go(qgemm_tcgen05_kernel<true, 2, 3>) : go(qgemm_tcgen05_kernel<true, 1, 4>);
    return mt == 2 ? go(qgemm_tcgen05_kernel<false, 2, 4>) : go(qgemm_tcgen05_kernel<false, 1, 4>);
}

}  // namespace gptq
