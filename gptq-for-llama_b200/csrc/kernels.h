// Internal launch interface between capi.cu (validation + dispatch) and the kernel files.
#pragma once
#include <cuda_runtime.h>

#include "gptq_b200.h"

namespace gptq {

struct QLinearArgs {
    const void* x;
    int64_t ldx;
    gptq_qweight w;   // gate for the fused MLP
    gptq_qweight w2;  // up for the fused MLP (unused otherwise)
    bool dual;        // fused SwiGLU MLP
    const void* bias;      // fp16 [N] or null
    const void* residual;  // fp16 [M, ldr] or null: out = residual + fp16(acc)   (decode engine)
    int64_t ldr;
    const void* norm_w;    // fp16 [K] or null: RMSNorm(x; norm_w, eps) fused in front of the product
    float eps;
    void* out;
    int64_t ldo;
    int M;
    void* workspace;
    size_t ws_bytes;
    cudaStream_t stream;
};

// generic.cu -- CUDA-core kernels, any bits in {2,3,4,8}, any g_idx, any M.
cudaError_t launch_qlinear_generic(const QLinearArgs& a);
cudaError_t launch_qlinear_transpose_generic(const void* g, int64_t ldg, const gptq_qweight& w, void* out, int64_t ldo, int M, cudaStream_t stream);
cudaError_t launch_dequant(const gptq_qweight& w, void* out, int64_t ldo, cudaStream_t stream);

// qmatvec.cu -- tuned int4 decode kernel (M <= 8, no act-order): stream-K over 2 CTAs/SM.
struct SkinnyPlan {
    int nslabs, nk, grid, max_contrib, seg_steps;
    long long total_units;
};
SkinnyPlan plan_skinny(int M, int K, int N);
size_t skinny_workspace_bytes(int M, int K, int N, bool dual);
bool skinny_supported(const QLinearArgs& a);
cudaError_t launch_qlinear_skinny(const QLinearArgs& a, bool pdl);

// qgemm_tcgen05.cu -- batched (prefill) int4 GEMM on tcgen05 tensor cores, M > 8
bool gemm_tc_supported(const QLinearArgs& a);
cudaError_t launch_qlinear_gemm_tc(const QLinearArgs& a);

// decode_mega.cu -- persistent single-kernel decode step (batch 1, int4 kernel-form layers)
bool mega_supported(const gptq_llama_model& m, const gptq_llama_state& st);
size_t mega_scratch_bytes(const gptq_llama_model& m, int max_seq);
cudaError_t launch_decode_mega(const gptq_llama_model& m, const gptq_llama_state& st, uint8_t* scratch, cudaStream_t stream);

// elementwise.cu
cudaError_t launch_rope(void* qk, int64_t token_stride, const int64_t* position_ids, int64_t pos_batch_stride, int bsz, int seq, int rows, int head_dim,
                        float base, cudaStream_t stream);
cudaError_t launch_rmsnorm(const void* x, int64_t ldx, const void* weight, void* y, int64_t ldy, int M, int N, float eps, cudaStream_t stream);

// pack.cu
cudaError_t launch_pack_rows(const int32_t* vals, int32_t* packed, int R, int C, int bits, bool along_cols, cudaStream_t stream);
cudaError_t launch_unpack_rows(const int32_t* packed, int32_t* vals, int R, int C, int bits, bool along_cols, cudaStream_t stream);

}  // namespace gptq
