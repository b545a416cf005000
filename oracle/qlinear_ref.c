/* C restatement (OpenMP) of the reference's quantized-linear arithmetic -- TEST / BASELINE INFRASTRUCTURE ONLY.
 * Never linked into or called by the product (gptq-for-llama_b200/); used by tests/ as a second checker and by
 * bench.py's cpu_baseline / --impl reference leg as "the reference's arithmetic on the host cores".
 *
 * Follows matmul_248_kernel (quant/quant_linear.py:84-137 of the reference):
 *   W[k,n] = fp16( fp16(q[k,n] - (z[g_idx[k],n] + 1)) * s[g_idx[k],n] )     (:120-128, the +1 is unmasked)
 *   out[m,n] = fp16( sum_k fp32(x[m,k]) * fp32(W[k,n]) ) (+ bias in fp16, :376)     (:111,:130,:137)
 * and fusedmatmul_248_kernel (quant/fused_mlp.py:128-168): c = fp16( silu(acc1) * acc2 ), fp32.
 * 3-bit uses the 96-bit little-endian run layout (unpinned by the reference, see gptq_oracle.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef _Float16 h16;

static inline float h2f(uint16_t u) { h16 h; memcpy(&h, &u, 2); return (float)h; }
static inline uint16_t f2h(float f) { h16 h = (h16)f; uint16_t u; memcpy(&u, &h, 2); return u; }

static inline int field(const int32_t* run_base, long stride, int j, int bits) {
    /* value j (0..31) of the run whose words are run_base[0], run_base[stride], ... */
    int bit = bits * j, wi = bit >> 5, sh = bit & 31;
    uint32_t v = ((uint32_t)run_base[(long)wi * stride]) >> sh;
    if (sh + bits > 32) v |= ((uint32_t)run_base[(long)(wi + 1) * stride]) << (32 - sh);
    return (int)(v & ((1u << bits) - 1u));
}

/* acc[m*N + n] (fp32) = sum_k x[m,k] * W[k,n] */
static void accumulate(const uint16_t* x, const int32_t* qweight, const uint16_t* scales, const int32_t* qzeros, const int32_t* g_idx, float* acc, int M, int K,
                       int N, int bits) {
    const int zwords = N / 32 * bits;
#pragma omp parallel for schedule(static)
    for (int nb = 0; nb < N / 32; ++nb) { /* a thread owns 32 columns: one run of qzeros */
        float a[8][32];
        int mdone = 0;
        while (mdone < M) {
            const int mc = (M - mdone) < 8 ? (M - mdone) : 8;
            memset(a, 0, sizeof(a));
            int cur_g = -1;
            float zf[32], sf[32];
            for (int r = 0; r < K / 32; ++r) {
                for (int j = 0; j < 32; ++j) {
                    const int k = r * 32 + j;
                    const int g = g_idx[k];
                    if (g != cur_g) {
                        cur_g = g;
                        for (int c = 0; c < 32; ++c) {
                            zf[c] = (float)(field(qzeros + (long)g * zwords + nb * bits, 1, c, bits) + 1);
                            sf[c] = h2f(scales[(long)g * N + nb * 32 + c]);
                        }
                    }
                    float xv[8];
                    for (int m = 0; m < mc; ++m) xv[m] = h2f(x[(long)(mdone + m) * K + k]);
                    const int32_t* runp = qweight + (long)r * bits * N + nb * 32;
                    for (int c = 0; c < 32; ++c) {
                        const float wq = (float)field(runp + c, N, j, bits);
                        const float w = (float)(h16)((float)(h16)(wq - zf[c]) * sf[c]); /* one fp16 rounding of the exact product */
                        for (int m = 0; m < mc; ++m) a[m][c] += xv[m] * w;
                    }
                }
            }
            for (int m = 0; m < mc; ++m)
                for (int c = 0; c < 32; ++c) acc[(long)(mdone + m) * N + nb * 32 + c] = a[m][c];
            mdone += mc;
        }
    }
}

int gptq_ref_qlinear(const uint16_t* x, const int32_t* qweight, const uint16_t* scales, const int32_t* qzeros, const int32_t* g_idx, const uint16_t* bias,
                     uint16_t* out, int M, int K, int N, int bits) {
    if (!(bits == 2 || bits == 3 || bits == 4 || bits == 8) || K % 32 || N % 32) return -1;
    float* acc = (float*)malloc(sizeof(float) * (size_t)M * N);
    if (!acc) return -2;
    accumulate(x, qweight, scales, qzeros, g_idx, acc, M, K, N, bits);
    for (long i = 0; i < (long)M * N; ++i) {
        h16 o = (h16)acc[i];
        if (bias) o = (h16)((float)o + h2f(bias[i % N]));
        memcpy(&out[i], &o, 2);
    }
    free(acc);
    return 0;
}

int gptq_ref_fused_mlp(const uint16_t* x, const int32_t* qw1, const uint16_t* s1, const int32_t* qz1, const int32_t* g1, const int32_t* qw2, const uint16_t* s2,
                       const int32_t* qz2, const int32_t* g2, uint16_t* out, int M, int K, int N, int bits) {
    if (!(bits == 2 || bits == 3 || bits == 4 || bits == 8) || K % 32 || N % 32) return -1;
    float* a1 = (float*)malloc(sizeof(float) * (size_t)M * N);
    float* a2 = (float*)malloc(sizeof(float) * (size_t)M * N);
    if (!a1 || !a2) return -2;
    accumulate(x, qw1, s1, qz1, g1, a1, M, K, N, bits);
    accumulate(x, qw2, s2, qz2, g2, a2, M, K, N, bits);
    for (long i = 0; i < (long)M * N; ++i) {
        const float silu = a1[i] * (1.0f / (1.0f + expf(-a1[i])));
        out[i] = f2h(silu * a2[i]);
    }
    free(a1);
    free(a2);
    return 0;
}
