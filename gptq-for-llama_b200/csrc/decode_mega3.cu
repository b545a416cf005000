// Persistent decode step, second layout: ONE CTA per SM made of three 8-warp consumer groups and one producer warp.
//
// decode_mega.cu (two 10-warp CTAs per SM) runs its matvec loop at ~600 cycles per 4 KB tile per CTA although the
// pipes would allow ~400 and HBM ~300: with four consumer warps per scheduler the loop is bound by per-warp
// dependency latency (profiles/r1_mega_timeline.txt).  This layout trades the dual-accumulator gate|up loop (96
// registers) for a single-matrix loop (<= 80 registers) and spends the registers on warps:
//   * 24 consumer warps per SM (6 per scheduler instead of 4): three groups, each with its own weight ring and its own
//     contiguous unit range -- 444 ranges instead of 296;
//   * gate and up are one operation over a virtual [K, 2I] matrix (a unit belongs to one of them), so every matvec runs
//     the same loop;
//   * one producer warp, one quad of lanes per group, lane r copying packed row r of every tile (1 KB bulk copies);
//   * the staging passes (RMSNorm, residual add) run once per SM with 768 threads instead of once per CTA twice per SM,
//     and the grid barrier has 148 participants instead of 296.
// Operation list, arithmetic and scratch layout are those of decode_mega.cu (see there); the reference rounding points
// are identical, only the fp32 summation order of the split-K partials differs.
#include <cstdlib>

#include "common.cuh"
#include "int4_core.cuh"
#include "kernels.h"

namespace gptq {
namespace {

using namespace int4;

constexpr int kGroups = 3;
constexpr int kGWarps = 8;
constexpr int kGThreads = 32 * kGWarps;         // 256: one group covers a 256-column slab
constexpr int kConsumers = kGroups * kGThreads;  // 768
constexpr int kCWarps = kGroups * kGWarps;       // 24
constexpr int kBlock = kConsumers + 32;          // + the producer warp
constexpr int kSlabCols = 256;
constexpr int kRowPitch = 1024 + 32;  // smem pitch of a 1 KB weight row (+32 B: the 4 rows of a k-step hit distinct bank groups)
constexpr int kTile = 4 * kRowPitch;  // one ring stage: 4 packed rows x 256 columns (4 KB of weights)
constexpr int kStages = 12;           // per group: 3 x 12 x 4 KB = 144 KB of weights in flight per SM
constexpr int kHD = 128;
constexpr int kAttnChunk = 256;  // keys per attention work item (same split as decode_mega.cu: the scratch layout is shared)
constexpr int kAttnIter = 2;     // keys per thread group and pass
constexpr int kAttnPass = 32 * kAttnIter;
constexpr int kRec = kHD + 4;  // floats per split-KV partial record: m, l, 2 pad, o[128]
constexpr int kMaxLayers = 80;
constexpr int kSegHalves = 2048;                     // per-group slice of xs for segment-staged inputs (o_proj, down_proj)
constexpr int kAttnScratchFloats = 192 + 32 * 132;  // per group: q[128], m[32], l[32], o[32][132]

#ifdef GPTQ_TRACE
}  // namespace
__device__ unsigned long long* g_mega3_trace = nullptr;
namespace {
#define MTRACE(id)                                                                                 \
    do {                                                                                           \
        if (g_mega3_trace != nullptr && threadIdx.x == 0 && (id) < 64) {                            \
            unsigned long long t_;                                                                 \
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_));                                  \
            g_mega3_trace[blockIdx.x * 64 + (id)] = t_;                                            \
        }                                                                                          \
    } while (0)
#else
#define MTRACE(id) \
    do {           \
    } while (0)
#endif

struct MatDesc {
    const uint32_t* qw;
    const __half* sc;
    const uint32_t* qz;
};
struct LayerDesc {
    MatDesc qkv, o, gate, up, down;
    const __half* input_norm;
    const __half* post_norm;
    const int32_t* qkv_perm;  // act-order input gathers (gptq_llama_layer), nullptr = identity
    const int32_t* o_perm;
    const int32_t* mlp_perm;
};
struct Mega3Params {
    int n_layers, H, I, V, n_heads, groupsize, max_seq, nsplit;
    float eps, inv_base, scale;
    const __half* embed;
    const __half* final_norm;
    const __half* lm_head;
    const int32_t* tokens;
    const int32_t* positions;
    __half* k_cache;
    __half* v_cache;
    size_t layer_stride;  // halves per layer in the caches
    __half* logits;
    int32_t* next_token;
    __half* resid[2];
    float* acc_qkv;
    float* acc_o;
    float* acc_g;
    float* acc_u;
    float* acc_d;
    float* part;
    float* rope_cs;
    unsigned long long* bar;
    int xs_halves;  // halves reserved for xs in shared memory (>= H and >= kGroups * kSegHalves)
    LayerDesc layers[kMaxLayers];
};

// One matvec operation as its work units see it: `nmat` matrices of N columns each (gate|up: 2), unit u =
// (virtual slab, k-step) numbered slab-major; virtual slab v belongs to matrix v / slabs, slab v % slabs.
struct MatView {
    MatDesc w[2];
    float* out[2];
    int K, N, slabs, nmat;
};

__device__ __forceinline__ MatView mat_view(const Mega3Params& p, int idx) {
    const LayerDesc& L = p.layers[idx >> 2];
    MatView v;
    v.nmat = 1;
    v.out[1] = nullptr;
    switch (idx & 3) {
        case 0: v.w[0] = v.w[1] = L.qkv; v.K = p.H; v.N = 3 * p.H; v.out[0] = p.acc_qkv; break;
        case 1: v.w[0] = v.w[1] = L.o; v.K = p.H; v.N = p.H; v.out[0] = p.acc_o; break;
        case 2: v.w[0] = L.gate; v.w[1] = L.up; v.K = p.H; v.N = p.I; v.out[0] = p.acc_g; v.out[1] = p.acc_u; v.nmat = 2; break;
        default: v.w[0] = v.w[1] = L.down; v.K = p.I; v.N = p.H; v.out[0] = p.acc_d; break;
    }
    v.slabs = v.N / kSlabCols;
    return v;
}

// this group's contiguous unit range of an operation with U units
__device__ __forceinline__ void unit_range(unsigned U, int grp, int& u0, int& u1) {
    const unsigned nbg = gridDim.x * kGroups, gi = blockIdx.x * kGroups + grp;
    u0 = (int)(((unsigned long long)gi * U) / nbg);
    u1 = (int)(((unsigned long long)(gi + 1) * U) / nbg);
}

// ---- pipeline state of one consumer thread (its group's ring) ----
struct Pipe {
    uint32_t tile;    // smem address of this lane's 16 B in the current stage
    uint32_t bar;     // smem address of full[current stage]; empty[s] sits kStages * 8 bytes after full[s]
    uint32_t parity;  // expected parity of the current round
    int left;         // stages until the ring wraps
};

__device__ __forceinline__ void cta_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kConsumers) : "memory"); }  // all consumer warps
__device__ __forceinline__ void grp_sync(int grp) { asm volatile("bar.sync %0, %1;" ::"r"(2 + grp), "n"(kGThreads) : "memory"); }

// Producer: lanes 4g..4g+3 of the producer warp feed group g; lane r copies packed row r of every tile of the group's
// unit ranges, walking the whole token's matvec list on its own (so weights of the next operation stream during the
// grid barriers, the attention and the staging passes).
__device__ void producer_loop(const Mega3Params& p, uint32_t ring, uint32_t full, uint32_t empty, int grp, int r) {
    int stage = 0, use = 0;
    const int n_ops = p.n_layers * 4;
    const uint32_t dst_lane = ring + r * kRowPitch;
#pragma unroll 1
    for (int op = 0; op < n_ops; ++op) {
        const MatView m = mat_view(p, op);
        const int nk = m.K / 32;
        int u0, u1;
        unit_range((unsigned)(m.nmat * m.slabs) * (unsigned)nk, grp, u0, u1);
        if (u1 <= u0) continue;
        const size_t row_bytes = (size_t)m.N * 4;
        int vslab = u0 / nk, ks = u0 - vslab * nk;
        const uint8_t* src = nullptr;
        bool fresh = true;
#pragma unroll 1
        for (int u = u0; u < u1; ++u) {
            if (fresh) {
                const int mat = vslab >= m.slabs ? 1 : 0, slab = vslab - mat * m.slabs;
                src = reinterpret_cast<const uint8_t*>(m.w[mat].qw) + (size_t)(ks * 4 + r) * row_bytes + (size_t)slab * (kSlabCols * 4);
                fresh = false;
            }
            if (use > 0) mbar_wait_backoff(empty + stage * 8, (use - 1) & 1u);  // the consumers released the previous use of this stage
            const uint32_t bar = full + stage * 8;
            if (r == 0) mbar_expect_tx(bar, 4096);  // the only pending arrival: the phase cannot complete before it
            bulk_copy_g2s(dst_lane + stage * kTile, src, 1024, bar);
            if (++stage == kStages) {
                stage = 0;
                ++use;
            }
            src += 4 * row_bytes;
            if (++ks == nk) {
                ks = 0;
                ++vslab;
                fresh = true;
            }
        }
    }
}

// ---- grid barrier: monotonic 64-bit arrival counter (never reset: every launch adds a multiple of gridDim.x) ----
__device__ __forceinline__ void grid_barrier(unsigned long long* bar, unsigned long long& target) {
    cta_sync();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(bar) : "memory");
        unsigned long long v;
        do {
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(bar) : "memory");
        } while (v < target);
        fence_acq_rel_gpu();
    }
    cta_sync();
}

__device__ __forceinline__ void zero_slice(float* buf, int n) {
    const int per = ((n / 4 + gridDim.x - 1) / gridDim.x);
    const int lo = blockIdx.x * per, hi = min(n / 4, lo + per);
    for (int i = lo + threadIdx.x; i < hi; i += kConsumers) reinterpret_cast<float4*>(buf)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

__device__ __forceinline__ float block_sum(float v, float* red_s) {
    v = warp_sum(v);
    cta_sync();
    if ((threadIdx.x & 31) == 0) red_s[threadIdx.x >> 5] = v;
    cta_sync();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kCWarps; ++w) t += red_s[w];
    return t;
}

// store 8 consecutive k (natural order, as 4 half2 words) k-permuted: (k0,k4)(k1,k5)(k2,k6)(k3,k7)
__device__ __forceinline__ void store_perm8(__half* dst, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
    uint4 o;
    o.x = __byte_perm(w0, w2, 0x5410);
    o.y = __byte_perm(w0, w2, 0x7632);
    o.z = __byte_perm(w1, w3, 0x5410);
    o.w = __byte_perm(w1, w3, 0x7632);
    *reinterpret_cast<uint4*>(dst) = o;
}

// x = rmsnorm(src [+ fp16(acc)]) for the whole row, staged k-permuted in xs (all 768 consumer threads); the updated
// residual stream is written to resid_out by slices.  ACT: position k' of xs holds feature perm[k'] (norm_w regrouped).
template <bool ACT>
__device__ void stage_norm(const Mega3Params& p, const __half* src, const float* acc, const __half* norm_w, __half* resid_out, __half* xs, __half* tmp,
                           float* red_s, const int32_t* perm) {
    const int H = p.H, tid = threadIdx.x;
    float ss = 0.f;
    for (int c = tid; c < H / 8; c += kConsumers) {
        const uint4 v = *reinterpret_cast<const uint4*>(src + c * 8);
        uint32_t xv[4] = {v.x, v.y, v.z, v.w};
        if (acc != nullptr) {
            const float4 a0 = *reinterpret_cast<const float4*>(acc + c * 8), a1 = *reinterpret_cast<const float4*>(acc + c * 8 + 4);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
            for (int j = 0; j < 4; ++j)  // residual + fp16(linear output): an fp16 add, as in HF's decoder layer
                xv[j] = h2_as_u32(__hadd2(u32_as_h2(xv[j]), __floats2half2_rn(av[2 * j], av[2 * j + 1])));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = __half22float2(u32_as_h2(xv[j]));
            ss = fmaf(f.x, f.x, ss);
            ss = fmaf(f.y, f.y, ss);
        }
        *reinterpret_cast<uint4*>(tmp + c * 8) = make_uint4(xv[0], xv[1], xv[2], xv[3]);
    }
    const float tot = block_sum(ss, red_s);
    const float rstd = 1.0f / sqrtf(tot / (float)H + p.eps);
    const int per = (H / 8 + gridDim.x - 1) / gridDim.x;
    const int wlo = blockIdx.x * per, whi = min(H / 8, wlo + per);
    for (int c = tid; c < H / 8; c += kConsumers) {
        const uint4 v = *reinterpret_cast<const uint4*>(tmp + c * 8);
        if (resid_out != nullptr && c >= wlo && c < whi) *reinterpret_cast<uint4*>(resid_out + c * 8) = v;
        const uint4 nw = *reinterpret_cast<const uint4*>(norm_w + c * 8);
        const uint32_t wv[4] = {nw.x, nw.y, nw.z, nw.w};
        uint32_t xv[4] = {v.x, v.y, v.z, v.w};
        if constexpr (ACT) {
            if (perm != nullptr) {  // gather from shared memory; the index vector and the (regrouped) weights travel together
                const ::int4 p0 = *reinterpret_cast<const ::int4*>(perm + c * 8), p1 = *reinterpret_cast<const ::int4*>(perm + c * 8 + 4);
                const int k[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) xv[j] = h2_as_u32(__halves2half2(tmp[k[2 * j]], tmp[k[2 * j + 1]]));
            }
        }
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 xf = __half22float2(u32_as_h2(xv[j])), wf = __half22float2(u32_as_h2(wv[j]));
            o[j] = h2_as_u32(__floats2half2_rn(__fmul_rn(__fmul_rn(xf.x, rstd), wf.x), __fmul_rn(__fmul_rn(xf.y, rstd), wf.y)));
        }
        store_perm8(xs + c * 8, o[0], o[1], o[2], o[3]);
    }
}

enum XMode { X_FULL = 0, X_ATTN = 1, X_SWIGLU = 2 };

// One matvec operation for one consumer group: consume the tiles of its unit range from the group's ring, RED the results.
// X_FULL: xs holds the whole input row (stage_norm); otherwise the group stages each segment's k-range into its slice xseg.
template <int XMODE, bool ACT>
__device__ void run_matvec(const Mega3Params& p, Pipe& pipe, const MatView& v, int grp, const __half* xs, __half* xseg, const int32_t* perm) {
    const int gtid = threadIdx.x - grp * kGThreads, warp = gtid >> 5, lane = gtid & 31, g = lane >> 2, t = lane & 3;
    const int nk = v.K / 32, N = v.N;
    int u_begin, u_end;
    unit_range((unsigned)(v.nmat * v.slabs) * (unsigned)nk, grp, u_begin, u_end);
    const int gs_steps = p.groupsize >> 5;

    int u = u_begin;
#pragma unroll 1
    while (u < u_end) {
        const int vslab = u / nk;
        const int ks0 = u - vslab * nk;
        const int nsteps = min(nk - ks0, u_end - u);
        const int mat = vslab >= v.slabs ? 1 : 0;
        const int col = (vslab - mat * v.slabs) * kSlabCols + warp * 32 + 4 * g;
        const int zshift = (col & 4) * 4;
        const MatDesc wd = v.w[mat];

        GroupRaw raw;
        GroupConst gc;
        const int grp0 = (ks0 * 32) / p.groupsize;
        const __half* scp = wd.sc + (size_t)grp0 * N + col;
        const uint32_t* qzp = wd.qz + (size_t)grp0 * (N >> 3) + (col >> 3);
        raw = load_group_raw(scp, qzp);
        scp += N;
        qzp += N >> 3;

        uint32_t xaddr;
        if constexpr (XMODE == X_FULL) {
            xaddr = smem_u32(xs) + (ks0 * 32 + t * 8) * 2;
        } else {
            grp_sync(grp);  // previous readers of this group's slice are done
            const int kbeg = ks0 * 32;
            if constexpr (XMODE == X_SWIGLU) {  // h = fp16(silu(acc_gate) * acc_up)  (quant/fused_mlp.py:163-165)
                for (int c = gtid; c < nsteps * 4; c += kGThreads) {
                    const int k = kbeg + c * 8;
                    uint32_t o[4];
                    const float4 g0 = *reinterpret_cast<const float4*>(p.acc_g + k), g1 = *reinterpret_cast<const float4*>(p.acc_g + k + 4);
                    const float4 u0 = *reinterpret_cast<const float4*>(p.acc_u + k), u1 = *reinterpret_cast<const float4*>(p.acc_u + k + 4);
                    o[0] = h2_as_u32(__floats2half2_rn(swiglu(g0.x, u0.x), swiglu(g0.y, u0.y)));
                    o[1] = h2_as_u32(__floats2half2_rn(swiglu(g0.z, u0.z), swiglu(g0.w, u0.w)));
                    o[2] = h2_as_u32(__floats2half2_rn(swiglu(g1.x, u1.x), swiglu(g1.y, u1.y)));
                    o[3] = h2_as_u32(__floats2half2_rn(swiglu(g1.z, u1.z), swiglu(g1.w, u1.w)));
                    store_perm8(xseg + c * 8, o[0], o[1], o[2], o[3]);
                }
            } else {  // attention output: one thread per feature combines the split-KV partials of its head
                const int nvalid = min(p.nsplit, p.positions[0] / kAttnChunk + 1);
                for (int e = gtid; e < nsteps * 32; e += kGThreads) {
                    int k = kbeg + e;
                    if constexpr (ACT) {
                        if (perm != nullptr) k = perm[k];  // regrouped rows: position k' of the matvec input is attention feature perm[k']
                    }
                    const int head = k / kHD, d = k - head * kHD;
                    const float* src = p.part + (size_t)head * p.nsplit * kRec;
                    float M = -INFINITY;
                    for (int sI = 0; sI < nvalid; ++sI) M = fmaxf(M, src[(size_t)sI * kRec]);
                    float L = 0.f, O = 0.f;
#pragma unroll 4
                    for (int sI = 0; sI < nvalid; ++sI) {
                        const float* ps = src + (size_t)sI * kRec;
                        const float wgt = expf(ps[0] - M);
                        L = fmaf(ps[1], wgt, L);
                        O = fmaf(ps[4 + d], wgt, O);
                    }
                    const int j8 = e & 7;
                    xseg[(e & ~7) + ((j8 & 3) << 1) + (j8 >> 2)] = __float2half_rn(O / L);  // k-permuted position inside the run of 8
                }
            }
            grp_sync(grp);
            xaddr = smem_u32(xseg) + (t * 8) * 2;
        }

        float acc[2][4];
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[h][i] = 0.f;

        build_group_const(gc, raw, zshift);
        int steps_left_in_grp = gs_steps - (ks0 % gs_steps);
        if (steps_left_in_grp < nsteps) {
            raw = load_group_raw(scp, qzp);
            scp += N;
            qzp += N >> 3;
        }

        // software pipeline: the NEXT tile's wait + shared-memory load are issued before the current tile's math
        auto fetch = [&](uint4& q, uint32_t& bar_of_q) {
            mbar_wait(pipe.bar, pipe.parity);  // the tile has landed
            q = lds128(pipe.tile);
            bar_of_q = pipe.bar;
            pipe.tile += kTile;
            pipe.bar += 8;
            if (--pipe.left == 0) {  // ring wrap: next round, other parity
                pipe.left = kStages;
                pipe.tile -= kStages * kTile;
                pipe.bar -= kStages * 8;
                pipe.parity ^= 1u;
            }
        };
        uint4 q_cur;
        uint32_t bar_cur;
        fetch(q_cur, bar_cur);
#pragma unroll 2
        for (int step = 0; step < nsteps; ++step) {
            if (steps_left_in_grp == 0) {
                build_group_const(gc, raw, zshift);
                steps_left_in_grp = gs_steps;
                if (step + gs_steps < nsteps) {
                    raw = load_group_raw(scp, qzp);
                    scp += N;
                    qzp += N >> 3;
                }
            }
            --steps_left_in_grp;
            const uint4 xf = lds128(xaddr);
            xaddr += 64;
            uint4 q_next = q_cur;
            uint32_t bar_next = bar_cur;
            if (step + 1 < nsteps) fetch(q_next, bar_next);
            uint32_t wf[4][4];
            dequant8<0>(q_cur.x, gc.za01, gc.zb01, gc.s01, wf[0]);
            dequant8<1>(q_cur.y, gc.za01, gc.zb01, gc.s01, wf[1]);
            dequant8<0>(q_cur.z, gc.za23, gc.zb23, gc.s23, wf[2]);
            dequant8<1>(q_cur.w, gc.za23, gc.zb23, gc.s23, wf[3]);
            mma_16816(acc[0], wf[0][0], wf[1][0], wf[0][1], wf[1][1], xf.x, xf.y);
            mma_16816(acc[0], wf[0][2], wf[1][2], wf[0][3], wf[1][3], xf.z, xf.w);
            mma_16816(acc[1], wf[2][0], wf[3][0], wf[2][1], wf[3][1], xf.x, xf.y);
            mma_16816(acc[1], wf[2][2], wf[3][2], wf[2][3], wf[3][3], xf.z, xf.w);
            __syncwarp();  // every lane has consumed the registers it read from the current tile's stage
            if (lane == 0) mbar_arrive(bar_cur + kStages * 8);  // empty[stage]
            q_cur = q_next;
            bar_cur = bar_next;
        }

        // batch row 0 lives in the t == 0 lanes: acc[0][0] -> col, [0][2] -> col+1, [1][0] -> col+2, [1][2] -> col+3
        if (t == 0) {
            float* o0 = v.out[mat] + col;
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o0), "f"(acc[0][0]), "f"(acc[0][2]), "f"(acc[1][0]), "f"(acc[1][2]) : "memory");
        }
        u += nsteps;
    }
}

// Attention work items (head, split) over the consumer GROUPS: RoPE(q,k) from this step's cos/sin, KV append, partial
// softmax(qK^T)V in passes of 64 keys (two keys per 8-lane thread group and pass), online-softmax merge across passes.
__device__ void run_attention(const Mega3Params& p, int layer, int grp, float* smem_f) {
    const int gtid = threadIdx.x - grp * kGThreads;
    const int pos = p.positions[0];
    const int T = pos + 1;
    float* q_s = smem_f;          // [128]
    float* red_m = smem_f + 128;  // [32]
    float* red_l = smem_f + 160;  // [32]
    float* red_o = smem_f + 192;  // [32][132]
    __half* kc_base = p.k_cache + layer * p.layer_stride;
    __half* vc_base = p.v_cache + layer * p.layer_stride;
    const int n_items = p.n_heads * p.nsplit;
    const int nbg = gridDim.x * kGroups;
#pragma unroll 1
    for (int item = blockIdx.x * kGroups + grp; item < n_items; item += nbg) {
        const int head = item / p.nsplit, split = item - head * p.nsplit;
        const int c0 = split * kAttnChunk;
        if (c0 >= T) continue;
        const int c1 = min(c0 + kAttnChunk, T);
        __half* kc = kc_base + (size_t)head * p.max_seq * kHD;
        __half* vc = vc_base + (size_t)head * p.max_seq * kHD;
        grp_sync(grp);  // smem reuse across items
        if (gtid < kHD) {
            const int i = gtid & 63;
            const bool hi = gtid >= 64;
            const float c = p.rope_cs[i], s = p.rope_cs[64 + i];
            const float* aq = p.acc_qkv + head * kHD;
            const float qx = __half2float(__float2half_rn(aq[i])), qy = __half2float(__float2half_rn(aq[i + 64]));  // the qkv projection output is fp16
            const float qr = hi ? __fadd_rn(__fmul_rn(qx, s), __fmul_rn(qy, c)) : __fsub_rn(__fmul_rn(qx, c), __fmul_rn(qy, s));
            q_s[gtid] = __half2float(__float2half_rn(qr));
            if (pos >= c0 && pos < c1) {  // this item owns the new key/value: append them
                const float* ak = aq + p.H;
                const float* av = aq + 2 * p.H;
                const float kx = __half2float(__float2half_rn(ak[i])), ky = __half2float(__float2half_rn(ak[i + 64]));
                const float kr = hi ? __fadd_rn(__fmul_rn(kx, s), __fmul_rn(ky, c)) : __fsub_rn(__fmul_rn(kx, c), __fmul_rn(ky, s));
                kc[(size_t)pos * kHD + gtid] = __float2half_rn(kr);
                vc[(size_t)pos * kHD + gtid] = __float2half_rn(av[gtid]);
            }
        }
        grp_sync(grp);
        const int tg = gtid >> 3, j = gtid & 7;  // 32 thread groups of 8 lanes; lane j owns dims [16j, 16j+16)
        float qr[16];
#pragma unroll
        for (int d = 0; d < 16; ++d) qr[d] = q_s[16 * j + d];
        float mloc = -INFINITY, lloc = 0.f, o[16];
#pragma unroll
        for (int d = 0; d < 16; ++d) o[d] = 0.f;
#pragma unroll 1
        for (int p0 = c0; p0 < c1; p0 += kAttnPass) {
            uint4 kreg[kAttnIter][2], vreg[kAttnIter][2];
#pragma unroll
            for (int it = 0; it < kAttnIter; ++it) {
                const int tk = min(p0 + tg + it * 32, c1 - 1);
                const uint4* kp = reinterpret_cast<const uint4*>(kc + (size_t)tk * kHD + 16 * j);
                kreg[it][0] = kp[0];
                kreg[it][1] = kp[1];
            }
#pragma unroll
            for (int it = 0; it < kAttnIter; ++it) {
                const int tk = min(p0 + tg + it * 32, c1 - 1);
                const uint4* vp = reinterpret_cast<const uint4*>(vc + (size_t)tk * kHD + 16 * j);
                vreg[it][0] = vp[0];
                vreg[it][1] = vp[1];
            }
            float sc[kAttnIter];
            float mpass = -INFINITY;
#pragma unroll
            for (int it = 0; it < kAttnIter; ++it) {
                const int tk = p0 + tg + it * 32;
                const uint32_t w[8] = {kreg[it][0].x, kreg[it][0].y, kreg[it][0].z, kreg[it][0].w, kreg[it][1].x, kreg[it][1].y, kreg[it][1].z, kreg[it][1].w};
                float s = 0.f;
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float2 f = __half22float2(u32_as_h2(w[e]));
                    s = fmaf(qr[2 * e], f.x, s);
                    s = fmaf(qr[2 * e + 1], f.y, s);
                }
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                s += __shfl_xor_sync(0xffffffffu, s, 4);
                s = (tk < c1) ? s * p.scale : -INFINITY;
                sc[it] = s;
                mpass = fmaxf(mpass, s);
            }
            const float mnew = fmaxf(mloc, mpass);
            if (mnew != -INFINITY) {
                const float alpha = (mloc == -INFINITY) ? 0.f : expf(mloc - mnew);
                lloc *= alpha;
#pragma unroll
                for (int d = 0; d < 16; ++d) o[d] *= alpha;
#pragma unroll
                for (int it = 0; it < kAttnIter; ++it) {
                    const int tk = p0 + tg + it * 32;
                    const float pw = (tk < c1) ? expf(sc[it] - mnew) : 0.f;
                    lloc += pw;
                    const uint32_t w[8] = {vreg[it][0].x, vreg[it][0].y, vreg[it][0].z, vreg[it][0].w, vreg[it][1].x, vreg[it][1].y, vreg[it][1].z, vreg[it][1].w};
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float2 f = __half22float2(u32_as_h2(w[e]));
                        o[2 * e] = fmaf(pw, f.x, o[2 * e]);
                        o[2 * e + 1] = fmaf(pw, f.y, o[2 * e + 1]);
                    }
                }
                mloc = mnew;
            }
        }
        if (j == 0) {
            red_m[tg] = mloc;
            red_l[tg] = lloc;
        }
#pragma unroll
        for (int d = 0; d < 16; ++d) red_o[tg * 132 + 16 * j + d] = o[d];
        grp_sync(grp);
        if (gtid < kHD) {
            float M = -INFINITY;
#pragma unroll 8
            for (int gI = 0; gI < 32; ++gI) M = fmaxf(M, red_m[gI]);
            float L = 0.f, O = 0.f;
#pragma unroll 8
            for (int gI = 0; gI < 32; ++gI) {
                const float wgt = (red_m[gI] == -INFINITY) ? 0.f : expf(red_m[gI] - M);
                L = fmaf(red_l[gI], wgt, L);
                O = fmaf(red_o[gI * 132 + gtid], wgt, O);
            }
            float* dst = p.part + ((size_t)head * p.nsplit + split) * kRec;
            dst[4 + gtid] = O;
            if (gtid == 0) {
                dst[0] = M;
                dst[1] = L;
            }
        }
    }
}

template <bool ACT>
__global__ void __launch_bounds__(kBlock, 1) llama_decode_mega3_kernel(const __grid_constant__ Mega3Params p) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ float red_s[kCWarps];
    __shared__ __align__(8) unsigned long long bars_s[kGroups][2 * kStages];
    const int tid = threadIdx.x, lane = tid & 31;
    // smem: [rings: 3 x 12 x 4224 B][region R]; R = xs (xs_halves) + tmp (H halves), aliased by the attention scratch
    // (3 x 17,664 B): xs / tmp are dead between the qkv matvec and the o_proj staging
    constexpr int kRingBytes = ((kGroups * kStages * kTile + 127) / 128) * 128;
    __half* xs = reinterpret_cast<__half*>(smem_raw + kRingBytes);
    __half* tmp = xs + p.xs_halves;
    float* attn_scratch = reinterpret_cast<float*>(smem_raw + kRingBytes);
    if (tid == 0) {
        for (int g = 0; g < kGroups; ++g)
            for (int s = 0; s < kStages; ++s) {
                mbar_init(smem_u32(&bars_s[g][s]), 1);
                mbar_init(smem_u32(&bars_s[g][kStages + s]), kGWarps);
            }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();  // the only block-wide barrier: after it the producer warp and the consumers never meet again
    if (tid >= kConsumers) {
        if (lane < 4 * kGroups) {
            const int g = lane >> 2;
            producer_loop(p, smem_u32(smem_raw) + g * (kStages * kTile), smem_u32(&bars_s[g][0]), smem_u32(&bars_s[g][kStages]), g, lane & 3);
        }
        return;
    }
    const int grp = tid / kGThreads, gtid = tid - grp * kGThreads;
    Pipe pipe;
    pipe.tile = smem_u32(smem_raw) + grp * (kStages * kTile) + (lane & 3) * kRowPitch + (gtid >> 5) * 128 + (lane >> 2) * 16;
    pipe.bar = smem_u32(&bars_s[grp][0]);
    pipe.parity = 0;
    pipe.left = kStages;
    __half* xseg = xs + grp * kSegHalves;

    unsigned long long gen;  // barrier target (meaningful in thread 0): the counter value when this launch began
    {
        unsigned long long v;
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p.bar) : "memory");
        gen = v - (v % gridDim.x);  // CTAs that already arrived at the first barrier have added < gridDim.x
    }

    // this step's RoPE angles (quant/fused_attn.py:43,91): freq_i = exp(i * inv_base) * pos
    if (blockIdx.x == 0 && tid < 64) {
        const float f = expf((float)tid * p.inv_base) * (float)p.positions[0];
        p.rope_cs[tid] = cosf(f);
        p.rope_cs[64 + tid] = sinf(f);
    }

    const __half* resid_src = p.embed + (size_t)p.tokens[0] * p.H;
    const float* resid_acc = nullptr;
    int cur = 1;
#pragma unroll 1
    for (int l = 0; l < p.n_layers; ++l) {
        const LayerDesc& L = p.layers[l];
        // ---- Q ----
        MTRACE(l * 12 + 0);
        stage_norm<ACT>(p, resid_src, resid_acc, L.input_norm, p.resid[cur ^ 1], xs, tmp, red_s, L.qkv_perm);
        MTRACE(l * 12 + 1);
        cur ^= 1;
        zero_slice(p.acc_g, p.I);  // last read by the previous layer's D
        zero_slice(p.acc_u, p.I);
        cta_sync();
        run_matvec<X_FULL, ACT>(p, pipe, mat_view(p, l * 4 + 0), grp, xs, xseg, nullptr);
        MTRACE(l * 12 + 2);
        grid_barrier(p.bar, gen);
        MTRACE(l * 12 + 3);
        // ---- A ----
        zero_slice(p.acc_d, p.H);  // last read by this layer's Q
        run_attention(p, l, grp, attn_scratch + grp * kAttnScratchFloats);
        MTRACE(l * 12 + 4);
        grid_barrier(p.bar, gen);
        MTRACE(l * 12 + 5);
        // ---- O ----
        zero_slice(p.acc_qkv, 3 * p.H);
        run_matvec<X_ATTN, ACT>(p, pipe, mat_view(p, l * 4 + 1), grp, xs, xseg, L.o_perm);
        MTRACE(l * 12 + 6);
        grid_barrier(p.bar, gen);
        MTRACE(l * 12 + 7);
        // ---- G ----
        stage_norm<ACT>(p, p.resid[cur], p.acc_o, L.post_norm, p.resid[cur ^ 1], xs, tmp, red_s, L.mlp_perm);
        cur ^= 1;
        cta_sync();
        MTRACE(l * 12 + 8);
        run_matvec<X_FULL, ACT>(p, pipe, mat_view(p, l * 4 + 2), grp, xs, xseg, nullptr);
        MTRACE(l * 12 + 9);
        grid_barrier(p.bar, gen);
        // ---- D ----
        zero_slice(p.acc_o, p.H);
        MTRACE(l * 12 + 10);
        run_matvec<X_SWIGLU, ACT>(p, pipe, mat_view(p, l * 4 + 3), grp, xs, xseg, nullptr);
        MTRACE(l * 12 + 11);
        grid_barrier(p.bar, gen);
        resid_src = p.resid[cur];
        resid_acc = p.acc_d;
    }
    // ---- L: final norm + lm_head (fp16 [V, H] rows, one warp per row) ----
    stage_norm<false>(p, resid_src, resid_acc, p.final_norm, nullptr, xs, tmp, red_s, nullptr);
    zero_slice(p.acc_g, p.I);
    zero_slice(p.acc_u, p.I);
    cta_sync();
    {
        const int chunks = p.H / 8, cwarp = tid >> 5;
#pragma unroll 1
        for (int row = blockIdx.x * kCWarps + cwarp; row < p.V; row += gridDim.x * kCWarps) {
            const uint4* wr = reinterpret_cast<const uint4*>(p.lm_head + (size_t)row * p.H);
            float a = 0.f;
#pragma unroll 4
            for (int c = lane; c < chunks; c += 32) {
                uint4 wv;
                asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(wv.x), "=r"(wv.y), "=r"(wv.z), "=r"(wv.w) : "l"(wr + c));
                const uint4 xv = *reinterpret_cast<const uint4*>(xs + c * 8);  // k-permuted: (k0,k4)(k1,k5)(k2,k6)(k3,k7)
                const uint32_t wp[4] = {__byte_perm(wv.x, wv.z, 0x5410), __byte_perm(wv.x, wv.z, 0x7632), __byte_perm(wv.y, wv.w, 0x5410),
                                        __byte_perm(wv.y, wv.w, 0x7632)};
                const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 wf = __half22float2(u32_as_h2(wp[e])), xf = __half22float2(u32_as_h2(xw[e]));
                    a = fmaf(wf.x, xf.x, a);
                    a = fmaf(wf.y, xf.y, a);
                }
            }
            a = warp_sum(a);
            if (lane == 0) p.logits[row] = __float2half_rn(a);
        }
    }
    grid_barrier(p.bar, gen);
    zero_slice(p.acc_d, p.H);
    if (blockIdx.x == 0 && p.next_token != nullptr) {  // greedy argmax (lowest index wins ties)
        float best = -INFINITY;
        int idx = 0x7fffffff;
        for (int i = tid; i < p.V; i += kConsumers) {
            const float v = __half2float(p.logits[i]);
            if (v > best || (v == best && i < idx)) {
                best = v;
                idx = i;
            }
        }
        __shared__ float sv[kCWarps];
        __shared__ int si[kCWarps];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
            if (ov > best || (ov == best && oi < idx)) {
                best = ov;
                idx = oi;
            }
        }
        if (lane == 0) {
            sv[tid >> 5] = best;
            si[tid >> 5] = idx;
        }
        cta_sync();
        if (tid == 0) {
            for (int w = 1; w < kCWarps; ++w)
                if (sv[w] > best || (sv[w] == best && si[w] < idx)) {
                    best = sv[w];
                    idx = si[w];
                }
            p.next_token[0] = idx;
        }
    }
}

inline size_t al256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace

// Same scratch layout as launch_decode_mega (mega_scratch_bytes): the two layouts are interchangeable per launch.
cudaError_t launch_decode_mega3(const gptq_llama_model& m, const gptq_llama_state& st, uint8_t* scratch, cudaStream_t stream) {
    static_assert(sizeof(Mega3Params) < 32000, "kernel parameter space");
    Mega3Params p{};
    p.n_layers = m.n_layers; p.H = m.hidden; p.I = m.intermediate; p.V = m.vocab; p.n_heads = m.n_heads;
    p.groupsize = m.layers[0].qkv.groupsize;
    p.max_seq = st.max_seq;
    p.nsplit = ceil_div(st.max_seq, kAttnChunk);
    p.eps = m.rms_eps;
    p.inv_base = (float)(-2.0 * log((double)m.rope_base) / (double)m.head_dim);
    p.scale = 1.0f / sqrtf((float)m.head_dim);
    p.embed = reinterpret_cast<const __half*>(m.embed);
    p.final_norm = reinterpret_cast<const __half*>(m.final_norm);
    p.lm_head = reinterpret_cast<const __half*>(m.lm_head);
    p.tokens = st.tokens;
    p.positions = st.positions;
    p.k_cache = reinterpret_cast<__half*>(st.k_cache);
    p.v_cache = reinterpret_cast<__half*>(st.v_cache);
    p.layer_stride = (size_t)m.n_heads * st.max_seq * m.head_dim;
    p.logits = reinterpret_cast<__half*>(st.logits);
    p.next_token = st.next_tokens;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        uint8_t* q = scratch + off;
        off += al256(bytes);
        return q;
    };
    p.resid[0] = reinterpret_cast<__half*>(take((size_t)m.hidden * 2));
    p.resid[1] = reinterpret_cast<__half*>(take((size_t)m.hidden * 2));
    p.acc_qkv = reinterpret_cast<float*>(take((size_t)3 * m.hidden * 4));
    p.acc_o = reinterpret_cast<float*>(take((size_t)m.hidden * 4));
    p.acc_d = reinterpret_cast<float*>(take((size_t)m.hidden * 4));
    p.acc_g = reinterpret_cast<float*>(take((size_t)m.intermediate * 4));
    p.acc_u = reinterpret_cast<float*>(take((size_t)m.intermediate * 4));
    p.part = reinterpret_cast<float*>(take((size_t)m.n_heads * p.nsplit * kRec * 4));
    p.rope_cs = reinterpret_cast<float*>(take(128 * 4));
    p.bar = reinterpret_cast<unsigned long long*>(take(256));
    bool act = false;
    for (int l = 0; l < m.n_layers; ++l) {
        const gptq_llama_layer& ly = m.layers[l];
        auto md = [](const gptq_qweight& w) {
            MatDesc d;
            d.qw = reinterpret_cast<const uint32_t*>(w.qweight);
            d.sc = reinterpret_cast<const __half*>(w.scales);
            d.qz = reinterpret_cast<const uint32_t*>(w.qzeros);
            return d;
        };
        p.layers[l].qkv = md(ly.qkv);
        p.layers[l].o = md(ly.o);
        p.layers[l].gate = md(ly.gate);
        p.layers[l].up = md(ly.up);
        p.layers[l].down = md(ly.down);
        p.layers[l].input_norm = reinterpret_cast<const __half*>(ly.input_norm);
        p.layers[l].post_norm = reinterpret_cast<const __half*>(ly.post_norm);
        p.layers[l].qkv_perm = ly.qkv_perm;
        p.layers[l].o_perm = ly.o_perm;
        p.layers[l].mlp_perm = ly.mlp_perm;
        act = act || ly.qkv_perm != nullptr || ly.o_perm != nullptr || ly.mlp_perm != nullptr;
    }
    int dev = 0, sms = 0, occ = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return cudaErrorInvalidDevice;
    p.xs_halves = max(m.hidden, kGroups * kSegHalves);
    const size_t region = max((size_t)(p.xs_halves + m.hidden) * 2, (size_t)kGroups * kAttnScratchFloats * 4);
    const size_t smem = (size_t)(((kGroups * kStages * kTile + 127) / 128) * 128) + region;
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    // every segment a group stages (a contiguous k-range inside one slab) must fit its slice of xs
    const long long groups = (long long)sms * kGroups;
    const long long seg_o = ((long long)(m.hidden / kSlabCols) * (m.hidden / 32) + groups - 1) / groups;
    const long long seg_d = ((long long)(m.hidden / kSlabCols) * (m.intermediate / 32) + groups - 1) / groups;
    if (max(seg_o, seg_d) * 32 > kSegHalves) return cudaErrorInvalidConfiguration;
    auto kernel = act ? llama_decode_mega3_kernel<true> : llama_decode_mega3_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kBlock, smem) != cudaSuccess || occ < 1) return cudaErrorCooperativeLaunchTooLarge;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(sms);
    cfg.blockDim = dim3(kBlock);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;  // all CTAs co-resident: the grid barrier cannot deadlock
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, p);
}

}  // namespace gptq

#ifdef GPTQ_TRACE
extern "C" int gptq_debug_set_mega3_trace(void* buf) {
    unsigned long long* b = reinterpret_cast<unsigned long long*>(buf);
    return (int)cudaMemcpyToSymbol(gptq::g_mega3_trace, &b, sizeof(b));
}
#endif
