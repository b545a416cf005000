// Persistent decode step: ONE cooperative kernel per token (batch 1, int4, no act-order).
//
// The kernel-chain engine (decode.cu) spends most of a token in per-kernel latencies (parameter fetch, first
// DRAM round trip, split-K fix-up, launch drain): ~12-23 us per matvec whose weights stream in 1.5-7 us.  Here
// all 296 CTAs (2 per SM) stay resident for the whole token and walk the same list of operations:
//
//   per layer:  Q  qkv matvec      x = rmsnorm(resid [+ fp16(acc_down)])           -> RED acc_qkv
//               A  attention       q,k,v = fp16(acc_qkv); RoPE; KV append; split-KV -> part
//               O  o_proj matvec   x = combine(part)                                -> RED acc_o
//               G  gate/up matvec  x = rmsnorm(resid + fp16(acc_o))                 -> RED acc_gate, acc_up
//               D  down matvec     x = fp16(silu(acc_gate) * acc_up)                -> RED acc_down
//   then        L  lm_head         x = rmsnorm(resid + fp16(acc_down)), fp16 rows   -> logits;   argmax
//
// separated by grid barriers (one atomic + one polled generation word, ~1.5 us).  Split-K partial sums are
// accumulated with red.global.add.f32 into fp32 vectors that the NEXT operation rounds to fp16 exactly where the
// reference rounds (a QuantLinear output is fp16), so there is no fix-up pass; each vector is re-zeroed one
// operation after its last reader.  The weight ring (cp.async, 16 x 512 B per warp, int4_core.cuh) is fed by a
// producer cursor that runs ahead ACROSS operations: while a CTA waits at a barrier, the next matvec's weights
// are already in flight, so HBM stays busy through the synchronisation points.
//
// Arithmetic per element is the same as in qmatvec.cu / decode.cu (reference-exact dequant, fp32 accumulate);
// only the fp32 summation order of the <= 20 split-K partials is unordered (atomics).
#include <cstdlib>

#include "common.cuh"
#include "int4_core.cuh"
#include "kernels.h"

namespace gptq {
namespace {

using namespace int4;

constexpr int kWarps = 8;
constexpr int kThreads = 256;
constexpr int kSlabCols = 256;
constexpr int kProducers = 2;                      // producer warps: tile i is fetched by producer i % kProducers
constexpr int kBlock = kThreads + 32 * kProducers;  // 8 consumer warps + the producer warps
constexpr int kRowPitch = 1024 + 32;   // smem pitch of a 1 KB weight row: +32 B makes the 4 rows of a k-step hit distinct bank groups
constexpr int kTile = 4 * kRowPitch;   // one ring stage: 4 packed rows x 256 columns of one matrix (4 KB of weights)
constexpr int kStages = 15;            // 15 x 4224 B = 63,360 B of weights in flight per CTA
constexpr int kHD = 128;
constexpr int kAttnChunk = 256;  // keys per attention work item (two passes of 128 with an online-softmax merge)
constexpr int kAttnPass = 128;
constexpr int kRec = kHD + 4;     // floats per split-KV partial record: m, l, 2 pad, o[128] (keeps o 16-byte aligned)
constexpr int kMaxLayers = 80;
#ifndef GPTQ_STEP_UNROLL
#define GPTQ_STEP_UNROLL 2
#endif
constexpr int kStepUnroll = GPTQ_STEP_UNROLL;  // unroll factor of the per-tile loop (development knob)

#ifdef GPTQ_TRACE
}  // namespace
__device__ unsigned long long* g_mega_trace = nullptr;
namespace {

#define MTRACE(id)                                                                                                       \
    do {                                                                                                                 \
        if (g_mega_trace != nullptr && threadIdx.x == 0 && (id) < 64) {                                                   \
            unsigned long long t_;                                                                                       \
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_));                                                        \
            g_mega_trace[blockIdx.x * 64 + (id)] = t_;                                                                   \
        }                                                                                                                \
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
struct MegaParams {
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
    // scratch (device)
    __half* resid[2];  // residual stream ping-pong, fp16 [H]
    float* acc_qkv;    // [3H]
    float* acc_o;      // [H]
    float* acc_g;      // [I]
    float* acc_u;      // [I]
    float* acc_d;      // [H]
    float* part;       // [heads][nsplit][130]
    float* rope_cs;    // [128]: cos[64], sin[64] of this step's position
    unsigned long long* bar;  // monotonic arrival counter of the grid barrier
    LayerDesc layers[kMaxLayers];
};

// One matvec of the op list, as the weight producer sees it.
struct MatOp {
    const uint32_t* qw[2];
    int K, N, ntiles_per_step;  // 2 for the fused gate/up
};

__device__ __forceinline__ MatOp mat_op(const MegaParams& p, int idx) {
    // matvec ops in execution order: 4 per layer (qkv, o, gate|up, down)
    const LayerDesc& L = p.layers[idx >> 2];
    MatOp m;
    m.ntiles_per_step = 1;
    m.qw[1] = nullptr;
    switch (idx & 3) {
        case 0: m.qw[0] = L.qkv.qw; m.K = p.H; m.N = 3 * p.H; break;
        case 1: m.qw[0] = L.o.qw; m.K = p.H; m.N = p.H; break;
        case 2: m.qw[0] = L.gate.qw; m.qw[1] = L.up.qw; m.K = p.H; m.N = p.I; m.ntiles_per_step = 2; break;
        default: m.qw[0] = L.down.qw; m.K = p.I; m.N = p.H; break;
    }
    return m;
}

// ---- pipeline state --------------------------------------------------------------------------------------
// A dedicated producer warp streams 4 KB tiles (4 packed rows x 1 KB: whole DRAM-page-sized row segments, one
// cp.async.bulk each) into a 15-stage ring shared by the CTA; the 8 consumer warps each read their 32-column
// stripe of every tile.  full[s]: producer -> consumers (expect_tx 4096 B); empty[s]: 8 consumer warps -> producer.
// The producer walks the matvec list of the whole token on its own, so it runs ahead across grid barriers.
struct Pipe {
    uint32_t ring;   // smem address of stage 0
    uint32_t full;   // smem address of full[0]  (8 B each; empty[s] sits kStages * 8 bytes after full[s])
    uint32_t empty;  // smem address of empty[0]
    // consumer cursors (per thread): stages are consumed in strict rotation, so one parity bit per round suffices
    uint32_t tile;    // smem address of this lane's 16 B in the current stage
    uint32_t bar;     // smem address of full[current stage]
    uint32_t parity;  // expected parity of the current round
    int left;         // stages until the ring wraps
};

__device__ __forceinline__ void cta_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }  // the 8 consumer warps only

__device__ void producer_loop(const MegaParams& p, uint32_t ring, uint32_t full, uint32_t empty, int pw) {
    int stage = 0, use = 0, mine = 0;  // ring position / use count of the next tile; mine: tiles until this producer's turn
    mine = pw;
    const int n_ops = p.n_layers * 4;
#pragma unroll 1
    for (int op = 0; op < n_ops; ++op) {
        const MatOp m = mat_op(p, op);
        const unsigned nk = m.K / 32, U = (unsigned)(m.N / kSlabCols) * nk, nb = gridDim.x;
        const int u0 = (int)((blockIdx.x * U) / nb), u1 = (int)(((blockIdx.x + 1) * U) / nb);
        if (u1 <= u0) continue;
        const int slab0 = u0 / nk;
        int ks = u0 - slab0 * nk;
        const size_t row_bytes = (size_t)m.N * 4;
        const uint8_t* src[2];
        src[0] = reinterpret_cast<const uint8_t*>(m.qw[0]) + (size_t)(ks * 4) * row_bytes + (size_t)slab0 * (kSlabCols * 4);
        src[1] = m.qw[1] ? reinterpret_cast<const uint8_t*>(m.qw[1]) + (size_t)(ks * 4) * row_bytes + (size_t)slab0 * (kSlabCols * 4) : nullptr;
        const long long wrap = (long long)(kSlabCols * 4) - (long long)nk * 4 * (long long)row_bytes;  // next slab, back to packed row 0
#pragma unroll 1
        for (int u = u0; u < u1; ++u) {
#pragma unroll 1
            for (int w = 0; w < m.ntiles_per_step; ++w) {
                if (mine == 0) {
                    if (use > 0) mbar_wait(empty + stage * 8, (use - 1) & 1u);  // consumers released the previous use of this stage
                    const uint32_t bar = full + stage * 8;
                    const uint32_t dst = ring + stage * kTile;
                    mbar_expect_tx(bar, 4096);
#pragma unroll
                    for (int r = 0; r < 4; ++r) bulk_copy_g2s(dst + r * kRowPitch, src[w] + r * row_bytes, 1024, bar);
                    mine = kProducers;
                }
                --mine;
                if (++stage == kStages) {
                    stage = 0;
                    ++use;
                }
            }
            src[0] += 4 * row_bytes;
            if (src[1]) src[1] += 4 * row_bytes;
            if (++ks == (int)nk) {
                ks = 0;
                src[0] += wrap;
                if (src[1]) src[1] += wrap;
            }
        }
    }
}

// ---- grid barrier: one monotonic 64-bit arrival counter (never reset: every launch adds a multiple of gridDim.x) ----
// arrive = red.release (fire and forget), wait = poll the same word with ld.acquire until it reaches this barrier's
// target: about 1.5 L2 round trips.  `target` lives in thread 0 of each CTA.
__device__ __forceinline__ void grid_barrier(unsigned long long* bar, unsigned long long& target) {
    cta_sync();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(bar) : "memory");
        unsigned long long v;
        do {
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(bar) : "memory");
        } while (v < target);
        fence_acq_rel_gpu();  // also drops this SM's stale L1 lines of data other CTAs have rewritten
    }
    cta_sync();
}

__device__ __forceinline__ void zero_slice(float* buf, int n) {
    // this CTA's share of a distributed memset (n is a multiple of 4)
    const int per = ((n / 4 + gridDim.x - 1) / gridDim.x);
    const int lo = blockIdx.x * per, hi = min(n / 4, lo + per);
    for (int i = lo + threadIdx.x; i < hi; i += kThreads) reinterpret_cast<float4*>(buf)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

__device__ __forceinline__ float block_sum(float v, float* red_s) {
    v = warp_sum(v);
    cta_sync();
    if ((threadIdx.x & 31) == 0) red_s[threadIdx.x >> 5] = v;
    cta_sync();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) t += red_s[w];
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

// x = rmsnorm(src [+ fp16(acc)]) for the whole row (K = H), staged k-permuted in xs; the updated residual stream
// (src + fp16(acc)) is written to resid_out by slices.  src/acc/resid_out are global.
// ACT: the matvec's packed rows were regrouped by the host (act-order): position k' of xs holds feature perm[k'].
template <bool ACT>
__device__ void stage_norm(const MegaParams& p, const __half* src, const float* acc, const __half* norm_w, __half* resid_out, __half* xs, __half* tmp,
                           float* red_s, const int32_t* perm) {
    const int H = p.H, tid = threadIdx.x;
    float ss = 0.f;
    for (int c = tid; c < H / 8; c += kThreads) {
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
    if constexpr (ACT) {
        if (perm != nullptr) {
            // norm_w is given in regrouped order (gptq_b200.h): the index vector and the weights travel together, the gather
            // itself reads shared memory
            for (int c = tid; c < H / 8; c += kThreads) {
                if (resid_out != nullptr && c >= wlo && c < whi) *reinterpret_cast<uint4*>(resid_out + c * 8) = *reinterpret_cast<const uint4*>(tmp + c * 8);
                const ::int4 p0 = *reinterpret_cast<const ::int4*>(perm + c * 8), p1 = *reinterpret_cast<const ::int4*>(perm + c * 8 + 4);
                const uint4 nw = *reinterpret_cast<const uint4*>(norm_w + c * 8);
                const int k[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
                const uint32_t wv[4] = {nw.x, nw.y, nw.z, nw.w};
                uint32_t o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 wf = __half22float2(u32_as_h2(wv[j]));
                    const float x0 = __half2float(tmp[k[2 * j]]), x1 = __half2float(tmp[k[2 * j + 1]]);
                    o[j] = h2_as_u32(__floats2half2_rn(__fmul_rn(__fmul_rn(x0, rstd), wf.x), __fmul_rn(__fmul_rn(x1, rstd), wf.y)));
                }
                store_perm8(xs + c * 8, o[0], o[1], o[2], o[3]);
            }
            return;
        }
    }
    for (int c = tid; c < H / 8; c += kThreads) {
        const uint4 v = *reinterpret_cast<const uint4*>(tmp + c * 8);
        if (resid_out != nullptr && c >= wlo && c < whi) *reinterpret_cast<uint4*>(resid_out + c * 8) = v;
        const uint4 nw = *reinterpret_cast<const uint4*>(norm_w + c * 8);
        const uint32_t xv[4] = {v.x, v.y, v.z, v.w}, wv[4] = {nw.x, nw.y, nw.z, nw.w};
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

// One matvec op for this CTA: consume the tiles of its unit range from the ring, RED the results.
// xs holds either the full K row (X_FULL, staged by stage_norm before the call) or is (re)staged per segment here.
template <bool DUAL, int XMODE, bool ACT = false>
__device__ void run_matvec(const MegaParams& p, Pipe& pipe, const MatDesc& w0, const MatDesc& w1, int K, int N,
                           float* out0, float* out1, __half* xs, const int32_t* perm = nullptr) {
    constexpr int NW = DUAL ? 2 : 1;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const unsigned nk = K / 32, U = (unsigned)(N / kSlabCols) * nk, nb = gridDim.x;
    const int u_begin = (int)((blockIdx.x * U) / nb), u_end = (int)(((blockIdx.x + 1) * U) / nb);
    const int gs_steps = p.groupsize >> 5;
    const MatDesc* wd[2] = {&w0, &w1};

#ifdef GPTQ_TRACE
    long long wait_cycles = 0;
    const long long tm0 = clock64();
#endif
    int u = u_begin;
#pragma unroll 1
    while (u < u_end) {
        const int slab = u / nk;
        const int ks0 = u - slab * nk;
        const int nsteps = min((int)nk - ks0, u_end - u);
        const int col = slab * kSlabCols + warp * 32 + 4 * g;
        const int zshift = (col & 4) * 4;

        GroupRaw raw[NW];
        GroupConst gc[NW];
        const __half* scp[NW];
        const uint32_t* qzp[NW];
        {
            const int grp0 = (ks0 * 32) / p.groupsize;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                scp[w] = wd[w]->sc + (size_t)grp0 * N + col;
                qzp[w] = wd[w]->qz + (size_t)grp0 * (N >> 3) + (col >> 3);
                raw[w] = load_group_raw(scp[w], qzp[w]);
                scp[w] += N;
                qzp[w] += N >> 3;
            }
        }

        uint32_t xaddr;
        if constexpr (XMODE == X_FULL) {
            xaddr = smem_u32(xs) + (ks0 * 32 + t * 8) * 2;
        } else {
            // stage this segment's k-range [ks0*32, (ks0+nsteps)*32) of the op input
            cta_sync();  // previous readers of xs are done
            const int kbeg = ks0 * 32;
            if constexpr (XMODE == X_SWIGLU) {  // h = fp16(silu(acc_gate) * acc_up)  (quant/fused_mlp.py:163-165)
                for (int c = tid; c < nsteps * 4; c += kThreads) {
                    const int k = kbeg + c * 8;
                    uint32_t o[4];
                    const float4 g0 = *reinterpret_cast<const float4*>(p.acc_g + k), g1 = *reinterpret_cast<const float4*>(p.acc_g + k + 4);
                    const float4 u0 = *reinterpret_cast<const float4*>(p.acc_u + k), u1 = *reinterpret_cast<const float4*>(p.acc_u + k + 4);
                    o[0] = h2_as_u32(__floats2half2_rn(swiglu(g0.x, u0.x), swiglu(g0.y, u0.y)));
                    o[1] = h2_as_u32(__floats2half2_rn(swiglu(g0.z, u0.z), swiglu(g0.w, u0.w)));
                    o[2] = h2_as_u32(__floats2half2_rn(swiglu(g1.x, u1.x), swiglu(g1.y, u1.y)));
                    o[3] = h2_as_u32(__floats2half2_rn(swiglu(g1.z, u1.z), swiglu(g1.w, u1.w)));
                    store_perm8(xs + c * 8, o[0], o[1], o[2], o[3]);
                }
            } else {  // attention output: one thread per feature combines the split-KV partials of its head (coalesced over d)
                const int nvalid = min(p.nsplit, p.positions[0] / kAttnChunk + 1);
                for (int e = tid; e < nsteps * 32; e += kThreads) {
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
                    xs[(e & ~7) + ((j8 & 3) << 1) + (j8 >> 2)] = __float2half_rn(O / L);  // k-permuted position inside the run of 8
                }
            }
            cta_sync();
            xaddr = smem_u32(xs) + (t * 8) * 2;
        }

        float acc[NW][2][4];
#pragma unroll
        for (int w = 0; w < NW; ++w)
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[w][h][i] = 0.f;

#pragma unroll
        for (int w = 0; w < NW; ++w) build_group_const(gc[w], raw[w], zshift);
        int steps_left_in_grp = gs_steps - (ks0 % gs_steps);
        if (steps_left_in_grp < nsteps) {
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                raw[w] = load_group_raw(scp[w], qzp[w]);
                scp[w] += N;
                qzp[w] += N >> 3;
            }
        }

        // software pipeline: the NEXT tile's wait + shared-memory load are issued before the current tile's math, so the
        // ~120 cycles of mbarrier-probe + LDS latency overlap with ~80 instructions of dequant/MMA
        auto fetch = [&](uint4& q, uint32_t& bar_of_q) {
#ifdef GPTQ_TRACE
            const long long tw0 = clock64();
#endif
            mbar_wait(pipe.bar, pipe.parity);  // the tile has landed
#ifdef GPTQ_TRACE
            wait_cycles += clock64() - tw0;
#endif
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
#pragma unroll kStepUnroll
        for (int step = 0; step < nsteps; ++step) {
            if (steps_left_in_grp == 0) {
#pragma unroll
                for (int w = 0; w < NW; ++w) build_group_const(gc[w], raw[w], zshift);
                steps_left_in_grp = gs_steps;
                if (step + gs_steps < nsteps) {
#pragma unroll
                    for (int w = 0; w < NW; ++w) {
                        raw[w] = load_group_raw(scp[w], qzp[w]);
                        scp[w] += N;
                        qzp[w] += N >> 3;
                    }
                }
            }
            --steps_left_in_grp;
            const uint4 xf = lds128(xaddr);
            xaddr += 64;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                uint4 q_next = q_cur;
                uint32_t bar_next = bar_cur;
                if (w + 1 < NW || step + 1 < nsteps) fetch(q_next, bar_next);  // prefetch the following tile of this segment
                uint32_t wf[4][4];
                dequant8<0>(q_cur.x, gc[w].za01, gc[w].zb01, gc[w].s01, wf[0]);
                dequant8<1>(q_cur.y, gc[w].za01, gc[w].zb01, gc[w].s01, wf[1]);
                dequant8<0>(q_cur.z, gc[w].za23, gc[w].zb23, gc[w].s23, wf[2]);
                dequant8<1>(q_cur.w, gc[w].za23, gc[w].zb23, gc[w].s23, wf[3]);
                mma_16816(acc[w][0], wf[0][0], wf[1][0], wf[0][1], wf[1][1], xf.x, xf.y);
                mma_16816(acc[w][0], wf[0][2], wf[1][2], wf[0][3], wf[1][3], xf.z, xf.w);
                mma_16816(acc[w][1], wf[2][0], wf[3][0], wf[2][1], wf[3][1], xf.x, xf.y);
                mma_16816(acc[w][1], wf[2][2], wf[3][2], wf[2][3], wf[3][3], xf.z, xf.w);
                __syncwarp();  // every lane has consumed the registers it read from the current tile's stage
                if (lane == 0) mbar_arrive(bar_cur + kStages * 8);  // empty[stage]
                q_cur = q_next;
                bar_cur = bar_next;
            }
        }

        // batch row 0 lives in the t == 0 lanes: acc[.][0][0] -> col, [0][2] -> col+1, [1][0] -> col+2, [1][2] -> col+3
        if (t == 0) {
            float* o0 = out0 + col;
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o0), "f"(acc[0][0][0]), "f"(acc[0][0][2]), "f"(acc[0][1][0]), "f"(acc[0][1][2]) : "memory");
            if constexpr (DUAL) {
                float* o1 = out1 + col;
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o1), "f"(acc[1][0][0]), "f"(acc[1][0][2]), "f"(acc[1][1][0]), "f"(acc[1][1][2])
                             : "memory");
            }
        }
        u += nsteps;
    }
#ifdef GPTQ_TRACE
    if (g_mega_trace != nullptr && threadIdx.x == 0) {
        const int base = DUAL ? 40 : (XMODE == X_FULL ? 44 : (XMODE == X_ATTN ? 48 : 52));
        g_mega_trace[blockIdx.x * 64 + base] = (unsigned long long)(clock64() - tm0);
        g_mega_trace[blockIdx.x * 64 + base + 1] = (unsigned long long)wait_cycles;
        g_mega_trace[blockIdx.x * 64 + base + 2] = (unsigned long long)(u_end - u_begin);
    }
#endif
}

// Attention work items (head, split): RoPE(q,k) from this step's cos/sin, KV append, partial softmax(qK^T)V.
__device__ void run_attention(const MegaParams& p, int layer, float* smem_f) {
    const int tid = threadIdx.x;
    const int pos = p.positions[0];
    const int T = pos + 1;
    float* q_s = smem_f;                     // [128]
    float* red_m = smem_f + 128;             // [32]
    float* red_l = smem_f + 160;             // [32]
    float* red_o = smem_f + 192;             // [32][132]
    __half* kc_base = p.k_cache + layer * p.layer_stride;
    __half* vc_base = p.v_cache + layer * p.layer_stride;
    const int n_items = p.n_heads * p.nsplit;
#pragma unroll 1
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int head = item / p.nsplit, split = item - head * p.nsplit;
        const int c0 = split * kAttnChunk;
        if (c0 >= T) continue;
        const int c1 = min(c0 + kAttnChunk, T);
        __half* kc = kc_base + (size_t)head * p.max_seq * kHD;
        __half* vc = vc_base + (size_t)head * p.max_seq * kHD;
        cta_sync();  // smem reuse across items
        if (tid < kHD) {
            const int i = tid & 63;
            const bool hi = tid >= 64;
            const float c = p.rope_cs[i], s = p.rope_cs[64 + i];
            const float* aq = p.acc_qkv + head * kHD;
            const float qx = __half2float(__float2half_rn(aq[i])), qy = __half2float(__float2half_rn(aq[i + 64]));  // the qkv projection output is fp16
            const float qr = hi ? __fadd_rn(__fmul_rn(qx, s), __fmul_rn(qy, c)) : __fsub_rn(__fmul_rn(qx, c), __fmul_rn(qy, s));
            q_s[tid] = __half2float(__float2half_rn(qr));
            if (pos >= c0 && pos < c1) {  // this item owns the new key/value: append them
                const float* ak = aq + p.H;
                const float* av = aq + 2 * p.H;
                const float kx = __half2float(__float2half_rn(ak[i])), ky = __half2float(__float2half_rn(ak[i + 64]));
                const float kr = hi ? __fadd_rn(__fmul_rn(kx, s), __fmul_rn(ky, c)) : __fsub_rn(__fmul_rn(kx, c), __fmul_rn(ky, s));
                kc[(size_t)pos * kHD + tid] = __float2half_rn(kr);
                vc[(size_t)pos * kHD + tid] = __float2half_rn(av[tid]);
            }
        }
        cta_sync();
        const int grp = tid >> 3, j = tid & 7;  // 32 groups of 8 lanes; lane j owns dims [16j, 16j+16)
        constexpr int ITER = kAttnPass / 32;
        float qr[16];
#pragma unroll
        for (int d = 0; d < 16; ++d) qr[d] = q_s[16 * j + d];
        // running online-softmax state of this lane group over the item's passes
        float mloc = -INFINITY, lloc = 0.f, o[16];
#pragma unroll
        for (int d = 0; d < 16; ++d) o[d] = 0.f;
#pragma unroll 1
        for (int p0 = c0; p0 < c1; p0 += kAttnPass) {
            // all K and V rows of this lane for the pass are requested up front (clamped indices): two DRAM round trips per pass
            uint4 kreg[ITER][2], vreg[ITER][2];
#pragma unroll
            for (int it = 0; it < ITER; ++it) {
                const int tk = min(p0 + grp + it * 32, c1 - 1);
                const uint4* kp = reinterpret_cast<const uint4*>(kc + (size_t)tk * kHD + 16 * j);
                kreg[it][0] = kp[0];
                kreg[it][1] = kp[1];
            }
#pragma unroll
            for (int it = 0; it < ITER; ++it) {
                const int tk = min(p0 + grp + it * 32, c1 - 1);
                const uint4* vp = reinterpret_cast<const uint4*>(vc + (size_t)tk * kHD + 16 * j);
                vreg[it][0] = vp[0];
                vreg[it][1] = vp[1];
            }
            float sc[ITER];
            float mpass = -INFINITY;
#pragma unroll
            for (int it = 0; it < ITER; ++it) {
                const int tk = p0 + grp + it * 32;
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
                for (int it = 0; it < ITER; ++it) {
                    const int tk = p0 + grp + it * 32;
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
            red_m[grp] = mloc;
            red_l[grp] = lloc;
        }
#pragma unroll
        for (int d = 0; d < 16; ++d) red_o[grp * 132 + 16 * j + d] = o[d];
        cta_sync();
        if (tid < kHD) {
            float M = -INFINITY;
#pragma unroll 8
            for (int gI = 0; gI < 32; ++gI) M = fmaxf(M, red_m[gI]);
            float L = 0.f, O = 0.f;
#pragma unroll 8
            for (int gI = 0; gI < 32; ++gI) {
                const float wgt = (red_m[gI] == -INFINITY) ? 0.f : expf(red_m[gI] - M);
                L = fmaf(red_l[gI], wgt, L);
                O = fmaf(red_o[gI * 132 + tid], wgt, O);
            }
            float* dst = p.part + ((size_t)head * p.nsplit + split) * kRec;
            dst[4 + tid] = O;
            if (tid == 0) {
                dst[0] = M;
                dst[1] = L;
            }
        }
    }
}

template <bool ACT>
__global__ void __launch_bounds__(kBlock, 2) llama_decode_mega_kernel(const __grid_constant__ MegaParams p) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ float red_s[kWarps];
    __shared__ __align__(8) unsigned long long bars_s[2 * kStages];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // smem: [ring: 15 x 4224 B][xs: H halves][tmp: H halves / attention scratch]
    constexpr int kRingBytes = ((kStages * kTile + 127) / 128) * 128;
    __half* xs = reinterpret_cast<__half*>(smem_raw + kRingBytes);
    __half* tmp = xs + p.H;
    Pipe pipe;
    pipe.ring = smem_u32(smem_raw);
    pipe.full = smem_u32(&bars_s[0]);
    pipe.empty = smem_u32(&bars_s[kStages]);
    pipe.tile = pipe.ring + (lane & 3) * kRowPitch + (warp & 7) * 128 + (lane >> 2) * 16;  // row t of the tile, this lane's 4 columns
    pipe.bar = pipe.full;
    pipe.parity = 0;
    pipe.left = kStages;
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(pipe.full + s * 8, 1);
            mbar_init(pipe.empty + s * 8, kWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();  // the only block-wide barrier: after it the producer warp and the consumers never meet again
    if (warp >= kWarps) {
        if (lane == 0) producer_loop(p, pipe.ring, pipe.full, pipe.empty, warp - kWarps);
        return;
    }
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

    // residual stream: every stage_norm reads buffer `cur` (or the embedding row) and writes the other one
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
        run_matvec<false, X_FULL>(p, pipe, L.qkv, L.qkv, p.H, 3 * p.H, p.acc_qkv, nullptr, xs);
        MTRACE(l * 12 + 2);
        grid_barrier(p.bar, gen);
        MTRACE(l * 12 + 3);
        // ---- A ----
        zero_slice(p.acc_d, p.H);  // last read by this layer's Q
        run_attention(p, l, reinterpret_cast<float*>(tmp));
        MTRACE(l * 12 + 4);
        grid_barrier(p.bar, gen);
        MTRACE(l * 12 + 5);
        // ---- O ----
        zero_slice(p.acc_qkv, 3 * p.H);
        run_matvec<false, X_ATTN, ACT>(p, pipe, L.o, L.o, p.H, p.H, p.acc_o, nullptr, xs, L.o_perm);
        MTRACE(l * 12 + 6);
        grid_barrier(p.bar, gen);
        MTRACE(l * 12 + 7);
        // ---- G ----
        stage_norm<ACT>(p, p.resid[cur], p.acc_o, L.post_norm, p.resid[cur ^ 1], xs, tmp, red_s, L.mlp_perm);
        cur ^= 1;
        cta_sync();
        MTRACE(l * 12 + 8);
        run_matvec<true, X_FULL>(p, pipe, L.gate, L.up, p.H, p.I, p.acc_g, p.acc_u, xs);
        MTRACE(l * 12 + 9);
        grid_barrier(p.bar, gen);
        // ---- D ----
        zero_slice(p.acc_o, p.H);
        MTRACE(l * 12 + 10);
        run_matvec<false, X_SWIGLU>(p, pipe, L.down, L.down, p.I, p.H, p.acc_d, nullptr, xs);
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
        const int chunks = p.H / 8;
#pragma unroll 1
        for (int row = blockIdx.x * kWarps + warp; row < p.V; row += gridDim.x * kWarps) {
            const uint4* wr = reinterpret_cast<const uint4*>(p.lm_head + (size_t)row * p.H);
            float a = 0.f;
#pragma unroll 4
            for (int c = lane; c < chunks; c += 32) {
                uint4 wv;
                asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(wv.x), "=r"(wv.y), "=r"(wv.z), "=r"(wv.w) : "l"(wr + c));
                // xs is k-permuted: chunk c holds (k0,k4)(k1,k5)(k2,k6)(k3,k7)
                const uint4 xv = *reinterpret_cast<const uint4*>(xs + c * 8);
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
        for (int i = tid; i < p.V; i += kThreads) {
            const float v = __half2float(p.logits[i]);
            if (v > best || (v == best && i < idx)) {
                best = v;
                idx = i;
            }
        }
        __shared__ float sv[kWarps];
        __shared__ int si[kWarps];
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
            sv[warp] = best;
            si[warp] = idx;
        }
        cta_sync();
        if (tid == 0) {
            for (int w = 1; w < kWarps; ++w)
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

// ---------------------------------------------------------------------------------------------------
bool mega_supported(const gptq_llama_model& m, const gptq_llama_state& st) {
    if (getenv("GPTQ_NO_MEGA") != nullptr) return false;
    if (st.batch != 1 || m.n_layers > kMaxLayers || m.head_dim != kHD) return false;
    if (m.hidden % kSlabCols || m.intermediate % kSlabCols || m.hidden > 8192 || m.intermediate > 28672) return false;
    const int gs = m.layers[0].qkv.groupsize;
    if (gs <= 0 || gs % 32) return false;
    for (int l = 0; l < m.n_layers; ++l) {
        const gptq_llama_layer& ly = m.layers[l];
        const gptq_qweight* ws[5] = {&ly.qkv, &ly.o, &ly.gate, &ly.up, &ly.down};
        if ((ly.qkv_perm != nullptr || ly.o_perm != nullptr || ly.mlp_perm != nullptr) && (m.hidden % 8 != 0)) return false;
        for (const gptq_qweight* w : ws) {
            if (w->bits != 4 || w->groupsize != gs) return false;
            if ((reinterpret_cast<uintptr_t>(w->qweight) & 15) || (reinterpret_cast<uintptr_t>(w->scales) & 7)) return false;
        }
    }
    return true;
}

size_t mega_scratch_bytes(const gptq_llama_model& m, int max_seq) {
    const int nsplit = ceil_div(max_seq, kAttnChunk);
    return al256((size_t)m.hidden * 2) * 2 + al256((size_t)3 * m.hidden * 4) + al256((size_t)m.hidden * 4) * 2 + al256((size_t)m.intermediate * 4) * 2 +
           al256((size_t)m.n_heads * nsplit * kRec * 4) + al256(128 * 4) + 256;
}

cudaError_t launch_decode_mega(const gptq_llama_model& m, const gptq_llama_state& st, uint8_t* scratch, cudaStream_t stream) {
    static_assert(sizeof(MegaParams) < 32000, "kernel parameter space");
    MegaParams p{};
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
    bool act = false;  // any act-order gather: the ACT instantiation (the plain one carries no trace of the feature)
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
    // smem: rings + xs (max(H, widest staged segment)) + tmp (H halves or the attention scratch)
    const size_t xs_halves = (size_t)m.hidden;  // segments of the down projection are far shorter than H (checked below)
    const size_t tmp_bytes = max((size_t)m.hidden * 2, (size_t)(192 + 32 * 132) * 4);
    const size_t smem = (size_t)(((kStages * kTile + 127) / 128) * 128) + xs_halves * 2 + tmp_bytes;
    // cooperative launch: every CTA must be co-resident (2 per SM on B200); if the device cannot host them, the caller
    // falls back to the kernel-chain engine instead of risking a barrier deadlock
    int dev = 0, sms = 0, occ = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return cudaErrorInvalidDevice;
    auto kernel = act ? llama_decode_mega_kernel<true> : llama_decode_mega_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kBlock, smem) != cudaSuccess || occ < 1) return cudaErrorCooperativeLaunchTooLarge;
    const int grid = (occ >= 2 ? 2 : 1) * sms;
    // the per-CTA k-segment of the down projection must fit in xs
    const long long seg_steps = ((long long)(m.hidden / kSlabCols) * (m.intermediate / 32) + grid - 1) / grid;
    if (seg_steps * 32 > (long long)xs_halves || smem > 110 * 1024) return cudaErrorInvalidConfiguration;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
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
extern "C" int gptq_debug_set_mega_trace(void* buf) {
    unsigned long long* b = reinterpret_cast<unsigned long long*>(buf);
    return (int)cudaMemcpyToSymbol(gptq::g_mega_trace, &b, sizeof(b));
}
#endif
