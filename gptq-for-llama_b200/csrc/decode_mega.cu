// Persistent decode step: ONE cooperative kernel per token (batch 1, int4 kernel-form layers; optionally one tensor-parallel shard).
//
// One CTA per SM (148 on B200), each with two consumer TEAMS of 8 warps and one producer warp per team.
// Everything a token reads from HBM -- packed weights with their scales/zeros, the KV cache, the fp16 lm_head --
// is streamed by the producer warps through the TMA unit (cp.async.bulk.tensor / cp.async.bulk) into a per-team ring of 17 KB stages in
// shared memory, in one fixed order per team for the whole token, so the producers run ahead across operations
// and across grid barriers: while the consumers synchronise, the next operation's bytes are already landing.
//
//   per layer:  Q  qkv matvec      x = rmsnorm(resid [+ fp16(acc_down)])           -> RED acc_qkv
//               A  attention       q,k,v = fp16(acc_qkv); RoPE; KV append; a head's 32-key units dealt to its teams -> part
//               O  o_proj matvec   x = combine(part)                                -> RED acc_o
//               G  gate|up matvec  x = rmsnorm(resid + fp16(acc_o))                 -> RED acc_gate, acc_up
//               D  down matvec     x = fp16(silu(acc_gate) * acc_up)                -> RED acc_down
//   then        L  lm_head         x = rmsnorm(resid + fp16(acc_down)), fp16 rows   -> logits;   argmax
//
// Matvec stage = up to 4 k-steps (16 packed rows = 128 k) x 256 columns of ONE quantisation group, plus that
// group's 256 scales and 256 zeros.  The weights of a stage arrive as ONE cp.async.bulk.tensor request (3-D view of the
// packed matrix: 128-byte column chunk x packed row x chunk index, box 32 x 16 x 8, SWIZZLE_128B so that the four rows a
// quarter-warp reads land on distinct banks); one tensor map per row stride serves every layer (the matrices of a
// class are addressed through the chunk coordinate).  Arithmetic (quant/quant_linear.py:113-133 regrouped): inside a group
//      sum_k x_k (w_k - z) s  =  s * ( sum_k x_k w_k  -  z * sum_k x_k )
// so the consumers feed the RAW nibbles to the tensor pipe (mma.sync m16n8k16, operands swapped: A = 16 output
// columns x 16 k of weights, B = x): a nibble masked in place IS an fp16 subnormal (n * 2^-24, or 16 n * 2^-24 for
// the odd nibbles, whose x is pre-scaled by 1/16 when it is staged), products and the fp32 accumulation are
// exact, and scale / zero are applied ONCE per group on the fp32 accumulator together with the group's sum of
// x (computed when x is staged).  5 integer instructions + 1 HMMA per packed word; the result differs from the
// reference only by NOT rounding every dequantised weight to fp16 (it is closer to the exact product); the
// 1e-3 tests in tests/ hold it to the oracle.  -DGPTQ_DQ_EXACT_INT selects the variant that converts nibbles to
// exact fp16 integers (magic-number trick, 4 more fp16 instructions per word) and leaves x unscaled.
//
// Split-K partial sums are accumulated with red.global.add.f32 into fp32 vectors that the NEXT operation
// rounds to fp16 exactly where the reference rounds (a QuantLinear output is fp16); each vector is re-zeroed
// one operation after its last reader.  Only the fp32 summation order of those partials is unordered.
// Tensor parallelism (gptq_llama_tp): the o_proj / down_proj sums also go to the peer GPUs' accumulators (NVLink, see MegaParams).
#include <cuda.h>
#include <cudaTypedefs.h>

#include "common.cuh"
#include "int4_core.cuh"
#include "kernels.h"

namespace gptq {
namespace {

using namespace int4;

constexpr int kTeams = 2;                             // consumer teams per CTA
constexpr int kTeamWarps = 8;
constexpr int kTeamThreads = kTeamWarps * 32;         // 256
constexpr int kConsumers = kTeams * kTeamThreads;     // 512
constexpr int kConsumerWarps = kTeams * kTeamWarps;   // 16
constexpr int kBlock = kConsumers + 32 * kTeams;      // + one producer warp per team
constexpr int kSlabCols = 256;
constexpr int kStageSteps = 4;         // k-steps (of 32 k = 4 packed rows) per stage
constexpr int kStepBytes = 4096;       // 4 packed rows x 256 columns
constexpr int kScaleOff = kStageSteps * kStepBytes;     // 16384: 256 fp16 scales of the stage's group
constexpr int kZeroOff = kScaleOff + 512;               // 16896: 32 qzeros words (256 nibbles)
constexpr int kStageBytes = 17 * 1024;                  // 17408: stages are 1 KB aligned (128-byte swizzle atoms of the TMA boxes)
constexpr int kHD = 128;
constexpr int kKeysPerUnit = 32;       // attention work unit: 32 keys of one head = 8 KB of K + 8 KB of V = one stage
constexpr int kVOff = 8192 + 256;      // V rows of a KV stage (K rows at offset 0)
constexpr int kRec = kHD + 4;          // floats per attention partial record: m, l, 2 pad, o[128]
constexpr int kMaxLayers = 80;
constexpr int kMaxStages = 8;
constexpr int kTeamScratch = 4224;     // per-team scratch (attention merge buffers / lm_head partials)
constexpr int kLmStageBytes = 16384;   // lm_head bytes per stage (whole rows)

#ifdef GPTQ_DQ_EXACT_INT
constexpr bool kSubnormal = false;
#else
constexpr bool kSubnormal = true;
#endif

#ifdef GPTQ_TRACE
}  // namespace
__device__ unsigned long long* g_mega_trace = nullptr;
namespace {

#define MTRACE(id)                                                                                                       \
    do {                                                                                                                 \
        if (g_mega_trace != nullptr && threadIdx.x == 0 && (id) < 36) {                                                   \
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

#ifdef GPTQ_TRACE
// per-op counters of layer 2, team 0 of every CTA: ids 48 + 3 * op + {0: cycles in the op, 1: of which waiting for stages, 2: stages}
#define OPTRACE_BEGIN(ring)                       \
    const long long optr_t0 = clock64();          \
    const long long optr_w0 = (ring).waited;      \
    const int optr_s0 = (ring).stages
#define OPTRACE_END(ring, layer, op)                                                                                   \
    do {                                                                                                               \
        if (g_mega_trace != nullptr && threadIdx.x == 0 && (layer) == 2) {                                              \
            g_mega_trace[blockIdx.x * 64 + 48 + 3 * (op)] = (unsigned long long)(clock64() - optr_t0);                  \
            g_mega_trace[blockIdx.x * 64 + 48 + 3 * (op) + 1] = (unsigned long long)((ring).waited - optr_w0);          \
            g_mega_trace[blockIdx.x * 64 + 48 + 3 * (op) + 2] = (unsigned long long)((ring).stages - optr_s0);          \
        }                                                                                                              \
    } while (0)
#else
#define OPTRACE_BEGIN(ring) \
    do {                    \
    } while (0)
#define OPTRACE_END(ring, layer, op) \
    do {                             \
    } while (0)
#endif

struct MatDesc {
    const __half* sc;
    const uint32_t* qz;
    int gs_steps;  // k-steps per quantisation group (groupsize / 32)
    int tmap;      // tensor-map class of qweight (0: N = 3H, 1: N = H, 2: N = I)
    int chunk0;    // chunk coordinate of the matrix: (qweight - class base) / 128 bytes
};
struct LayerDesc {
    MatDesc qkv, o, gate, up, down;
    const __half* input_norm;
    const __half* post_norm;
    const int32_t* qkv_perm;  // act-order input gathers (gptq_llama_layer), nullptr = identity
    const int32_t* o_perm;
    const int32_t* mlp_perm;
};
constexpr int kMaxTP = 8;
struct MegaParams {
    // H: hidden size (the residual stream, replicated under tensor parallelism); n_heads, Hq = 128 * n_heads, I: this rank's attention heads
    // / attention width / MLP width (the full ones on a single GPU); V: vocabulary, [v0, v1): the lm_head rows of this rank
    int n_layers, H, Hq, I, V, v0, v1, n_heads, max_seq, n_stages, lm_rows;
    int tp_size, tp_rank, tp_exchange;  // tp_exchange: reduce locally first, then hand slices to the ranks (else: every team REDs into every rank)
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
    float* part;       // [teams][kRec]  attention partial records (m, l, pad, pad, o[128]), one per team and layer
    float* rope_cs;    // [128]: cos[64], sin[64] of this step's position
    unsigned long long* bar;  // [0]: monotonic arrival counter of the grid barrier, [1]: its value when the previous launch ended
    // tensor parallelism (tp_size > 1): a rank first reduces its o_proj / down_proj partial sums locally (acc_*_loc), then every CTA adds its
    // slice of that vector into EVERY rank's accumulator over NVLink (peer pointers, red.sys: 4 bytes x H x ranks per hand-off instead of
    // every team's partials); the lm_head slices are stored into every rank's logits; the hand-offs that follow use a cross-GPU barrier:
    // xbar = this rank's arrival counters, one 256-byte line per source rank, line kMaxTP = value at launch end
    float* acc_o_loc;  // [H] this rank's partial sums of o_proj / down_proj before they are handed to the ranks (tp_size > 1)
    float* acc_d_loc;
    float* acc_o_peer[kMaxTP];
    float* acc_d_peer[kMaxTP];
    __half* logits_peer[kMaxTP];
    unsigned long long* xbar_peer[kMaxTP];
    LayerDesc layers[kMaxLayers];
    CUtensorMap tmaps[6];  // [class]: 16-row boxes (a full stage), [3 + class]: 4-row boxes (one k-step)
};

// ---- work split ------------------------------------------------------------------------------------------------
// U units of an operation are dealt to the nb teams as contiguous ranges [T*U/nb, (T+1)*U/nb).
__device__ __forceinline__ void team_range(unsigned T, unsigned U, unsigned nb, int& a, int& b) {
    a = (int)((T * U) / nb);  // T * U < 2^32 (checked by mega_plan)
    b = (int)(((T + 1) * U) / nb);
}
// Attention: the teams are dealt to the heads (nb / n_heads or one more each), a head's 32-key units to its teams: no team meets two
// heads, so every team writes exactly one partial record per layer (a neutral one if it has no unit).
struct HeadTeams {
    int first, count;  // teams [first, first + count) serve the head
};
__device__ __forceinline__ HeadTeams head_teams(int head, int n_heads, unsigned nb) {
    const int base = (int)nb / n_heads, rem = (int)nb % n_heads;
    HeadTeams h;
    h.first = head * base + min(head, rem);
    h.count = base + (head < rem ? 1 : 0);
    return h;
}
// (head, unit range [b0, b1) inside the head) of team T for a context of upb units per head
__device__ __forceinline__ void attn_range(unsigned T, int n_heads, unsigned nb, int upb, int& head, int& b0, int& b1) {
    const int base = (int)nb / n_heads, rem = (int)nb % n_heads;
    const int split = rem * (base + 1);
    int idx, cnt;
    if ((int)T < split) {
        head = (int)T / (base + 1);
        idx = (int)T - head * (base + 1);
        cnt = base + 1;
    } else {
        head = rem + ((int)T - split) / base;
        idx = ((int)T - split) % base;
        cnt = base;
    }
    b0 = idx * upb / cnt;
    b1 = (idx + 1) * upb / cnt;
}

// k-steps of the stage that starts `gpos` steps into its quantisation group, in a segment with `left` steps to go
__device__ __forceinline__ int stage_steps(int gpos, int gs_steps, int left) { return min(min(kStageSteps, gs_steps - gpos), left); }

// position of this step, clamped to the cache (the host rejects pos >= max_seq; the kernel must not write outside)
__device__ __forceinline__ int step_pos(const MegaParams& p) { return min(max(p.positions[0], 0), p.max_seq - 1); }

// ---- team / CTA synchronisation ------------------------------------------------------------------------------------
__device__ __forceinline__ void team_sync(int team) { asm volatile("bar.sync %0, 256;" ::"r"(team + 1) : "memory"); }
__device__ __forceinline__ void cta_sync() { asm volatile("bar.sync 3, 512;" ::: "memory"); }  // all consumer warps (not the producers)

// ---- producer ------------------------------------------------------------------------------------------------------
struct ProdRing {
    uint32_t ring, full, empty;
    int stage, use, nstages;
#ifdef GPTQ_TRACE
    long long blocked;  // cycles spent waiting for a free stage
#endif
};
__device__ __forceinline__ uint32_t prod_acquire(ProdRing& r, uint32_t& bar) {
#ifdef GPTQ_TRACE
    const long long t0 = clock64();
#endif
    if (r.use > 0) mbar_wait_backoff(r.empty + r.stage * 8, (r.use - 1) & 1u);  // the consumers released the previous use of this stage
#ifdef GPTQ_TRACE
    r.blocked += clock64() - t0;
#endif
    bar = r.full + r.stage * 8;
    return r.ring + r.stage * kStageBytes;
}
__device__ __forceinline__ void prod_advance(ProdRing& r) {
    if (++r.stage == r.nstages) {
        r.stage = 0;
        ++r.use;
    }
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst), "l"(tm), "r"(c0),
                 "r"(c1), "r"(c2), "r"(bar)
                 : "memory");
}

// weights of one matvec (NM matrices side by side: gate|up is one virtual matrix of 2 * N/256 slabs)
template <int NM>
__device__ void produce_matvec(ProdRing& r, const MegaParams& p, const MatDesc* const (&md)[NM], int K, int N, unsigned T, unsigned nb) {
    const int nk = K / 32, nslab = N / kSlabCols;
    int u, u1;
    team_range(T, (unsigned)(NM * nslab) * nk, nb, u, u1);
#pragma unroll 1
    while (u < u1) {
        const int slab_v = u / nk;
        int ks = u - slab_v * nk;
        const int nseg = min(nk - ks, u1 - u);
        const int mi = (NM > 1 && slab_v >= nslab) ? 1 : 0;
        const int slab = slab_v - mi * nslab;
        const MatDesc& m = *md[mi];
        const int chunk = m.chunk0 + slab * (kSlabCols / 32);
        int left = nseg;
        int g = ks / m.gs_steps, gpos = ks - g * m.gs_steps;  // group of step ks and the position inside it
#pragma unroll 1
        while (left > 0) {
            const int n = stage_steps(gpos, m.gs_steps, left);
            uint32_t bar;
            const uint32_t dst = prod_acquire(r, bar);
            mbar_expect_tx(bar, n * kStepBytes + 640);
            if (n == kStageSteps) {
                tma_load_3d(dst, &p.tmaps[m.tmap], 0, ks * 4, chunk, bar);
            } else {
                for (int j = 0; j < n; ++j) tma_load_3d(dst + j * kStepBytes, &p.tmaps[3 + m.tmap], 0, (ks + j) * 4, chunk, bar);
            }
            bulk_copy_g2s(dst + kScaleOff, m.sc + (size_t)g * N + slab * kSlabCols, 512, bar);
            bulk_copy_g2s(dst + kZeroOff, m.qz + (size_t)g * (N >> 3) + slab * (kSlabCols / 8), 128, bar);
            prod_advance(r);
            ks += n;
            left -= n;
            gpos += n;
            if (gpos == m.gs_steps) {
                gpos = 0;
                ++g;
            }
        }
        u += nseg;
    }
}

__device__ void produce_kv(ProdRing& r, const MegaParams& p, int layer, unsigned T, unsigned nb) {
    const int Tlen = step_pos(p) + 1;
    const int upb = (Tlen + kKeysPerUnit - 1) / kKeysPerUnit;
    int head, b, b1;
    attn_range(T, p.n_heads, nb, upb, head, b, b1);
    const __half* kc = p.k_cache + layer * p.layer_stride;
    const __half* vc = p.v_cache + layer * p.layer_stride;
#pragma unroll 1
    for (; b < b1; ++b) {
        const int nrows = min(kKeysPerUnit, Tlen - b * kKeysPerUnit);
        const size_t off = ((size_t)head * p.max_seq + (size_t)b * kKeysPerUnit) * kHD;
        uint32_t bar;
        const uint32_t dst = prod_acquire(r, bar);
        mbar_expect_tx(bar, nrows * 512);
        bulk_copy_g2s(dst, kc + off, nrows * 256, bar);
        bulk_copy_g2s(dst + kVOff, vc + off, nrows * 256, bar);
        prod_advance(r);
    }
}

__device__ void produce_lm_head(ProdRing& r, const MegaParams& p, unsigned T, unsigned nb) {
    const int R = p.lm_rows;
    int u, u1;
    const int nloc = p.v1 - p.v0;
    team_range(T, (unsigned)((nloc + R - 1) / R), nb, u, u1);
#pragma unroll 1
    for (; u < u1; ++u) {
        const int nrows = min(R, nloc - u * R);
        uint32_t bar;
        const uint32_t dst = prod_acquire(r, bar);
        const uint32_t bytes = (uint32_t)nrows * p.H * 2;
        mbar_expect_tx(bar, bytes);
        bulk_copy_g2s(dst, p.lm_head + (size_t)u * R * p.H, bytes, bar);
        prod_advance(r);
    }
}

__device__ void producer_loop(const MegaParams& p, ProdRing r, unsigned T, unsigned nb) {
#pragma unroll 1
    for (int l = 0; l < p.n_layers; ++l) {
        const LayerDesc& L = p.layers[l];
        {
            const MatDesc* const md[1] = {&L.qkv};
            produce_matvec<1>(r, p, md, p.H, 3 * p.Hq, T, nb);
        }
        produce_kv(r, p, l, T, nb);
        {
            const MatDesc* const md[1] = {&L.o};
            produce_matvec<1>(r, p, md, p.Hq, p.H, T, nb);
        }
        {
            const MatDesc* const md[2] = {&L.gate, &L.up};
            produce_matvec<2>(r, p, md, p.H, p.I, T, nb);
        }
        {
            const MatDesc* const md[1] = {&L.down};
            produce_matvec<1>(r, p, md, p.I, p.H, T, nb);
        }
    }
    produce_lm_head(r, p, T, nb);
#ifdef GPTQ_TRACE
    if (g_mega_trace != nullptr && (T & 1) == 0) g_mega_trace[(T / kTeams) * 64 + 63] = (unsigned long long)r.blocked;
#endif
}

// ---- consumer side of the ring -----------------------------------------------------------------------------------------
struct ConsRing {
    uint32_t ring, full;  // stage 0 / full[0]; empty[s] sits kMaxStages * 8 bytes after full[s]
    uint32_t tile, bar;   // current stage
    uint32_t parity;
    int left, nstages;
#ifdef GPTQ_TRACE
    long long waited;  // cycles spent waiting for stages to land
    int stages;
#endif
};
__device__ __forceinline__ uint32_t cons_wait(ConsRing& c) {
#ifdef GPTQ_TRACE
    const long long t0 = clock64();
#endif
    mbar_wait(c.bar, c.parity);
#ifdef GPTQ_TRACE
    c.waited += clock64() - t0;
    ++c.stages;
#endif
    return c.tile;
}
// every lane of the warp has finished reading the stage
__device__ __forceinline__ void cons_release(ConsRing& c, int lane) {
    __syncwarp();
    if (lane == 0) mbar_arrive(c.bar + kMaxStages * 8);
    c.tile += kStageBytes;
    c.bar += 8;
    if (--c.left == 0) {
        c.left = c.nstages;
        c.tile = c.ring;
        c.bar = c.full;
        c.parity ^= 1u;
    }
}

// ---- grid barrier: one monotonic 64-bit arrival counter ----------------------------------------------------------------
// arrive = red.release (fire and forget), wait = poll the same word with ld.acquire until it reaches this barrier's
// target.  `target` lives in thread 0 of each CTA.
__device__ __forceinline__ void grid_barrier(unsigned long long* bar, unsigned long long& target) {
    cta_sync();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(bar) : "memory");
        unsigned long long v;
        do {
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(bar) : "memory");
        } while (v < target);
#ifdef GPTQ_BARRIER_FENCE
        fence_acq_rel_gpu();
#endif
    }
    cta_sync();
}

// Cross-GPU barrier of all CTAs of all tensor-parallel ranks: every CTA adds 1 to ITS source line of every rank's counter block and polls its
// own rank's lines.  ONE system-scope fence per CTA makes its remote REDs (partial sums in the peers' accumulators) visible before its
// arrivals; the arrivals themselves and the polls are relaxed (the counters and the data they guard live in the polling GPU's own L2, and
// everything that is read afterwards bypasses L1).
__device__ __forceinline__ void tp_barrier(const MegaParams& p, unsigned long long& target) {
    cta_sync();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        asm volatile("fence.acq_rel.sys;" ::: "memory");
        for (int q = 0; q < p.tp_size; ++q)
            asm volatile("red.relaxed.sys.global.add.u64 [%0], 1;" ::"l"(p.xbar_peer[q] + 32 * p.tp_rank) : "memory");
        const unsigned long long* mine = p.xbar_peer[p.tp_rank];
        for (int q = 0; q < p.tp_size; ++q) {
            unsigned long long v;
            do {
                asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine + 32 * q) : "memory");
            } while (v < target);
        }
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    cta_sync();
}

// This rank's locally reduced vector (n floats, complete after a grid barrier) is added into every rank's accumulator and cleared for its next use.
__device__ __forceinline__ void tp_exchange(const MegaParams& p, float* loc, float* const* peers, int n) {
    const int per = (n + gridDim.x - 1) / gridDim.x;
    const int lo = blockIdx.x * per, hi = min(n, lo + per);
    for (int j = lo + threadIdx.x; j < hi; j += kConsumers) {
        const float v = ld_cg(loc + j);
        loc[j] = 0.f;
        for (int q = 0; q < p.tp_size; ++q) asm volatile("red.relaxed.sys.global.add.f32 [%0], %1;" ::"l"(peers[q] + j), "f"(v) : "memory");
    }
}

__device__ __forceinline__ void zero_slice(float* buf, int n) {
    // this CTA's share of a distributed memset (n is a multiple of 4)
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
    for (int w = 0; w < kConsumerWarps; ++w) t += red_s[w];
    return t;
}

// ---- x staging ----------------------------------------------------------------------------------------------------------
// Matvec input layout: 8 consecutive k (natural order, 4 half2 words w0..w3) are stored k-permuted as
// (k0,k4)(k1,k5)(k2,k6)(k3,k7): the B fragments of the two MMAs of a packed word; the pairs that meet the ODD
// nibbles (k1,k5 / k3,k7; mask 0x00f000f0 = 16 n) are pre-scaled by 1/16 (kSubnormal only).
__device__ __forceinline__ uint4 perm8(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
    uint4 o;
    o.x = __byte_perm(w0, w2, 0x5410);
    o.y = __byte_perm(w0, w2, 0x7632);
    o.z = __byte_perm(w1, w3, 0x5410);
    o.w = __byte_perm(w1, w3, 0x7632);
    if (kSubnormal) {
        const __half2 sixteenth = __float2half2_rn(0.0625f);
        o.y = h2_as_u32(__hmul2(u32_as_h2(o.y), sixteenth));
        o.w = h2_as_u32(__hmul2(u32_as_h2(o.w), sixteenth));
    }
    return o;
}
// position of natural index j (0..7) inside the permuted run of 8, and whether it is pre-scaled
__device__ __forceinline__ int perm_pos(int j) { return ((j & 3) << 1) + (j >> 2); }
__device__ __forceinline__ bool perm_scaled(int j) { return kSubnormal && (j & 1); }

// sum of the EFFECTIVE x of one staged run of 8 (what the tensor pipe will multiply the nibbles with)
__device__ __forceinline__ float run_sum(uint4 v) {
    const float2 a = __half22float2(u32_as_h2(v.x)), b = __half22float2(u32_as_h2(v.y)), c = __half22float2(u32_as_h2(v.z)), d = __half22float2(u32_as_h2(v.w));
    const float even = (a.x + a.y) + (c.x + c.y), odd = (b.x + b.y) + (d.x + d.y);
    return kSubnormal ? fmaf(odd, 16.0f, even) : even + odd;
}
// xsum[s] = sum over the 32 k of step s of the staged x (4 runs of 8), for s < nsteps; `nthreads` threads (a multiple of 32) cooperate
__device__ __forceinline__ void compute_xsum(const __half* xs, int nsteps, float* xsum, int tid, int nthreads) {
    const int n4 = nsteps * 4;
    for (int base = 0; base < n4; base += nthreads) {
        const int idx = base + tid;
        float v = 0.f;
        if (idx < n4) v = run_sum(*reinterpret_cast<const uint4*>(xs + idx * 8));
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        if (idx < n4 && (idx & 3) == 0) xsum[idx >> 2] = v;
    }
}

// x = rmsnorm(src [+ fp16(acc)]) for the whole row (K = H), staged in xs (matvec layout, or natural order if PLAIN)
// with its per-step sums in xsum; the updated residual stream (src + fp16(acc)) is written to resid_out by slices.
// All 512 consumer threads.  ACT: the matvec's packed rows were regrouped by the host: position k' holds feature perm[k'].
template <bool ACT, bool PLAIN>
__device__ void stage_norm(const MegaParams& p, const __half* src, const float* acc, const __half* norm_w, __half* resid_out, __half* xs, float* xsum,
                           __half* tmp, float* red_s, const int32_t* perm) {
    const int H = p.H, tid = threadIdx.x, nch = H / 8;
    float ss = 0.f;
    uint4 nwv[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int c = tid + i * kConsumers;
        nwv[i] = make_uint4(0, 0, 0, 0);
        if (c < nch) {
            nwv[i] = *reinterpret_cast<const uint4*>(norm_w + c * 8);  // same round trip as src / acc
            const uint4 v = ld_cg_u4(src + c * 8);  // written by other CTAs (residual slices): L2, not this SM's L1
            uint32_t xv[4] = {v.x, v.y, v.z, v.w};
            if (acc != nullptr) {
                const float4 a0 = ld_cg4(acc + c * 8), a1 = ld_cg4(acc + c * 8 + 4);
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
    }
    const float tot = block_sum(ss, red_s);  // its barriers also publish tmp
    const float rstd = 1.0f / sqrtf(tot / (float)H + p.eps);
    if (resid_out != nullptr) {
        const int per = (nch + gridDim.x - 1) / gridDim.x;
        const int c = blockIdx.x * per + tid;
        if (tid < per && c < nch) *reinterpret_cast<uint4*>(resid_out + c * 8) = *reinterpret_cast<const uint4*>(tmp + c * 8);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int c = tid + i * kConsumers;
        if (c < nch) {
            float xf[8];
            bool gathered = false;
            if constexpr (ACT) {
                if (perm != nullptr) {  // norm_w is given in regrouped order (gptq_b200.h); the gather itself reads shared memory
                    const ::int4 p0 = *reinterpret_cast<const ::int4*>(perm + c * 8), p1 = *reinterpret_cast<const ::int4*>(perm + c * 8 + 4);
                    const int k[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
                    for (int j = 0; j < 8; ++j) xf[j] = __half2float(tmp[k[j]]);
                    gathered = true;
                }
            }
            if (!gathered) {
                const uint4 v = *reinterpret_cast<const uint4*>(tmp + c * 8);
                const uint32_t xv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = __half22float2(u32_as_h2(xv[j]));
                    xf[2 * j] = f.x;
                    xf[2 * j + 1] = f.y;
                }
            }
            const uint32_t wv[4] = {nwv[i].x, nwv[i].y, nwv[i].z, nwv[i].w};
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {  // fp32 normalise, fp32 weight multiply, one rounding to fp16 (quant/triton_norm.py:30-38)
                const float2 wf = __half22float2(u32_as_h2(wv[j]));
                o[j] = h2_as_u32(__floats2half2_rn(__fmul_rn(__fmul_rn(xf[2 * j], rstd), wf.x), __fmul_rn(__fmul_rn(xf[2 * j + 1], rstd), wf.y)));
            }
            if constexpr (PLAIN) {
                *reinterpret_cast<uint4*>(xs + c * 8) = make_uint4(o[0], o[1], o[2], o[3]);
            } else {
                const uint4 pv = perm8(o[0], o[1], o[2], o[3]);
                *reinterpret_cast<uint4*>(xs + c * 8) = pv;
                // per-step sums of the staged x: the 4 runs of a k-step sit in 4 consecutive lanes (c < nch is uniform per warp)
                float v = run_sum(pv);
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                if ((c & 3) == 0) xsum[c >> 2] = v;
            }
        }
    }
    cta_sync();
}

// ---- matvec consumer ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_16816_z(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(0.f));
}

// the four A registers (k-pairs (0,4), (1,5), (2,6), (3,7)) of one packed word
__device__ __forceinline__ void nibble_regs(uint32_t q, uint32_t (&a)[4]) {
    const uint32_t q8 = q >> 8;  // (as IMAD.HI on the FMA pipe it is far slower: tools/ubench/loop.cu, 1102 vs 817 cycles per stage)
    if (kSubnormal) {  // masked in place: fp16 subnormals n * 2^-24 and 16 n * 2^-24
        a[0] = q & 0x000f000fu;
        a[1] = q & 0x00f000f0u;
        a[2] = q8 & 0x000f000fu;
        a[3] = q8 & 0x00f000f0u;
    } else {  // exact fp16 integers: (1024 + n) - 1024 and (1024 + 16 n) / 16 - 64
        const __half2 c1024 = __float2half2_rn(1024.0f), sixteenth = __float2half2_rn(0.0625f), m64 = __float2half2_rn(-64.0f);
        a[0] = h2_as_u32(__hsub2(nibbles_to_h2<0x000f000fu>(q), c1024));
        a[1] = h2_as_u32(__hfma2(nibbles_to_h2<0x00f000f0u>(q), sixteenth, m64));
        a[2] = h2_as_u32(__hsub2(nibbles_to_h2<0x000f000fu>(q8), c1024));
        a[3] = h2_as_u32(__hfma2(nibbles_to_h2<0x00f000f0u>(q8), sixteenth, m64));
    }
}

// one k-step (4 packed words of this lane = columns 4g..4g+3 x 8 k) into the two accumulators
template <bool FIRST>
__device__ __forceinline__ void step_mma(const uint4& q, const uint4& xf, float (&acc0)[4], float (&acc1)[4]) {
    uint32_t c0[4], c1[4], c2[4], c3[4];
    nibble_regs(q.x, c0);
    nibble_regs(q.y, c1);
    nibble_regs(q.z, c2);
    nibble_regs(q.w, c3);
    if (FIRST) {
        mma_16816_z(acc0, c0[0], c1[0], c0[1], c1[1], xf.x, xf.y);
        mma_16816_z(acc1, c2[0], c3[0], c2[1], c3[1], xf.x, xf.y);
    } else {
        mma_16816(acc0, c0[0], c1[0], c0[1], c1[1], xf.x, xf.y);
        mma_16816(acc1, c2[0], c3[0], c2[1], c3[1], xf.x, xf.y);
    }
    mma_16816(acc0, c0[2], c1[2], c0[3], c1[3], xf.z, xf.w);
    mma_16816(acc1, c2[2], c3[2], c2[3], c3[3], xf.z, xf.w);
}

enum XMode { X_FULL = 0, X_ATTN = 1, X_SWIGLU = 2 };

struct TeamCtx {
    int team, wt, lane, ttid;  // team, warp in team, lane, thread in team
    unsigned T, nb;            // global team index / number of teams
    __half* xseg;              // this team's half of the xs buffer (per-segment inputs of O and D)
    float* xsum_seg;
    uint8_t* scratch;          // kTeamScratch bytes
};


// Input of o_proj (XMODE == X_ATTN) / down_proj (X_SWIGLU) for the team's WHOLE unit range [u0, u1) (units = k-steps of 32, numbered
// slab-major; the range wraps at most once from the end of one slab's k-range to the start of the next): staged once, in unit order,
// into the team's half of the xs buffer with its per-step sums.  One L2 round trip for the common shapes.
template <int XMODE, bool ACT>
__device__ void stage_range(const MegaParams& p, const TeamCtx& tc, int nk, int u0, int u1, const int32_t* perm) {
    const int lane = tc.lane;
    const int nun = u1 - u0, nfeat = nun * 32;
    const int ksb = u0 % nk;  // k-step of the first unit
    auto k_of = [&](int e) {  // feature e of the range -> input index k
        int ks = ksb + (e >> 5);
        if (ks >= nk) ks -= nk;
        return ks * 32 + (e & 31);
    };
    team_sync(tc.team);  // previous readers of xseg are done
    if constexpr (XMODE == X_SWIGLU) {  // h = fp16(silu(acc_gate) * acc_up)  (quant/fused_mlp.py:163-165)
        for (int c = tc.ttid; c < nun * 4; c += kTeamThreads) {
            const int k = k_of(c * 8);
            const float4 g0 = ld_cg4(p.acc_g + k), g1 = ld_cg4(p.acc_g + k + 4);
            const float4 a0 = ld_cg4(p.acc_u + k), a1 = ld_cg4(p.acc_u + k + 4);
            const uint32_t o0 = h2_as_u32(__floats2half2_rn(swiglu(g0.x, a0.x), swiglu(g0.y, a0.y)));
            const uint32_t o1 = h2_as_u32(__floats2half2_rn(swiglu(g0.z, a0.z), swiglu(g0.w, a0.w)));
            const uint32_t o2 = h2_as_u32(__floats2half2_rn(swiglu(g1.x, a1.x), swiglu(g1.y, a1.y)));
            const uint32_t o3 = h2_as_u32(__floats2half2_rn(swiglu(g1.z, a1.z), swiglu(g1.w, a1.w)));
            *reinterpret_cast<uint4*>(tc.xseg + c * 8) = perm8(o0, o1, o2, o3);
        }
    } else {
        // attention output: softmax-merge of the partial records (m, l, o[128]) of the head's teams (head_teams: contiguous, one
        // record each).
        constexpr int kBatch = 12;  // records fetched per round trip (a 7B head has 9 or 10 teams)
        auto put = [&](int e, float v) {
            __half hv = __float2half_rn(v);
            const int j8 = e & 7;
            if (perm_scaled(j8)) hv = __hmul(hv, __float2half_rn(0.0625f));
            tc.xseg[(e & ~7) + perm_pos(j8)] = hv;
        };
        // heads of the range: [hA0, hA1] before the wrap, [0, hB1] after it
        const int nA = min(nun, nk - ksb);
        const int hA0 = (ksb * 32) / kHD, hA1 = ((ksb + nA) * 32 - 1) / kHD;
        const int nslotA = hA1 - hA0 + 1, nslotB = (nun > nA) ? ((nun - nA) * 32 - 1) / kHD + 1 : 0;
        bool fast = (nslotA + nslotB <= kTeamWarps) && (nun <= nk) && ((int)tc.nb / p.n_heads + 1 <= kBatch);
        if constexpr (ACT) fast = fast && (perm == nullptr);
        if (fast) {
            // ONE round trip (per 256 features): every thread fetches the o values of its feature from all records of its head while
            // warp w fetches (m, l) of the w-th head of the range and turns them into merge weights exp(m - M) / L.
            float* wts = reinterpret_cast<float*>(tc.scratch);  // [kTeamWarps][kBatch]
            float ov[kBatch];
            int slot = 0;
            auto fetch = [&](int e) {
                const bool active = e < nfeat;
                const int k = k_of(active ? e : 0);
                const int head = k / kHD, d = k - head * kHD;
                slot = ((e >> 5) < nA) ? head - hA0 : nslotA + head;
                const HeadTeams ht = head_teams(head, p.n_heads, tc.nb);
#pragma unroll
                for (int i = 0; i < kBatch; ++i) ov[i] = (active && i < ht.count) ? ld_cg(p.part + (size_t)(ht.first + i) * kRec + 4 + d) : 0.f;
            };
            auto emit = [&](int e) {
                if (e < nfeat) {
                    float O = 0.f;
#pragma unroll
                    for (int i = 0; i < kBatch; ++i) O = fmaf(wts[slot * kBatch + i], ov[i], O);
                    put(e, O);
                }
            };
            fetch(tc.ttid);
            if (tc.wt < nslotA + nslotB) {
                const int head = tc.wt < nslotA ? hA0 + tc.wt : tc.wt - nslotA;
                const HeadTeams ht = head_teams(head, p.n_heads, tc.nb);
                float m = -INFINITY, l = 0.f;
                if (lane < ht.count) {
                    m = ld_cg(p.part + (size_t)(ht.first + lane) * kRec);
                    l = ld_cg(p.part + (size_t)(ht.first + lane) * kRec + 1);
                }
                float M = m;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, o));  // lanes 0..15 hold the batch
                M = __shfl_sync(0xffffffffu, M, 0);
                const float w = (m == -INFINITY) ? 0.f : expf(m - M);
                float L = l * w;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) L += __shfl_xor_sync(0xffffffffu, L, o);
                L = __shfl_sync(0xffffffffu, L, 0);
                if (lane < kBatch) wts[tc.wt * kBatch + lane] = w / L;
            }
            team_sync(tc.team);
            emit(tc.ttid);
            for (int e = tc.ttid + kTeamThreads; e - tc.ttid < nfeat; e += kTeamThreads) {  // wider ranges (13B, 65B): further round trips
                fetch(e);
                emit(e);
            }
        } else {
            for (int e = tc.ttid; e < nfeat; e += kTeamThreads) {
                int k = k_of(e);
                if constexpr (ACT) {
                    if (perm != nullptr) k = perm[k];  // regrouped rows: position k' of the matvec input is attention feature perm[k']
                }
                const int head = k / kHD, d = k - head * kHD;
                const HeadTeams ht = head_teams(head, p.n_heads, tc.nb);
                float M = -INFINITY, Ls = 0.f, O = 0.f;
#pragma unroll 1
                for (int ib = 0; ib < ht.count; ib += kBatch) {
                    float mv[kBatch], lv[kBatch], ov[kBatch];
#pragma unroll
                    for (int i = 0; i < kBatch; ++i) {  // all loads of the batch are in flight together
                        const bool on = ib + i < ht.count;
                        const float* rc = p.part + (size_t)(ht.first + (on ? ib + i : 0)) * kRec;
                        const float m = ld_cg(rc), l = ld_cg(rc + 1), o = ld_cg(rc + 4 + d);
                        mv[i] = on ? m : -INFINITY;
                        lv[i] = on ? l : 0.f;
                        ov[i] = on ? o : 0.f;
                    }
                    float Mb = M;
#pragma unroll
                    for (int i = 0; i < kBatch; ++i) Mb = fmaxf(Mb, mv[i]);
                    const float w0 = (M == -INFINITY) ? 0.f : expf(M - Mb);
                    Ls *= w0;
                    O *= w0;
#pragma unroll
                    for (int i = 0; i < kBatch; ++i) {
                        const float w = (mv[i] == -INFINITY) ? 0.f : expf(mv[i] - Mb);
                        Ls = fmaf(lv[i], w, Ls);
                        O = fmaf(ov[i], w, O);
                    }
                    M = Mb;
                }
                put(e, O / Ls);
            }
        }
    }
    team_sync(tc.team);
    compute_xsum(tc.xseg, nun, tc.xsum_seg, tc.ttid, kTeamThreads);
    team_sync(tc.team);
}

// One matvec op for this team: consume the stages of its unit range from the ring, RED the results.
// NM = 2: gate|up as one virtual matrix (out0 = gate accumulators, out1 = up accumulators).
template <int NM, int XMODE, bool ACT = false>
__device__ void run_matvec(const MegaParams& p, ConsRing& ring, const TeamCtx& tc, int gs_steps, int K, int N, float* out0, float* out1, const __half* xs_full,
                           const float* xsum_full, const int32_t* perm = nullptr, float* const* peers = nullptr) {
    const int lane = tc.lane, g = lane >> 2, t = lane & 3;
    const int nk = K / 32, nslab = N / kSlabCols;
    int u, u_end;
    team_range(tc.T, (unsigned)(NM * nslab) * nk, tc.nb, u, u_end);
    // MMA row g of this lane carries the 4 columns of 16-byte unit cg of the warp's 128-byte chunk; cg is chosen so that the
    // swizzled units (unit ^ row) of a quarter-warp (g in {2q, 2q+1}, t = 0..3) are 8 distinct bank groups
    const int cg = (g >> 1) | ((g & 1) << 2);
    const int col_l = 4 * cg + t;  // the column (of the warp's 32) this lane finishes in the epilogue
    // 16-row box [chunk][row][128 B]: row 4j + t of chunk wt, unit cg ^ ((4j + t) & 7)
    const uint32_t lane_a0 = (uint32_t)((tc.wt * 16 + t) * 128 + ((cg ^ t) * 16));       // even k-steps of the stage
    const uint32_t lane_a1 = (uint32_t)((tc.wt * 16 + t) * 128 + ((cg ^ t ^ 4) * 16));   // odd k-steps
    // 4-row boxes (one per k-step, 4 KB apart) [chunk][row][128 B]: line wt * 4 + t
    const uint32_t lane_b = (uint32_t)((tc.wt * 4 + t) * 128 + ((cg ^ (4 * (tc.wt & 1) + t)) * 16));
    const uint32_t lane_s = (uint32_t)(kScaleOff + (tc.wt * 32 + col_l) * 2);      // scale of its column
    const uint32_t lane_z = (uint32_t)(kZeroOff + (tc.wt * 4 + (cg >> 1)) * 4);    // the qzeros word holding its zero
    const int zshift = (cg & 1) * 16 + t * 4;
    const float unit = kSubnormal ? 16777216.0f : 1.0f;  // the accumulators are in units of 2^-24
    const int u_begin = u;
    if constexpr (XMODE != X_FULL) {
        if (u < u_end) {
#ifdef GPTQ_TRACE
            const long long stg0 = clock64();
#endif
            stage_range<XMODE, ACT>(p, tc, nk, u, u_end, perm);
#ifdef GPTQ_TRACE
            if (g_mega_trace != nullptr && threadIdx.x == 0) g_mega_trace[blockIdx.x * 64 + 36 + XMODE] += (unsigned long long)(clock64() - stg0);
#endif
        }
    }

#pragma unroll 1
    while (u < u_end) {
        const int slab_v = u / nk;
        const int ks0 = u - slab_v * nk;
        const int nseg = min(nk - ks0, u_end - u);
        const int mi = (NM > 1 && slab_v >= nslab) ? 1 : 0;
        float* outp = (mi ? out1 : out0) + (slab_v - mi * nslab) * kSlabCols + tc.wt * 32 + col_l;

        uint32_t xaddr;
        const float* xsum;
        if constexpr (XMODE == X_FULL) {
            xaddr = smem_u32(xs_full) + (ks0 * 32 + t * 8) * 2;
            xsum = xsum_full + ks0;
        } else {
            xaddr = smem_u32(tc.xseg) + ((u - u_begin) * 32 + t * 8) * 2;
            xsum = tc.xsum_seg + (u - u_begin);
        }

        float tot = 0.f;  // lane (g, t) finishes column 4g + t of the warp's stripe
        int left = nseg, gpos = ks0 % gs_steps;
#pragma unroll 1
        while (left > 0) {
            const int n = stage_steps(gpos, gs_steps, left);
            const uint32_t st = cons_wait(ring);
            float acc0[4], acc1[4];
            float xs4;
            if (n == kStageSteps) {
                uint4 q[4], xf[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) q[j] = lds128(st + j * 512 + ((j & 1) ? lane_a1 : lane_a0));
#pragma unroll
                for (int j = 0; j < 4; ++j) xf[j] = lds128(xaddr + j * 64);
                step_mma<true>(q[0], xf[0], acc0, acc1);
                step_mma<false>(q[1], xf[1], acc0, acc1);
                step_mma<false>(q[2], xf[2], acc0, acc1);
                step_mma<false>(q[3], xf[3], acc0, acc1);
                xs4 = (xsum[0] + xsum[1]) + (xsum[2] + xsum[3]);
            } else {
                {
                    const uint4 q = lds128(st + lane_b), xf = lds128(xaddr);
                    step_mma<true>(q, xf, acc0, acc1);
                    xs4 = xsum[0];
                }
#pragma unroll 1
                for (int j = 1; j < n; ++j) {
                    const uint4 q = lds128(st + lane_b + j * kStepBytes), xf = lds128(xaddr + j * 64);
                    step_mma<false>(q, xf, acc0, acc1);
                    xs4 += xsum[j];
                }
            }
            // group epilogue: tot += s * (acc * unit - z * sum(x))   (z = stored zero + 1, quant/quant_linear.py:120-121).
            // All four t lanes hold the same four column sums (the batch columns of B are copies): lane t finishes column 4g + t.
            {
                uint32_t sh, zw;
                asm volatile("ld.shared.u16 %0, [%1];" : "=r"(sh) : "r"(st + lane_s));
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(zw) : "r"(st + lane_z));
                const float a = (t & 2) ? ((t & 1) ? acc1[2] : acc1[0]) : ((t & 1) ? acc0[2] : acc0[0]);
                const float z = (float)(((zw >> zshift) & 15u) + 1u);
                tot = fmaf(__half2float(__ushort_as_half((unsigned short)sh)), fmaf(a, unit, -z * xs4), tot);
            }
            cons_release(ring, lane);
            xaddr += n * 64;
            xsum += n;
            left -= n;
            gpos += n;
            if (gpos == gs_steps) gpos = 0;
        }
        if (peers == nullptr) {
            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(outp), "f"(tot) : "memory");  // the warp's 32 columns: one 128-byte line
        } else {  // tensor parallelism, direct mode: the partial sum goes into out0's counterpart on every rank
            const size_t off = (size_t)(outp - out0);
            for (int q = 0; q < p.tp_size; ++q) asm volatile("red.relaxed.sys.global.add.f32 [%0], %1;" ::"l"(peers[q] + off), "f"(tot) : "memory");
        }
        u += nseg;
    }
}

// ---- attention ----------------------------------------------------------------------------------------------------------
// Work units (head, 32 keys); a team serves one head (attn_range); one unit = one ring stage (K rows, V rows).  The team writes
// one partial record (m, l, o[128]) to p.part[T].
__device__ void run_attention(const MegaParams& p, ConsRing& ring, const TeamCtx& tc, int layer) {
    const int pos = step_pos(p), Tlen = pos + 1;
    const int upb = (Tlen + kKeysPerUnit - 1) / kKeysPerUnit;
    int head, b0, b_end;
    attn_range(tc.T, p.n_heads, tc.nb, upb, head, b0, b_end);
    float* red_o = reinterpret_cast<float*>(tc.scratch);                 // [8][128]  end of a segment
    float* red_ml = reinterpret_cast<float*>(tc.scratch + 4096);         // [8][2]
    float* q_s = reinterpret_cast<float*>(tc.scratch);                   // [128]     start of a segment (aliases red_o)
    __half* knew = reinterpret_cast<__half*>(tc.scratch + 512);          // [128]
    __half* vnew = knew + kHD;                                          // [128]
    const int ttid = tc.ttid, lane = tc.lane, grp = ttid >> 3, j = ttid & 7;  // 32 groups of 8 lanes: group = key, lane j owns dims 8j..8j+7 and 64+8j..64+8j+7
    const int ub_new = pos / kKeysPerUnit;  // the unit that holds this step's key
    __half* kc_l = p.k_cache + layer * p.layer_stride;
    __half* vc_l = p.v_cache + layer * p.layer_stride;
    float* rec = p.part + (size_t)tc.T * kRec;
    if (b0 >= b_end) {  // no unit for this team (short context): a neutral record keeps the merge uniform
        if (ttid < kHD) rec[4 + ttid] = 0.f;
        if (ttid == 0) {
            rec[0] = -INFINITY;
            rec[1] = 0.f;
        }
        return;
    }
    {
        const int nseg = b_end - b0;
        const bool owns_new = (ub_new >= b0 && ub_new < b_end);
        team_sync(tc.team);  // scratch reuse
        if (ttid < kHD) {
            const int i = ttid & 63;
            const bool hi = ttid >= 64;
            const float c = ld_cg(p.rope_cs + i), s = ld_cg(p.rope_cs + 64 + i);
            const float* aq = p.acc_qkv + head * kHD;
            const float qx = __half2float(__float2half_rn(ld_cg(aq + i))), qy = __half2float(__float2half_rn(ld_cg(aq + i + 64)));  // the qkv projection output is fp16
            const float qr = hi ? __fadd_rn(__fmul_rn(qx, s), __fmul_rn(qy, c)) : __fsub_rn(__fmul_rn(qx, c), __fmul_rn(qy, s));
            q_s[ttid] = __half2float(__float2half_rn(qr));
            if (owns_new) {  // this segment owns the new key/value: RoPE(k), append both to the cache
                const float* ak = aq + p.Hq;
                const float* av = aq + 2 * p.Hq;
                const float kx = __half2float(__float2half_rn(ld_cg(ak + i))), ky = __half2float(__float2half_rn(ld_cg(ak + i + 64)));
                const float kr = hi ? __fadd_rn(__fmul_rn(kx, s), __fmul_rn(ky, c)) : __fsub_rn(__fmul_rn(kx, c), __fmul_rn(ky, s));
                const __half kh = __float2half_rn(kr), vh = __float2half_rn(ld_cg(av + ttid));
                knew[ttid] = kh;
                vnew[ttid] = vh;
                const size_t off = ((size_t)head * p.max_seq + pos) * kHD + ttid;
                kc_l[off] = kh;
                vc_l[off] = vh;
            }
        }
        team_sync(tc.team);
        float qr[16];
#pragma unroll
        for (int d = 0; d < 8; ++d) {
            qr[d] = q_s[8 * j + d];
            qr[8 + d] = q_s[64 + 8 * j + d];
        }
        float mloc = -INFINITY, lloc = 0.f, o[16];
#pragma unroll
        for (int d = 0; d < 16; ++d) o[d] = 0.f;
#pragma unroll 1
        for (int b = b0; b < b0 + nseg; ++b) {
            const uint32_t st = cons_wait(ring);
            if (b == ub_new) {
                // the stage was fetched before this step's key/value existed: patch its row from the fresh values
                team_sync(tc.team);  // (uniform per team) every warp has seen the stage land
                const int row = pos - b * kKeysPerUnit;
                if (ttid < 16) {
                    sts128(st + row * 256 + ttid * 16, reinterpret_cast<const uint4*>(knew)[ttid]);
                    fence_proxy_async_smem();  // generic-proxy writes into a stage the TMA unit will overwrite later
                } else if (ttid < 32) {
                    sts128(st + kVOff + row * 256 + (ttid - 16) * 16, reinterpret_cast<const uint4*>(vnew)[ttid - 16]);
                    fence_proxy_async_smem();
                }
                team_sync(tc.team);
            }
            const int key = b * kKeysPerUnit + grp;
            const bool valid = key < Tlen;
            float s = 0.f;
            if (valid) {
                const uint4 k0 = lds128(st + grp * 256 + j * 16), k1 = lds128(st + grp * 256 + 128 + j * 16);
                const uint32_t w[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float2 f = __half22float2(u32_as_h2(w[e]));
                    s = fmaf(qr[2 * e], f.x, s);
                    s = fmaf(qr[2 * e + 1], f.y, s);
                }
            }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            if (valid) {
                s *= p.scale;
                const float mnew = fmaxf(mloc, s);
                const float alpha = (mloc == -INFINITY) ? 0.f : expf(mloc - mnew);
                const float pw = expf(s - mnew);
                lloc = fmaf(lloc, alpha, pw);
                const uint4 v0 = lds128(st + kVOff + grp * 256 + j * 16), v1 = lds128(st + kVOff + grp * 256 + 128 + j * 16);
                const uint32_t w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float2 f = __half22float2(u32_as_h2(w[e]));
                    o[2 * e] = fmaf(o[2 * e], alpha, pw * f.x);
                    o[2 * e + 1] = fmaf(o[2 * e + 1], alpha, pw * f.y);
                }
                mloc = mnew;
            }
            cons_release(ring, lane);
        }
        team_sync(tc.team);  // every warp has read q_s and passed the patch: the merge buffers (which alias q_s / knew / vnew) may be written
        // merge the 4 key groups of the warp (lanes xor 8, 16), then the 8 warps through shared memory
#pragma unroll
        for (int sh = 8; sh <= 16; sh <<= 1) {
            const float mo = __shfl_xor_sync(0xffffffffu, mloc, sh), lo = __shfl_xor_sync(0xffffffffu, lloc, sh);
            const float mn = fmaxf(mloc, mo);
            const float wa = (mloc == -INFINITY) ? 0.f : expf(mloc - mn), wb = (mo == -INFINITY) ? 0.f : expf(mo - mn);
            lloc = lloc * wa + lo * wb;
#pragma unroll
            for (int d = 0; d < 16; ++d) {
                const float oo = __shfl_xor_sync(0xffffffffu, o[d], sh);
                o[d] = o[d] * wa + oo * wb;
            }
            mloc = mn;
        }
        if (lane < 8) {
            if (lane == 0) {
                red_ml[tc.wt * 2] = mloc;
                red_ml[tc.wt * 2 + 1] = lloc;
            }
#pragma unroll
            for (int d = 0; d < 8; ++d) {
                red_o[tc.wt * kHD + 8 * lane + d] = o[d];
                red_o[tc.wt * kHD + 64 + 8 * lane + d] = o[8 + d];
            }
        }
        team_sync(tc.team);
        if (ttid < kHD) {
            float M = -INFINITY;
#pragma unroll
            for (int w = 0; w < kTeamWarps; ++w) M = fmaxf(M, red_ml[w * 2]);
            float L = 0.f, O = 0.f;
#pragma unroll
            for (int w = 0; w < kTeamWarps; ++w) {
                const float mw = red_ml[w * 2];
                const float wgt = (mw == -INFINITY) ? 0.f : expf(mw - M);
                L = fmaf(red_ml[w * 2 + 1], wgt, L);
                O = fmaf(red_o[w * kHD + ttid], wgt, O);
            }
            rec[4 + ttid] = O;
            if (ttid == 0) {
                rec[0] = M;
                rec[1] = L;
            }
        }
    }
}

// ---- lm_head: fp16 [V, H] rows through the ring, lm_rows whole rows per stage -------------------------------------------------
__device__ void run_lm_head(const MegaParams& p, ConsRing& ring, const TeamCtx& tc, const __half* xs_plain) {
    const int R = p.lm_rows, H = p.H, nch = H / 8;
    int u, u_end;
    const int nloc = p.v1 - p.v0;
    team_range(tc.T, (unsigned)((nloc + R - 1) / R), tc.nb, u, u_end);
    float* part_s = reinterpret_cast<float*>(tc.scratch);  // [2][R][8]
    // this thread's chunks of x (k = 8 * (ttid + 256 i)) stay in registers for the whole op
    float xr[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = tc.ttid + i * kTeamThreads;
        if (c < nch) {
            const uint4 v = *reinterpret_cast<const uint4*>(xs_plain + c * 8);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f = __half22float2(u32_as_h2(w[e]));
                xr[i][2 * e] = f.x;
                xr[i][2 * e + 1] = f.y;
            }
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) xr[i][e] = 0.f;
        }
    }
    int buf = 0;
#pragma unroll 1
    for (; u < u_end; ++u) {
        const int nrows = min(R, nloc - u * R);
        const uint32_t st = cons_wait(ring);
        float* ps = part_s + buf * (R * kTeamWarps);
#pragma unroll 1
        for (int r = 0; r < nrows; ++r) {
            float a = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int c = tc.ttid + i * kTeamThreads;
                if (c < nch) {
                    const uint4 wv = lds128(st + (uint32_t)(r * H * 2 + c * 16));
                    const uint32_t w[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 f = __half22float2(u32_as_h2(w[e]));
                        a = fmaf(f.x, xr[i][2 * e], a);
                        a = fmaf(f.y, xr[i][2 * e + 1], a);
                    }
                }
            }
            a = warp_sum(a);
            if (tc.lane == 0) ps[r * kTeamWarps + tc.wt] = a;
        }
        cons_release(ring, tc.lane);
        team_sync(tc.team);  // partials of this stage are visible; the other buffer is free again
        if (tc.ttid < nrows) {
            float a = 0.f;
#pragma unroll
            for (int w = 0; w < kTeamWarps; ++w) a += ps[tc.ttid * kTeamWarps + w];
            const __half lg = __float2half_rn(a);
            for (int q = 0; q < p.tp_size; ++q) p.logits_peer[q][(size_t)p.v0 + (size_t)u * R + tc.ttid] = lg;  // every rank holds all logits
        }
        buf ^= 1;
    }
}

template <bool ACT>
__global__ void __launch_bounds__(kBlock, 1) llama_decode_mega_kernel(const __grid_constant__ MegaParams p) {
    extern __shared__ __align__(16) uint8_t smem_dyn[];
    uint8_t* smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);  // TMA swizzle atoms: 1 KB aligned stages (1 KB of slack is allocated)
    __shared__ float red_s[kConsumerWarps];
    __shared__ __align__(8) unsigned long long bars_s[kTeams][2 * kMaxStages];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // smem (1 KB aligned): [team 0 ring][team 1 ring][xs: H halves][xsum: H/32 floats][tmp / team scratch]
    const uint32_t ring_bytes = (uint32_t)p.n_stages * kStageBytes;
    __half* xs = reinterpret_cast<__half*>(smem_raw + kTeams * ring_bytes);
    float* xsum = reinterpret_cast<float*>(xs + p.H);
    uint8_t* tmp_raw = reinterpret_cast<uint8_t*>(xsum + p.H / 32);
    tmp_raw += (16 - (reinterpret_cast<uintptr_t>(tmp_raw) & 15)) & 15;
    __half* tmp = reinterpret_cast<__half*>(tmp_raw);

    if (tid == 0) {
        for (int tm = 0; tm < kTeams; ++tm)
            for (int s = 0; s < p.n_stages; ++s) {
                mbar_init(smem_u32(&bars_s[tm][s]), 1);
                mbar_init(smem_u32(&bars_s[tm][kMaxStages + s]), kTeamWarps);
            }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();  // the only block-wide barrier: after it the producer warps and the consumers never meet again
    const unsigned nb = gridDim.x * kTeams;
    if (warp >= kConsumerWarps) {
        if (lane == 0) {
            const int tm = warp - kConsumerWarps;
            ProdRing r;
            r.ring = smem_u32(smem_raw) + tm * ring_bytes;
            r.full = smem_u32(&bars_s[tm][0]);
            r.empty = smem_u32(&bars_s[tm][kMaxStages]);
            r.stage = 0;
            r.use = 0;
            r.nstages = p.n_stages;
#ifdef GPTQ_TRACE
            r.blocked = 0;
#endif
            producer_loop(p, r, blockIdx.x * kTeams + tm, nb);
        }
        return;
    }
    TeamCtx tc;
    tc.team = warp / kTeamWarps;
    tc.wt = warp % kTeamWarps;
    tc.lane = lane;
    tc.ttid = tid - tc.team * kTeamThreads;
    tc.T = blockIdx.x * kTeams + tc.team;
    tc.nb = nb;
    tc.xseg = xs + tc.team * (p.H / 2);
    tc.xsum_seg = xsum + tc.team * (p.H / 64);
    tc.scratch = tmp_raw + tc.team * kTeamScratch;
    ConsRing ring;
    ring.ring = smem_u32(smem_raw) + tc.team * ring_bytes;
    ring.full = smem_u32(&bars_s[tc.team][0]);
    ring.tile = ring.ring;
    ring.bar = ring.full;
    ring.parity = 0;
    ring.left = p.n_stages;
    ring.nstages = p.n_stages;
#ifdef GPTQ_TRACE
    ring.waited = 0;
    ring.stages = 0;
#endif

    unsigned long long gen;  // barrier target (meaningful in thread 0): the counter value when this launch began
    {
        // bar[1] = counter value at the end of the previous launch (written by CTA 0 after its last barrier): CTAs that
        // start late may already see arrivals of this launch in bar[0], never in bar[1]
        unsigned long long v;
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p.bar + 1) : "memory");
        gen = v;
    }
    unsigned long long xgen = 0;  // the same for the cross-GPU barrier (tensor parallelism)
    if (p.tp_size > 1) asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(xgen) : "l"(p.xbar_peer[p.tp_rank] + 32 * kMaxTP) : "memory");
    // after the o_proj, down_proj and lm_head operations every rank needs every rank's contributions
    auto sync_all_ranks = [&](float* loc, float* const* peers) {
        if (p.tp_size == 1) {
            grid_barrier(p.bar, gen);
        } else {
            if (p.tp_exchange && loc != nullptr) {
                grid_barrier(p.bar, gen);
                tp_exchange(p, loc, peers, p.H);
            }
            tp_barrier(p, xgen);
        }
    };

    // this step's RoPE angles (quant/fused_attn.py:43,91): freq_i = exp(i * inv_base) * pos
    if (blockIdx.x == 0 && tid < 64) {
        const float f = expf((float)tid * p.inv_base) * (float)step_pos(p);
        p.rope_cs[tid] = cosf(f);
        p.rope_cs[64 + tid] = sinf(f);
    }

    // residual stream: every stage_norm reads buffer `cur` (or the embedding row) and writes the other one
    const int token = min(max(p.tokens[0], 0), p.V - 1);
    const __half* resid_src = p.embed + (size_t)token * p.H;
    const float* resid_acc = nullptr;
    int cur = 1;
#pragma unroll 1
    for (int l = 0; l < p.n_layers; ++l) {
        const LayerDesc& L = p.layers[l];
        // ---- Q ----
        MTRACE(l * 12 + 0);
        stage_norm<ACT, false>(p, resid_src, resid_acc, L.input_norm, p.resid[cur ^ 1], xs, xsum, tmp, red_s, L.qkv_perm);
        MTRACE(l * 12 + 1);
        cur ^= 1;
        zero_slice(p.acc_g, p.I);  // last read by the previous layer's D
        zero_slice(p.acc_u, p.I);
        {
            OPTRACE_BEGIN(ring);
            run_matvec<1, X_FULL>(p, ring, tc, L.qkv.gs_steps, p.H, 3 * p.Hq, p.acc_qkv, nullptr, xs, xsum);
            OPTRACE_END(ring, l, 0);
        }
        MTRACE(l * 12 + 2);
        grid_barrier(p.bar, gen);
        MTRACE(l * 12 + 3);
        // ---- A ----
        zero_slice(p.acc_d, p.H);  // last read by this layer's Q
        {
            OPTRACE_BEGIN(ring);
            run_attention(p, ring, tc, l);
            OPTRACE_END(ring, l, 1);
        }
        MTRACE(l * 12 + 4);
        grid_barrier(p.bar, gen);
        MTRACE(l * 12 + 5);
        // ---- O ----
        zero_slice(p.acc_qkv, 3 * p.Hq);
        {
            OPTRACE_BEGIN(ring);
            run_matvec<1, X_ATTN, ACT>(p, ring, tc, L.o.gs_steps, p.Hq, p.H, p.tp_exchange ? p.acc_o_loc : p.acc_o, nullptr, xs, xsum, L.o_perm, (p.tp_size > 1 && !p.tp_exchange) ? p.acc_o_peer : nullptr);
            OPTRACE_END(ring, l, 2);
        }
        MTRACE(l * 12 + 6);
        sync_all_ranks(p.acc_o_loc, p.acc_o_peer);
        MTRACE(l * 12 + 7);
        // ---- G ----
        stage_norm<ACT, false>(p, p.resid[cur], p.acc_o, L.post_norm, p.resid[cur ^ 1], xs, xsum, tmp, red_s, L.mlp_perm);
        cur ^= 1;
        MTRACE(l * 12 + 8);
        {
            OPTRACE_BEGIN(ring);
            run_matvec<2, X_FULL>(p, ring, tc, L.gate.gs_steps, p.H, p.I, p.acc_g, p.acc_u, xs, xsum);
            OPTRACE_END(ring, l, 3);
        }
        MTRACE(l * 12 + 9);
        grid_barrier(p.bar, gen);
        // ---- D ----
        zero_slice(p.acc_o, p.H);
        MTRACE(l * 12 + 10);
        {
            OPTRACE_BEGIN(ring);
            run_matvec<1, X_SWIGLU>(p, ring, tc, L.down.gs_steps, p.I, p.H, p.tp_exchange ? p.acc_d_loc : p.acc_d, nullptr, xs, xsum, nullptr, (p.tp_size > 1 && !p.tp_exchange) ? p.acc_d_peer : nullptr);
            OPTRACE_END(ring, l, 4);
        }
        MTRACE(l * 12 + 11);
        sync_all_ranks(p.acc_d_loc, p.acc_d_peer);
        resid_src = p.resid[cur];
        resid_acc = p.acc_d;
    }
    // ---- L: final norm + lm_head ----
    stage_norm<false, true>(p, resid_src, resid_acc, p.final_norm, nullptr, xs, xsum, tmp, red_s, nullptr);
    zero_slice(p.acc_g, p.I);
    zero_slice(p.acc_u, p.I);
    run_lm_head(p, ring, tc, xs);
    sync_all_ranks(nullptr, nullptr);
    zero_slice(p.acc_d, p.H);
    if (blockIdx.x == 0) {
        if (tid == 0) {  // every CTA has arrived at the last barrier: the counters rest at these values until the next launch
            p.bar[1] = gen;
            if (p.tp_size > 1) p.xbar_peer[p.tp_rank][32 * kMaxTP] = xgen;
        }
        if (p.next_token != nullptr) {  // greedy argmax (lowest index wins ties)
            float best = -INFINITY;
            int idx = 0x7fffffff;
            for (int i = tid; i < p.V; i += kConsumers) {
                const float v = __half2float(ld_cg_h(p.logits + i));  // written by other CTAs
                if (v > best || (v == best && i < idx)) {
                    best = v;
                    idx = i;
                }
            }
            __shared__ float sv[kConsumerWarps];
            __shared__ int si[kConsumerWarps];
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
                for (int w = 1; w < kConsumerWarps; ++w)
                    if (sv[w] > best || (sv[w] == best && si[w] < idx)) {
                        best = sv[w];
                        idx = si[w];
                    }
                p.next_token[0] = idx;
            }
        }
    }
}

inline size_t al256(size_t v) { return (v + 255) & ~(size_t)255; }

// 3-D view of packed matrices with row stride N * 4 bytes: {32 words (128 B), packed rows, 128-byte chunks}, box 32 x box_rows x 8
// (= box_rows rows of a 256-column slab), 128-byte swizzle.  cuTensorMapEncodeTiled comes through the runtime's driver entry
// point (libcuda is not linked: the library must load without a driver).
bool encode_weight_map(CUtensorMap* tm, const void* base, int rows, uint64_t chunks, int N, int box_rows) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || fn == nullptr) return false;
    const auto encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
    const cuuint64_t dims[3] = {32, (cuuint64_t)rows, (cuuint64_t)chunks};  // innermost first
    const cuuint64_t strides[2] = {(cuuint64_t)N * 4, 128};                 // bytes: packed row, chunk
    const cuuint32_t box[3] = {32, (cuuint32_t)box_rows, (cuuint32_t)(kSlabCols / 32)};
    const cuuint32_t estr[3] = {1, 1, 1};
    return encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// device properties that shape the launch (queried per call: no cached global state)
struct MegaPlan {
    int grid, n_stages, lm_rows;
    size_t smem;
};

// Shared-memory plan for this model on a device with `sms` SMs and `smem_max` bytes of opt-in shared memory per block;
// returns false if the shape does not fit the kernel's staging buffers.
bool mega_plan(const gptq_llama_model& m, int sms, size_t smem_max, size_t smem_static, MegaPlan& pl) {
    const int H = m.hidden, I = m.intermediate;
    const size_t fixed = (size_t)H * 2 + (size_t)(H / 32) * 4 + 16 + max((size_t)H * 2, (size_t)kTeams * kTeamScratch);
    const size_t other = fixed + smem_static + 1024;  // + the kernel's static shared memory + 1 KB alignment slack of the rings
    if (smem_max < other) return false;
    int st = (int)((smem_max - other) / ((size_t)kTeams * kStageBytes));
    st = min(st, kMaxStages);
    if (st < 2) return false;
    pl.grid = sms;
    pl.n_stages = st;
    pl.smem = fixed + 1024 + (size_t)kTeams * st * kStageBytes;
    pl.lm_rows = max(1, kLmStageBytes / (H * 2));
    if ((size_t)pl.lm_rows * H * 2 > (size_t)kStageBytes) return false;
    if ((size_t)2 * pl.lm_rows * kTeamWarps * 4 > (size_t)kTeamScratch) return false;
    // per-team k-segments of o_proj / down_proj are staged in half of the xs buffer, their step sums in half of xsum
    const long long nteams = (long long)sms * kTeams;
    const int Hq = m.n_heads * m.head_dim;  // attention width of this rank (= H on a single GPU)
    const long long seg_o = ((long long)(H / kSlabCols) * (Hq / 32) + nteams - 1) / nteams + 1;
    const long long seg_d = ((long long)(H / kSlabCols) * (I / 32) + nteams - 1) / nteams + 1;
    const long long seg = max(seg_o, seg_d);
    if (seg * 32 > H / 2 || seg > H / 64) return false;
    if (m.n_heads > nteams) return false;  // a team's attention range must meet at most two heads
    // 32-bit range arithmetic of team_range: (T + 1) * U must stay below 2^32 for every operation's unit count U
    const long long umax = max(max((long long)2 * (I / kSlabCols) * (H / 32), (long long)(H / kSlabCols) * (I / 32)),
                               max((long long)m.vocab, (long long)m.n_heads * 4096));
    if (umax * (nteams + 1) >= (1ll << 32)) return false;
    return true;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
bool mega_supported(const gptq_llama_model& m, const gptq_llama_state& st) {
    if (st.batch != 1 || m.n_layers > kMaxLayers || m.head_dim != kHD) return false;
    if (m.hidden % kSlabCols || m.intermediate % kSlabCols || m.hidden > 8192 || m.intermediate > 32768 || m.hidden % 64) return false;
    if ((3 * m.n_heads * m.head_dim) % kSlabCols) return false;  // the (local) fused qkv width is dealt in 256-column slabs
    if (st.tp != nullptr) {
        const gptq_llama_tp& tp = *st.tp;
        if (tp.size < 1 || tp.size > kMaxTP || tp.rank < 0 || tp.rank >= tp.size) return false;
        if (tp.vocab_begin < 0 || tp.vocab_end > m.vocab || tp.vocab_begin >= tp.vocab_end) return false;
        for (int q = 0; q < tp.size; ++q)
            if (q != tp.rank && (tp.peer_scratch[q] == nullptr || tp.peer_logits[q] == nullptr)) return false;
    } else if (m.n_heads * m.head_dim != m.hidden) {
        return false;
    }
    for (int l = 0; l < m.n_layers; ++l) {
        const gptq_llama_layer& ly = m.layers[l];
        const gptq_qweight* ws[5] = {&ly.qkv, &ly.o, &ly.gate, &ly.up, &ly.down};
        for (const gptq_qweight* w : ws) {
            if (w->bits != 4 || w->groupsize <= 0 || w->groupsize % 32) return false;
            if ((reinterpret_cast<uintptr_t>(w->qweight) & 15) || (reinterpret_cast<uintptr_t>(w->scales) & 15) || (reinterpret_cast<uintptr_t>(w->qzeros) & 15))
                return false;
        }
        if (ly.gate.groupsize != ly.up.groupsize) return false;
    }
    if ((reinterpret_cast<uintptr_t>(m.lm_head) & 15) || (reinterpret_cast<uintptr_t>(st.k_cache) & 15) || (reinterpret_cast<uintptr_t>(st.v_cache) & 15)) return false;
    return true;
}

size_t mega_scratch_bytes(const gptq_llama_model& m, int max_seq) {
    (void)max_seq;
    const size_t max_teams = 1024;  // >= kTeams * SM count of any device this library runs on
    return al256((size_t)m.hidden * 2) * 2 + al256((size_t)3 * m.hidden * 4) + al256((size_t)m.hidden * 4) * 2 + al256((size_t)m.intermediate * 4) * 2 +
           al256(max_teams * kRec * 4) + al256(128 * 4) + 256 + (kMaxTP + 1) * 256 + al256((size_t)m.hidden * 4) * 2;
}

cudaError_t launch_decode_mega(const gptq_llama_model& m, const gptq_llama_state& st, uint8_t* scratch, cudaStream_t stream) {
    static_assert(sizeof(MegaParams) < 32000, "kernel parameter space");
    int dev = 0, sms = 0, smem_optin = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess)
        return cudaErrorInvalidDevice;
    bool any_perm = false;
    for (int l = 0; l < m.n_layers; ++l) any_perm = any_perm || m.layers[l].qkv_perm != nullptr || m.layers[l].o_perm != nullptr || m.layers[l].mlp_perm != nullptr;
    auto kernel = any_perm ? llama_decode_mega_kernel<true> : llama_decode_mega_kernel<false>;
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, kernel) != cudaSuccess) return cudaErrorInvalidDeviceFunction;
    MegaPlan pl;
    if (sms * kTeams > 1024 || !mega_plan(m, sms, (size_t)smem_optin, fa.sharedSizeBytes, pl)) return cudaErrorInvalidConfiguration;

    MegaParams p{};
    p.n_layers = m.n_layers; p.H = m.hidden; p.Hq = m.n_heads * m.head_dim; p.I = m.intermediate; p.V = m.vocab; p.n_heads = m.n_heads;
    p.v0 = st.tp != nullptr ? st.tp->vocab_begin : 0;
    p.v1 = st.tp != nullptr ? st.tp->vocab_end : m.vocab;
    p.max_seq = st.max_seq;
    p.n_stages = pl.n_stages;
    p.lm_rows = pl.lm_rows;
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
    // the head of the region has the same layout on every tensor-parallel rank (it depends on the hidden size only): peers address
    // acc_o / acc_d / xbar of this rank through its scratch base
    p.resid[0] = reinterpret_cast<__half*>(take((size_t)m.hidden * 2));
    p.resid[1] = reinterpret_cast<__half*>(take((size_t)m.hidden * 2));
    p.acc_o = reinterpret_cast<float*>(take((size_t)m.hidden * 4));
    p.acc_d = reinterpret_cast<float*>(take((size_t)m.hidden * 4));
    unsigned long long* xbar = reinterpret_cast<unsigned long long*>(take((kMaxTP + 1) * 256));
    const gptq_llama_tp* tp = st.tp;
    p.tp_size = tp != nullptr ? tp->size : 1;
    p.tp_rank = tp != nullptr ? tp->rank : 0;
    // measured on 65B: direct peer REDs win up to 4 ranks, the local reduction + slice exchange at 8
    p.tp_exchange = (tp != nullptr && tp->size > 1) ? (tp->reduce_mode == 0 ? (tp->size > 4) : (tp->reduce_mode == 2)) : 0;
    for (int q = 0; q < p.tp_size; ++q) {
        uint8_t* base = (tp != nullptr && q != tp->rank) ? reinterpret_cast<uint8_t*>(tp->peer_scratch[q]) : scratch;
        p.acc_o_peer[q] = reinterpret_cast<float*>(base + (reinterpret_cast<uint8_t*>(p.acc_o) - scratch));
        p.acc_d_peer[q] = reinterpret_cast<float*>(base + (reinterpret_cast<uint8_t*>(p.acc_d) - scratch));
        p.xbar_peer[q] = reinterpret_cast<unsigned long long*>(base + (reinterpret_cast<uint8_t*>(xbar) - scratch));
        p.logits_peer[q] = reinterpret_cast<__half*>((tp != nullptr && q != tp->rank) ? tp->peer_logits[q] : st.logits);
    }
    p.acc_o_loc = reinterpret_cast<float*>(take((size_t)m.hidden * 4));
    p.acc_d_loc = reinterpret_cast<float*>(take((size_t)m.hidden * 4));
    p.acc_qkv = reinterpret_cast<float*>(take((size_t)3 * p.Hq * 4));
    p.acc_g = reinterpret_cast<float*>(take((size_t)m.intermediate * 4));
    p.acc_u = reinterpret_cast<float*>(take((size_t)m.intermediate * 4));
    p.part = reinterpret_cast<float*>(take((size_t)1024 * kRec * 4));
    p.rope_cs = reinterpret_cast<float*>(take(128 * 4));
    p.bar = reinterpret_cast<unsigned long long*>(take(256));
    // One tensor map per row stride serves every layer: class 0 = qkv (N = 3H), 1 = o and down (N = H), 2 = gate and up (N = I).
    // Its base is the lowest qweight address of the class; a matrix is addressed through the chunk coordinate (128-byte units).
    const int classN[3] = {3 * m.n_heads * m.head_dim, m.hidden, m.intermediate};
    uintptr_t base[3] = {UINTPTR_MAX, UINTPTR_MAX, UINTPTR_MAX}, top[3] = {0, 0, 0};
    int rows_max[3] = {0, 0, 0};
    auto visit = [&](const gptq_qweight& w, int c) {
        const uintptr_t a = reinterpret_cast<uintptr_t>(w.qweight);
        base[c] = a < base[c] ? a : base[c];
        top[c] = a > top[c] ? a : top[c];
        rows_max[c] = max(rows_max[c], w.K / 8);
    };
    for (int l = 0; l < m.n_layers; ++l) {
        const gptq_llama_layer& ly = m.layers[l];
        visit(ly.qkv, 0); visit(ly.o, 1); visit(ly.down, 1); visit(ly.gate, 2); visit(ly.up, 2);
    }
    for (int c = 0; c < 3; ++c) {
        if ((base[c] & 127) != 0) return cudaErrorInvalidConfiguration;
        const uint64_t chunks = (uint64_t)(top[c] - base[c]) / 128 + (uint64_t)classN[c] / 32;
        if (chunks >= (1ull << 31)) return cudaErrorInvalidConfiguration;
        for (int v = 0; v < 2; ++v)
            if (!encode_weight_map(&p.tmaps[3 * v + c], reinterpret_cast<const void*>(base[c]), rows_max[c], chunks, classN[c], v == 0 ? 4 * kStageSteps : 4))
                return cudaErrorNotSupported;
    }
    bool act = false;  // any act-order gather: the ACT instantiation (the plain one carries no trace of the feature)
    for (int l = 0; l < m.n_layers; ++l) {
        const gptq_llama_layer& ly = m.layers[l];
        bool aligned = true;
        auto md = [&](const gptq_qweight& w, int c) {
            MatDesc d;
            d.sc = reinterpret_cast<const __half*>(w.scales);
            d.qz = reinterpret_cast<const uint32_t*>(w.qzeros);
            d.gs_steps = w.groupsize / 32;
            d.tmap = c;
            const uintptr_t delta = reinterpret_cast<uintptr_t>(w.qweight) - base[c];
            aligned = aligned && (delta % 128 == 0);
            d.chunk0 = (int)(delta / 128);
            return d;
        };
        p.layers[l].qkv = md(ly.qkv, 0);
        p.layers[l].o = md(ly.o, 1);
        p.layers[l].gate = md(ly.gate, 2);
        p.layers[l].up = md(ly.up, 2);
        p.layers[l].down = md(ly.down, 1);
        if (!aligned) return cudaErrorInvalidConfiguration;
        p.layers[l].input_norm = reinterpret_cast<const __half*>(ly.input_norm);
        p.layers[l].post_norm = reinterpret_cast<const __half*>(ly.post_norm);
        p.layers[l].qkv_perm = ly.qkv_perm;
        p.layers[l].o_perm = ly.o_perm;
        p.layers[l].mlp_perm = ly.mlp_perm;
        act = act || ly.qkv_perm != nullptr || ly.o_perm != nullptr || ly.mlp_perm != nullptr;
    }
    (void)act;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
    if (e != cudaSuccess) return e;
    int occ = 0;
    // cooperative launch: every CTA must be co-resident (one per SM); if the device cannot host them, the caller falls back to
    // the kernel-chain engine instead of risking a barrier deadlock
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kBlock, pl.smem) != cudaSuccess || occ < 1) return cudaErrorCooperativeLaunchTooLarge;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(pl.grid);
    cfg.blockDim = dim3(kBlock);
    cfg.dynamicSmemBytes = pl.smem;
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
