#!/usr/bin/env python
"""Headline benchmark: tokens/sec of LLaMA-7B int4 g128 batch-1 decode on B200 (BASELINE.json metric),
plus the roofline of the dominant kernel and the CPU baseline, as ONE JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference's arithmetic on the host cores

A "step" is one decoded token: one replay of the captured CUDA graph of gptq_llama_decode_step over a
random-init LLaMA-7B-shaped GPTQ model (32 distinct layers, 3.6 GB of packed weights per step, i.e. far
larger than the 126 MB L2, so every step streams from HBM) at context position seq-1 = 2047.
N > 1 (torchrun): the path does not shard at this size ("replicas only"): every rank decodes
its own sequence on its own GPU, no data-path collective; value = total tokens/s, scaling = weak.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, 'gptq-for-llama_b200')):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

METRIC = 'tokens/sec LLaMA-7B int4 g128 batch=1; matvec HBM GB/s vs 8 TB/s roofline'
SEQ = 2048
NCU_TRAFFIC_BYTES = 4447960888  # per launch of llama_decode_mega_kernel: 4.4350 GB read + 12.9 MB written (profiles/r1_mega_final_summary.txt)
BITS, GROUP = 4, 128


def alg_bytes_qlinear(K, N, M=1, bits=BITS, gs=GROUP):
    """qweight + scales + qzeros + g_idx + x + out."""
    G = (K + gs - 1) // gs
    return K * N * bits // 8 + G * N * 2 + G * N * bits // 8 + 4 * K + 2 * M * K + 2 * M * N


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        return json.load(open(path)).get('hbm_gbs', 6650.0), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


# ----------------------------------------------------------------------------------------------------
# CPU arm: the oracle's restatement of the reference kernels (the reference has no CPU forward), timed on
# a bounded sample: the quantized linears of ONE decoder layer at M=1, scaled to a 32-layer token.
# ----------------------------------------------------------------------------------------------------
def cpu_layer_seconds(reps):
    from oracle import gptq_oracle as O
    from oracle import cref
    Q = cref if cref.available() else O  # C/OpenMP restatement (all host threads) when built, else the numpy/torch one
    hidden, inter = 4096, 11008
    shapes = {'qkv': (hidden, 3 * hidden), 'o': (hidden, hidden), 'gate': (hidden, inter), 'up': (hidden, inter), 'down': (inter, hidden)}
    W = {k: O.random_packed(K, N, BITS, GROUP, seed=i)[:4] for i, (k, (K, N)) in enumerate(shapes.items())}
    x = torch.randn(1, hidden, generator=torch.Generator().manual_seed(0)).half()
    nw = torch.ones(hidden).half()

    def layer():
        h = O.rmsnorm_fwd(x, nw, 1e-6)
        Q.qlinear_fwd(h, *W['qkv'], BITS)
        Q.qlinear_fwd(x, *W['o'], BITS)
        mid = Q.fused_mlp_fwd(h, W['gate'], W['up'], BITS)
        Q.qlinear_fwd(mid, *W['down'], BITS)

    layer()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        layer()
        ts.append(time.perf_counter() - t0)
    return statistics.median(ts)


def cpu_baseline(reps=5):
    t_layer = cpu_layer_seconds(reps)
    n_layers = 32
    return {
        'value': 1.0 / (t_layer * n_layers),
        'unit': 'tokens/s',
        'cores': os.cpu_count(),
        'kind': 'port',
        'sample': f'oracle C/OpenMP restatement of matmul_248/fusedmatmul_248 (oracle/qlinear_ref.c, all host threads) on the 5 quantized linears of 1 of 32 LLaMA-7B layers at M=1, '
                  f'median of {reps} passes ({t_layer:.3f} s/layer) x 32 layers; lm_head and attention excluded (favours the CPU)',
    }, t_layer


def run_reference(args):
    """`--impl reference`: the reference's own arithmetic on the host cores (torch threads = all cores)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    steps, warm = max(1, args.steps), max(0, args.warmup)
    reps = min(max(steps, 1), 20)
    for _ in range(min(warm, 1)):
        cpu_layer_seconds(1)
    base, t_layer = cpu_baseline(reps)
    ms = t_layer * 32 * 1e3
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': base['value'], 'unit': 'tokens/s', 'n_gpus': args.gpus, 'steps': reps, 'warmup': min(warm, 1),
        'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f16', 'data': 'synthetic',
        'config': {'workload': f'LLaMA-7B int4 g128 batch=1 decode, seq={SEQ}', 'note': 'CPU: bounded sample (1 layer x 32)'},
        'cpu_baseline': base,
        'e2e': {'value': base['value'], 'unit': 'tokens/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.FIELDS}', '--format=csv,noheader,nounits', '-lms', '100', '-i', str(index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        for ln in out.strip().splitlines():
            f = [t.strip() for t in ln.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'samples': len(sm), 'reasons': sorted(reasons)}


def timed(fn, steps, dist_on):
    """EXACTLY `steps` calls of fn bracketed by barrier + synchronize; device time from CUDA events."""
    import torch.distributed as dist
    if dist_on:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    if dist_on:
        dist.barrier()
    return a.elapsed_time(b) / 1e3


def run_ours(args):
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    dist_on = world > 1
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if dist_on:
        dist.init_process_group('nccl', device_id=dev)
    from gptq_b200 import engine, ops

    steps, warm = max(1, args.steps), max(3, args.warmup)
    dec = engine.synthetic_llama('7b', bits=BITS, groupsize=GROUP, device=str(dev), seed=rank, max_seq=SEQ)
    # synthetic context: the cache holds seq-1 = 2047 tokens of random K/V; the step decodes token 2048
    dec.k_cache.normal_(0, 0.5)
    dec.v_cache.normal_(0, 0.5)
    pos = SEQ - 1
    dec.positions.fill_(pos)
    dec.tokens.fill_(1)

    # ---- device-resident arm: inputs already in HBM ------------------------------------------------
    for _ in range(warm):
        dec.step()
    sampler = ClockSampler(local) if rank == 0 else None
    t_dev = timed(dec.step, steps, dist_on)

    # ---- end-to-end arm: host token in -> H2D -> step -> D2H logits, every step -----------------------
    tok_host = torch.ones(1, dtype=torch.int32).pin_memory()
    pos_host = torch.full((1, ), pos, dtype=torch.int32).pin_memory()
    logits_host = torch.empty(1, dec.vocab, dtype=torch.float16).pin_memory()

    def e2e_step():
        dec.tokens.copy_(tok_host, non_blocking=True)
        dec.positions.copy_(pos_host, non_blocking=True)
        dec.step()
        logits_host.copy_(dec.logits, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the caller needs the logits before it can pick the next token
        tok_host[0] = int(logits_host[0, :8].float().argmax())  # touch the result on the host

    for _ in range(warm):
        e2e_step()
    t_e2e = timed(e2e_step, steps, dist_on)
    clocks = sampler.stop() if sampler else None

    # ---- dominant kernel in isolation: fused gate/up matvec over the 32 layers' distinct weights ------
    x = torch.randn(1, dec.hidden, device=dev).half()
    gates = [(ly['gate'], ly['up']) for ly in dec.layers]

    def mlp_all():
        for g, u in gates:
            ops.fused_mlp(x, (g.qweight, g.scales, g.qzeros, g.g_idx), (u.qweight, u.scales, u.qzeros, u.g_idx), BITS, GROUP)

    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        mlp_all()
        side.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            mlp_all()
    torch.cuda.synchronize()
    for _ in range(3):
        graph.replay()
    reps = 10
    t_k = timed(graph.replay, reps, False) / (reps * len(gates))
    kbytes = 2 * alg_bytes_qlinear(dec.hidden, dec.intermediate) - 2 * dec.hidden  # two weights, x read once
    peak, peak_src = measured_peaks()
    achieved = kbytes / t_k / 1e9

    # max over ranks, whole-job aggregate
    tt = torch.tensor([t_dev, t_e2e], device=dev, dtype=torch.float64)
    if dist_on:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_dev, t_e2e = tt.tolist()
    if rank == 0:
        base, _ = cpu_baseline(5) if world == 1 else (None, None)
        line = {
            'metric': METRIC, 'value': world * steps / t_dev, 'unit': 'tokens/s', 'n_gpus': world, 'steps': steps, 'warmup': warm,
            'ms_per_step': t_dev / steps * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f16', 'data': 'synthetic',
            'config': {
                'workload': f'LLaMA-7B int4 g128 batch=1 decode, context {pos} (seq={SEQ}), 32 layers, random-init packed weights',
                'parallelism': 'replicas only (one independent sequence per GPU, no data-path collective)' if world > 1 else 'single GPU',
                'l2': 'each step streams 3.6 GB of weights + 1.07 GB of KV cache (inputs >> 126 MB L2); no explicit flush needed',
                'arithmetic': 'int4 weights dequantised to fp16 exactly as the reference kernel, fp16 x fp16 -> fp32 accumulate, fp16 store',
            },
            'e2e': {'value': world * steps / t_e2e, 'unit': 'tokens/s', 'h2d_bytes_per_step': 8, 'd2h_bytes_per_step': dec.vocab * 2,
                    'note': 'host token+position (pinned) -> H2D -> CUDA-graph decode step -> D2H fp16 logits, synchronised every step'},
            'gpu_launches': dec.launches_per_step() * steps,
            'roofline': None,
            'clocks': clocks,
        }
        # algorithmic bytes per token: 32 x quant linears + fp16 lm_head + KV cache read at this context
        H, I, V = dec.hidden, dec.intermediate, dec.vocab
        per_layer = alg_bytes_qlinear(H, 3 * H) + alg_bytes_qlinear(H, H) + 2 * alg_bytes_qlinear(H, I) + alg_bytes_qlinear(I, H)
        kv = 2 * 32 * SEQ * H * 2
        step_bytes = 32 * per_layer + V * H * 2 + kv
        t_step = t_dev / steps
        mlp = {'kernel': 'qmatvec_int4_kernel<dual> (standalone gptq_fused_mlp_fwd, 4096->11008 x2, timed alone over 32 distinct layers)',
               'achieved': achieved, 'frac': achieved / peak, 'bytes_per_launch': kbytes, 'us_per_launch': t_k * 1e6}
        if dec.launches_per_step() == 1:
            # the whole token is ONE persistent kernel: its launch duration is the step time measured above with CUDA events
            line['roofline'] = {'bound': 'hbm', 'kernel': 'llama_decode_mega_kernel (persistent decode step: 160 int4 matvecs + attention + lm_head)',
                                'achieved': step_bytes / t_step / 1e9, 'peak': peak, 'unit': 'GB/s', 'frac': step_bytes / t_step / 1e9 / peak,
                                'peak_source': peak_src, 'bytes_per_launch': step_bytes, 'us_per_launch': t_step * 1e6,
                                'frac_of_8TBs': step_bytes / t_step / 8e12,
                                'traffic': NCU_TRAFFIC_BYTES, 'traffic_source': 'dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full, profiles/',
                                'standalone_fused_mlp': mlp}
        else:
            line['roofline'] = {'bound': 'hbm', 'kernel': mlp['kernel'], 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                                'peak_source': peak_src, 'bytes_per_launch': kbytes, 'us_per_launch': t_k * 1e6, 'frac_of_8TBs': achieved / 8000.0,
                                'traffic': None, 'step_bytes': step_bytes, 'step_achieved_gbs': step_bytes / t_step / 1e9,
                                'step_frac': step_bytes / t_step / 1e9 / peak}
        if base is not None:
            line['cpu_baseline'] = base
        print(json.dumps(line))
    if dist_on:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
