"""Tensor-parallel decode timing (BASELINE config 5): torchrun --nproc-per-node N tools/tp_bench.py [size] [steps]
One rank per GPU; LLaMA-65B int4 g128, batch 1, context 2047; device time per token (CUDA events, max over ranks)."""
import json, os, sys, torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'gptq-for-llama_b200'))
from gptq_b200 import engine
size = sys.argv[1] if len(sys.argv) > 1 else '65b'
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
    dec = engine.synthetic_llama_tp(size, rank, world, device=str(dev), seed=0, max_seq=2048)
else:
    dec = engine.synthetic_llama(size, device=str(dev), seed=0, max_seq=2048)
dec.k_cache.normal_(0, 0.5); dec.v_cache.normal_(0, 0.5)
dec.positions.fill_(2047); dec.tokens.fill_(1)
for _ in range(5): dec.step()
torch.cuda.synchronize()
if world > 1: dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(steps): dec.step()
b.record(); torch.cuda.synchronize()
t = torch.tensor([a.elapsed_time(b) / steps], device=dev, dtype=torch.float64)
if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    H, I, L, NH = engine.LLAMA_SHAPES[size]
    bytes_w = L * ((H * 3 * H + H * H + 3 * H * I) // 2 + (H // 128) * (3 * H + H + 2 * I) * 5 // 2 + (I // 128) * H * 5 // 2)
    bytes_tok = bytes_w + 32000 * H * 2 + 2 * L * 2048 * H * 2
    print(json.dumps({'config': f'LLaMA-{size} int4 g128 batch=1 decode, context 2047, tensor-parallel over {world} GPU(s)', 'n_gpus': world, 'ms_per_token': t.item(),
                      'tokens_per_s': 1e3 / t.item(), 'algorithmic_GB_per_token': bytes_tok / 1e9, 'aggregate_GBps': bytes_tok / t.item() / 1e6,
                      'finite': bool(torch.isfinite(dec.logits).all()), 'launches_per_step': dec.launches_per_step()}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
