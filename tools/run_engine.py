"""Run a few decode steps of the synthetic 7B engine without CUDA graphs (for ncu)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'gptq-for-llama_b200'))
from gptq_b200 import engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dec = engine.synthetic_llama('7b', max_seq=2048, use_graph=False)
dec.k_cache.normal_(0, 0.5); dec.v_cache.normal_(0, 0.5)
dec.positions.fill_(2047); dec.tokens.fill_(1)
for _ in range(n):
    dec.step()
torch.cuda.synchronize()
print('launches/step', dec.launches_per_step())
