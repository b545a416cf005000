"""A few decode steps without a CUDA graph, for `ncu -k regex:llama_decode_mega` captures."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'gptq-for-llama_b200'))
from gptq_b200 import engine
size = sys.argv[1] if len(sys.argv) > 1 else '7b'
nl = int(sys.argv[2]) if len(sys.argv) > 2 else None
dec = engine.synthetic_llama(size, max_seq=2048, use_graph=False, n_layers=nl)
dec.k_cache.normal_(0, 0.5); dec.v_cache.normal_(0, 0.5)
dec.positions.fill_(2047); dec.tokens.fill_(1)
for _ in range(4):
    dec.step()
torch.cuda.synchronize()
print('done', bool(torch.isfinite(dec.logits).all()))
