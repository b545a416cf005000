"""Decode-only timing of the synthetic 7B engine (graph replay), for A/B builds: GPTQ_B200_LIB=<so> python tools/quick_bench.py"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'gptq-for-llama_b200'))
from gptq_b200 import engine
size = sys.argv[1] if len(sys.argv) > 1 else '7b'
bits = int(sys.argv[2]) if len(sys.argv) > 2 else 4
act = len(sys.argv) > 3 and sys.argv[3] == 'act'
dec = engine.synthetic_llama(size, bits=bits, act_order=act, max_seq=2048)
dec.k_cache.normal_(0, 0.5); dec.v_cache.normal_(0, 0.5)
dec.positions.fill_(2047); dec.tokens.fill_(1)
for _ in range(10): dec.step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(100): dec.step()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 100
print(f'bits={bits} act_order={act}', end=' ')
print(size, os.environ.get('GPTQ_B200_LIB', 'default'), f'{ms:.3f} ms/token  {1000/ms:.1f} tok/s  launches/step {dec.launches_per_step()}')
