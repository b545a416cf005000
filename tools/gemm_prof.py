"""One large prefill GEMM for ncu: M=8192, 4096x4096 int4."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'gptq-for-llama_b200'))
from gptq_b200 import ops
sys.path.insert(0, os.path.join(ROOT, 'tools'))
from microbench import rand_layer
dev = torch.device('cuda:0')
w = rand_layer(4096, 4096, 4, 128, dev)
x = torch.randn(8192, 4096, device=dev).half()
for _ in range(3):
    ops.matmul248(x, *w, 4, None, groupsize=128)
torch.cuda.synchronize()
