import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'gptq-for-llama_b200'))
from gptq_b200 import ops
from oracle import gptq_oracle as O
for (K, N, M, gs) in [(256, 128, 6, 64), (4096, 4096, 1, 128), (512, 256, 1, 128), (4096, 11008, 1, 128)]:
    qw, s, qz, g, _ = O.random_packed(K, N, 4, gs, seed=1)
    x = torch.randn(M, K, generator=torch.Generator().manual_seed(0)).half()
    ref = O.qlinear_fwd(x, qw, s, qz, g, 4)
    out = ops.matmul248(x.cuda(), qw.cuda(), s.cuda(), qz.cuda(), g.cuda(), 4, 15, groupsize=gs)
    torch.cuda.synchronize()
    err = (out.float().cpu() - ref.float()).abs().max().item()
    print(K, N, M, 'max err', err, 'rms', ref.float().pow(2).mean().sqrt().item(), flush=True)
