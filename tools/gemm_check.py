import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'gptq-for-llama_b200'))
from gptq_b200 import ops
from oracle import gptq_oracle as O
for (K, N, M) in [(4096, 4096, 512), (256, 128, 128), (1024, 256, 64)]:
    qw, s, qz, g, _ = [t.cuda() if t is not None else None for t in O.random_packed(K, N, 4, 128, seed=3)]
    x = torch.randn(M, K, generator=torch.Generator().manual_seed(0)).half().cuda()
    f = lambda v: ops.matmul248(v, qw, s, qz, g, 4, 15, groupsize=128)
    a, b, c = f(x), f(x), f(x * 2)
    torch.cuda.synchronize()
    print(K, N, M, 'nondeterministic elems', int((a != b).sum()), 'homogeneity mismatches', int((c != a * 2).sum()),
          'max |c-2a|', float((c.float() - 2 * a.float()).abs().max()), 'max |a|', float(a.float().abs().max()))
    W = ops.dequant(qw, s, qz, g, 4, 128)
    ref = (x.float() @ W.float())
    print('   max |a-ref|', float((a.float() - ref).abs().max()), ' max |c/2-ref|', float((c.float() / 2 - ref).abs().max()))
