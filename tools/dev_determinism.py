"""Run-to-run comparison of the persistent kernel's intermediate buffers (development).
python tools/dev_determinism.py [n_layers] [pos]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'gptq-for-llama_b200')):
    sys.path.insert(0, p)
from gptq_b200 import engine
nl = int(sys.argv[1]) if len(sys.argv) > 1 else 1
pos = int(sys.argv[2]) if len(sys.argv) > 2 else 0
dec = engine.synthetic_llama('7b', bits=4, groupsize=128, vocab=32000, seed=13, max_seq=2048, n_layers=nl, use_graph=False)
dec.k_cache.normal_(0, 0.5); dec.v_cache.normal_(0, 0.5)
kc0, vc0 = dec.k_cache.clone(), dec.v_cache.clone()
H, I = dec.hidden, dec.intermediate
al = lambda v: (v + 255) // 256 * 256
base = dec.mega_scratch_offset()
off = {}
o = base
for name, nbytes in (('resid0', H * 2), ('resid1', H * 2), ('acc_qkv', 3 * H * 4), ('acc_o', H * 4), ('acc_d', H * 4), ('acc_g', I * 4), ('acc_u', I * 4), ('part', 1024 * 2 * 132 * 4)):
    off[name] = (o, nbytes); o += al(nbytes)
def snap():
    s = {}
    for n in ('resid0', 'resid1'):
        a, nb = off[n]; s[n] = dec.scratch[a:a + nb].view(torch.float16).float().clone()
    a, nb = off['part']; s['part'] = dec.scratch[a:a + 296 * 2 * 132 * 4].view(torch.float32).clone()
    s['k_new'] = dec.k_cache[:, 0, :, pos].float().clone(); s['v_new'] = dec.v_cache[:, 0, :, pos].float().clone()
    s['logits'] = dec.logits[0].float().clone()
    return s
runs = []
for i in range(12):
    dec.k_cache.copy_(kc0); dec.v_cache.copy_(vc0)
    dec.tokens.fill_(5); dec.positions.fill_(pos)
    dec.step(); torch.cuda.synchronize()
    runs.append(snap())
print('lib', os.environ.get('GPTQ_B200_LIB', 'default'), 'layers', nl, 'pos', pos)
for n in runs[0]:
    ref = runs[0][n]
    rms = ref[torch.isfinite(ref)].pow(2).mean().sqrt().item()
    d = [(r[n] - ref).abs() for r in runs[1:]]
    mx = max(x[torch.isfinite(x)].max().item() if torch.isfinite(x).any() else 0.0 for x in d)
    nd = max(int((x > 0).sum()) for x in d)
    print(f'  {n:8s} rms {rms:.3e}  max run-to-run diff {mx:.3e} ({mx / max(rms, 1e-30):.2e} of rms)  max #differing elements {nd} / {ref.numel()}')
