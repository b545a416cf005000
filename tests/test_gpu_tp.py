"""GPU, 2 ranks, NCCL: tensor-parallel QuantLinear against the single-GPU result (needs >= 2 GPUs: gpurun --gpus 2)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import gptq_oracle as O

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        import quant
        from gptq_b200 import tp
        K, N, bits, gs = 2048, 1024, 4, 128
        qw, s, qz, g, _ = O.random_packed(K, N, bits, gs, seed=1)
        full = quant.QuantLinear(bits, gs, K, N, False)
        full.qweight, full.scales, full.qzeros, full.g_idx = qw, s, qz, g
        full = full.cuda()
        ok = True
        for M in (1, 4, 64):
            x = torch.randn(M, K, generator=torch.Generator().manual_seed(M)).half().cuda()
            ref = full(x)
            col = tp.TPQuantLinear(full, 'column', gather_output=True)
            row = tp.TPQuantLinear(full, 'row')
            ok = ok and torch.equal(col(x), ref)
            err = (row(x).float() - ref.float()).abs()
            bound = 2e-3 * torch.maximum(ref.float().abs(), ref.float().pow(2).mean().sqrt())
            ok = ok and bool((err <= bound).all())
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_tp_quantlinear_world2_nccl():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(2))
    assert res == {0: True, 1: True}
