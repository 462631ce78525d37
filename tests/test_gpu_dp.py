"""GPU, 2 ranks over NCCL (skipped with fewer than 2 GPUs): DP(2) of the fused update equals the
single-GPU update on the same global batch (fp32 path, injected noise)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from helpers import O, SEED, synthetic_batch, synthetic_noise

pytestmark = pytest.mark.gpu


def _agent(cfg, distributed, device, critic="Transformer"):
    import dgvit_b200 as dg
    return dg.SAC(2, 2, "GaussianTransformer", critic, False, False, False, SEED, LR_C=1e-3, LR_A=1e-3,
                  LR_ALPHA=1e-4, BUFFER_SIZE=16, TAU=5e-4, POLICY_FREQ=1, GAMMA=0.999, ALPHA=1.0, block=cfg.depth,
                  head=cfg.heads, l_f_size=cfg.dim, precision="fp32", device=device, distributed=distributed)


def _cuda(d, B):
    out = {}
    for k, v in d.items():
        if v is None:
            continue
        out[k] = (v.to(torch.uint8) if k.startswith("mask") else v.reshape(B, -1) if k in ("obs", "next_obs") else v).cuda().contiguous()
    return out


def _worker(rank, world, port, B, steps, out, critic="Transformer"):
    import torch.distributed as dist
    from dgvit_b200.parallel import shard_batch
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    cfg = O.Cfg(dim=32, depth=2, heads=2)
    ag = _agent(cfg, True, f"cuda:{rank}", critic)
    for s in range(steps):
        batch, noise = synthetic_batch(cfg, B, 100 + s), synthetic_noise(cfg, B, 200 + s)
        lb, ln = shard_batch(batch, world, rank), shard_batch(noise, world, rank)
        n = lb["obs"].shape[0]
        ag.update_from_batch(_cuda(lb, n), _cuda(ln, n), global_batch=B)
    torch.cuda.synchronize()
    if rank == 0:
        torch.save(dict(actor=ag.policy._arena.cpu(), critic=ag.critic._arena.cpu(), target=ag.critic_target._arena.cpu(),
                        losses=ag._loss_buffer().cpu(), log_alpha=ag.log_alpha.cpu()), out)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("critic", ["Transformer", "CNN"])
def test_dp2_equals_single_gpu(tmp_path, critic):
    B, steps = 8, 2
    out = str(tmp_path / "dp.pt")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, B, steps, out, critic), nprocs=2, join=True)
    got = torch.load(out)
    cfg = O.Cfg(dim=32, depth=2, heads=2)
    ag = _agent(cfg, False, "cuda:0", critic)
    for s in range(steps):
        batch, noise = synthetic_batch(cfg, B, 100 + s), synthetic_noise(cfg, B, 200 + s)
        ag.update_from_batch(_cuda(batch, B), _cuda(noise, B))
    torch.cuda.synchronize()
    for k, mod in (("actor", ag.policy), ("critic", ag.critic), ("target", ag.critic_target)):
        d = (got[k] - mod._arena.cpu()).abs()
        assert float((d > 1e-5).float().mean()) < 2e-3, k          # Adam amplifies rounding where |g| ~ eps
        assert float(d.max()) < 5e-3, (k, float(d.max()))
    assert torch.allclose(got["losses"][:2], ag._losses.cpu()[:2], rtol=1e-4, atol=1e-5)
    assert abs(float(got["log_alpha"]) - float(ag.log_alpha)) < 1e-6
