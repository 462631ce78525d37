"""CPU, world_size 2 over gloo: the data-parallel convention of the update.

Each rank computes the gradients of its shard with the losses normalised by the GLOBAL batch
(what the CUDA kernels do: `global_batch` in dgvit_sac), the flat gradient vectors are
SUM-all-reduced through dgvit_b200.parallel, and the result must equal the single-process
gradients on the whole batch.  The arithmetic is the oracle's (no GPU here); the sharding,
normalisation and reduction plumbing is the product's."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import O, reference_init, synthetic_batch, synthetic_noise


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _losses(pa, pc, batch, noise, cfg, Bg):
    """critic + policy loss of one shard, means taken over the global batch Bg."""
    s, ps, a, r = batch["obs"], batch["pobs"], batch["act"], batch["rew"]
    q1, q2 = O.critic_forward(pc, s, ps, a, cfg, noise["mask_c"])
    nq = r.expand(-1, 2) * 0.5
    qf = (((q1 - nq) ** 2).sum() + ((q2 - nq) ** 2).sum()) / (Bg * 2)
    pi, logp, _ = O.actor_sample(pa, s, ps, noise["eps_pi"], cfg, noise["mask_a"])
    pol = (0.7 * logp).sum() / Bg
    return qf, pol


def _worker(rank, world, port, B, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from dgvit_b200.parallel import allreduce_sum_, shard_batch, shard_bounds
    cfg = O.Cfg(dim=32, depth=1, heads=2)
    pa = {k: v.requires_grad_(True) for k, v in reference_init("actor", cfg, 3).items()}
    pc = {k: v.requires_grad_(True) for k, v in reference_init("critic", cfg, 4).items()}
    batch, noise = synthetic_batch(cfg, B, 5), synthetic_noise(cfg, B, 6)
    off, cnt = shard_bounds(B, world, rank)
    assert sum(shard_bounds(B, world, r)[1] for r in range(world)) == B
    qf, pol = _losses(pa, pc, shard_batch(batch, world, rank), shard_batch(noise, world, rank), cfg, B)
    qf.backward()
    pol.backward()
    flat_c = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in pc.values()])
    flat_a = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in pa.values()])
    allreduce_sum_(flat_c)
    allreduce_sum_(flat_a)
    if rank == 0:
        torch.save(dict(c=flat_c, a=flat_a, off=off, cnt=cnt), out)
    dist.destroy_process_group()


def test_dp2_matches_single_process(tmp_path):
    B, world = 6, 2
    out = str(tmp_path / "dp.pt")
    mp.spawn(_worker, args=(world, _free_port(), B, out), nprocs=world, join=True)
    got = torch.load(out)
    cfg = O.Cfg(dim=32, depth=1, heads=2)
    pa = {k: v.requires_grad_(True) for k, v in reference_init("actor", cfg, 3).items()}
    pc = {k: v.requires_grad_(True) for k, v in reference_init("critic", cfg, 4).items()}
    qf, pol = _losses(pa, pc, synthetic_batch(cfg, B, 5), synthetic_noise(cfg, B, 6), cfg, B)
    qf.backward()
    pol.backward()
    flat_c = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in pc.values()])
    flat_a = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in pa.values()])
    assert float((got["c"] - flat_c).abs().max()) <= 2e-5 * float(flat_c.abs().max())
    assert float((got["a"] - flat_a).abs().max()) <= 2e-5 * float(flat_a.abs().max())


def test_shard_bounds_cover_batch():
    from dgvit_b200.parallel import shard_bounds
    for B in (1, 7, 256, 4096):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(B, w, r) for r in range(w)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == B
            for (o1, c1), (o2, _) in zip(spans, spans[1:]):
                assert o1 + c1 == o2
