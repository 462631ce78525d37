"""Stand-alone operators of the hot path (thin wrappers over the C ABI)."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib as L
from .modules import QNetwork, _ArenaModule, _stream


def soft_update(target, source, tau):
    """vn/utils.py:31-33 over all parameters, one fused kernel over the flat arenas."""
    if isinstance(target, QNetwork) and isinstance(source, QNetwork):
        target.bind(); source.bind()
        n = target.layout().total
        assert n == source.layout().total
        L.check(L.lib().dgvit_polyak_flat(target._arena.data_ptr(), source._arena.data_ptr(), n, float(tau),
                                          _stream(target._arena.device)), "polyak_flat")
        return
    if not (isinstance(target, _ArenaModule) and isinstance(source, _ArenaModule)):
        raise TypeError("soft_update expects dgvit_b200 modules")
    t, s = target.net_struct(), source.net_struct()
    L.check(L.lib().dgvit_polyak(C.byref(t), C.byref(s), float(tau), _stream(target._arena.device)), "polyak")


def hard_update(target, source):
    """vn/utils.py:35-37."""
    soft_update(target, source, 1.0)


_scratch = {}


def depth_augment(raw: torch.Tensor, noise: Optional[torch.Tensor] = None,
                  rng_state: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Depth normalise -> +N(0,50) -> clip -> 5x5 blur -> 11x11 centre-band blur -> 4x bilinear
    resize -> /255 (vn/env_lab.py:420-434,78-90,69-76,295-299).  raw [n,H,W] f32 on CUDA;
    ``noise`` [n,H,W] f32 N(0,50) draws, or ``rng_state`` (int64[2] on device) to generate them.  ``out``: optional
    contiguous f32 destination of n*(H/4)*(W/4) elements, e.g. consecutive rows of a ``ReplayStore`` (the states go
    straight into the store, no intermediate copy)."""
    if not raw.is_cuda:
        raise RuntimeError("depth_augment runs on CUDA only (no CPU fallback)")
    if raw.dim() == 2:
        raw = raw.unsqueeze(0)
    raw = raw.contiguous().float()
    n, H, W = raw.shape
    if noise is not None:
        noise = noise.to(raw.device).contiguous().float().reshape(n, H, W)
    elif rng_state is None:
        rng_state = torch.tensor([torch.initial_seed() & 0x7FFFFFFFFFFFFFFF, int(torch.randint(0, 2**31, (1,)))],
                                 dtype=torch.int64, device=raw.device)
    key = (raw.device, n, H, W)
    sc = _scratch.get(key)
    if sc is None:
        nb = C.c_size_t()
        L.check(L.lib().dgvit_depth_scratch_bytes(n, H, W, C.byref(nb)), "depth_scratch_bytes")
        sc = torch.empty(nb.value, dtype=torch.uint8, device=raw.device)
        if len(_scratch) > 16:
            _scratch.clear()
        _scratch[key] = sc
    if out is None:
        out = torch.empty(n, H // 4, W // 4, dtype=torch.float32, device=raw.device)
    elif not (out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and out.numel() == n * (H // 4) * (W // 4)):
        raise ValueError("out must be a contiguous CUDA float32 tensor of n*(H/4)*(W/4) elements")
    L.check(L.lib().dgvit_depth_augment(raw.data_ptr(), L.ptr(noise), L.ptr(rng_state), n, H, W, out.data_ptr(),
                                        sc.data_ptr(), sc.numel(), _stream(raw.device)), "depth_augment")
    return out
