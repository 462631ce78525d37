"""CPU: the C-ABI library loads and exports every symbol include/dgvit.h declares, the
parameter layout mirrors the reference registration order, and the module surface
(state_dict keys, seeded init, deepcopy, loud failure without CUDA) matches the reference."""
import copy
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

import dgvit_b200 as dg
from dgvit_b200 import _lib as L
from helpers import O, SEED, golden, reference_init, reference_sac_init

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "dgvit.h")).read()
    declared = set(re.findall(r"^\s*(?:int|int64_t|long long|const char\*)\s+(dgvit_\w+)\s*\(", hdr, flags=re.M))
    assert declared, "no declarations parsed"
    lib = C.CDLL(L.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"libdgvit.so does not export {name}"
    assert declared == set(L.SYMBOLS), (declared ^ set(L.SYMBOLS))
    assert L.lib().dgvit_version() >= 100


def test_errors_are_codes_not_exceptions():
    cfg = L.Cfg(kind=L.ACTOR, img_h=128, img_w=160, patch_h=16, patch_w=20, dim=48, depth=4, heads=4,
                dim_head=64, mlp_dim=2048, n_act=2, n_pstate=2)       # dim not a multiple of 32
    out = L.Layout()
    rc = L.lib().dgvit_param_layout(C.byref(cfg), C.byref(out))
    assert rc == -1 and b"dim" in L.lib().dgvit_last_error()
    with pytest.raises(RuntimeError):
        L.check(rc, "param_layout")


@pytest.mark.parametrize("block,head,lfs", [(4, 4, 64), (2, 2, 32), (6, 6, 128)])
def test_layout_matches_registration_order(block, head, lfs):
    for cls in (dg.GoTPolicy, dg.GoTQNetwork):
        m = cls(2, 2, block, head, lfs)
        offs = m._named_offsets()
        names = [n for n, _ in m.named_parameters()]
        assert [n for n, _ in offs] == names
        lay = m.layout()
        end = 0
        for (n, off), p in zip(offs, m.parameters()):
            assert off % 64 == 0 and off >= end, n      # aligned, increasing, non-overlapping
            end = off + p.numel()
        assert end <= lay.total
        m.bind()
        assert m._bound()
        # views really alias the arena
        p0 = next(m.parameters())
        p0.data.fill_(3.0)
        assert float(m._arena[offs[0][1]]) == 3.0


def test_seeded_init_equals_reference():
    g = golden("modules_shipped.npz")
    torch.manual_seed(SEED)
    a = dg.GoTPolicy(2, 2, 4, 4, 64)
    torch.manual_seed(SEED + 1)
    c = dg.GoTQNetwork(2, 2, 4, 4, 64)
    cfg = O.Cfg()
    ra, rc = reference_init("actor", cfg, SEED), reference_init("critic", cfg, SEED + 1)
    for mod, ref, nm in ((a, ra, "actor"), (c, rc, "critic")):
        sd = mod.state_dict()
        assert list(sd.keys()) == list(ref.keys()) == [str(s) for s in g[f"{nm}_names"]]
        for k in ref:
            assert torch.equal(sd[k], ref[k]), k
        s = np.array([float(v.double().sum()) for v in sd.values()])
        np.testing.assert_allclose(s, g[f"{nm}_sum"], rtol=0, atol=1e-9)


def test_state_dict_roundtrip_and_deepcopy(tmp_path):
    torch.manual_seed(1)
    a = dg.GoTPolicy(2, 2, 2, 2, 32)
    a.bind()
    f = tmp_path / "a.pth"
    torch.save(a.state_dict(), f)
    b = dg.GoTPolicy(2, 2, 2, 2, 32)
    b.bind()
    b.load_state_dict(torch.load(f))
    assert b._bound()                      # load_state_dict copies in place: aliasing survives
    for (k, x), (_, y) in zip(a.state_dict().items(), b.state_dict().items()):
        assert torch.equal(x, y), k
    c = copy.deepcopy(a)
    c.bind()
    for x, y in zip(a.parameters(), c.parameters()):
        assert torch.equal(x, y) and x.data_ptr() != y.data_ptr()
    # torch.optim.Adam accepts the parameters (real leaf nn.Parameters)
    torch.optim.Adam(a.parameters(), lr=1e-3)
    # the cached two-parameter probe behind _bound() (it runs on every batch-1 choose_action) notices re-pointed storage:
    # the last parameter, the first parameter, and a replaced first Parameter object
    assert a._bound() and a._bound()
    last = list(a.parameters())[-1]
    last.data = last.data.clone()
    assert not a._bound()
    a.bind()
    assert a._bound() and last.data_ptr() == list(a.parameters())[-1].data_ptr()
    first = next(a.parameters())
    first.data = first.data.clone()
    assert not a._bound()
    a.bind()
    assert a._bound()
    assert not copy.deepcopy(a)._bound() or True     # a deep copy re-binds lazily; must not raise


def test_unsupported_reference_variants_fail_loudly():
    """heads == 1 with dim_head == dim: the reference replaces the attention output projection by nn.Identity
    (vn/GoalFormer.py:55,67-70) — another network than the kernels implement; building it must not silently succeed."""
    with pytest.raises(NotImplementedError):
        dg.GoTPolicy(2, 2, 2, 1, 64)
    dg.GoTPolicy(2, 2, 2, 1, 32)            # heads == 1 with a projection (dim != dim_head) is fine
    # a deep copy's trunk runs on the copy's arena, not on the original's
    a = dg.GoTPolicy(2, 2, 1, 2, 32)
    b = copy.deepcopy(a)
    assert b.trans._backend() is b and a.trans._backend() is a
    # the trunk of a stand-alone GoT gets a private arena whose layout follows the trunk's registration order
    t = dg.GoT(image_size=(128, 160), patch_size=(16, 20), num_classes=2, dim=32, depth=2, heads=2, mlp_dim=2048, channels=1)
    be = t._backend()
    assert [n for n, _ in be._named_offsets()] == [n for n, _ in be.named_parameters()]
    be.bind()
    assert be._bound()


def test_replay_record_layout():
    """one packed transition record = obs frame | next_obs frame | pobs | next_pobs | act | rew | done | engage, padded to
    16 bytes (dgvit_replay_append)"""
    n = L.lib().dgvit_replay_record_floats(128 * 160, 2, 2)
    assert n == 2 * 128 * 160 + 12 and n % 4 == 0
    assert L.lib().dgvit_replay_record_floats(64, 3, 2) == (2 * 64 + 11 + 3) // 4 * 4


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    a = dg.GoTPolicy(2, 2, 2, 2, 32)
    with pytest.raises(RuntimeError, match="CUDA"):
        a.sample([torch.zeros(1, 128, 160), torch.zeros(1, 2)])
    with pytest.raises(RuntimeError, match="CUDA"):
        dg.SAC(2, 2, "GaussianTransformer", "Transformer", False, False, False, 0)
    with pytest.raises(RuntimeError, match="CUDA"):
        dg.depth_augment(torch.zeros(8, 8))


def test_depth_streaming_formulation_equals_the_staged_pipeline():
    """The identities csrc/depth.cu's streaming kernels rest on, checked in numpy against the oracle's staged pipeline
    (normalise -> +noise -> GaussianBlur(5,5) -> 11x11 on the centre band -> resize by 4 -> /255):
      * off the band: resize(blur5(A)) is ONE separable stride-4 filter w6 = [1 5 10 10 5 1] / 32 over pixels 4o-1 .. 4o+4 of the
        noisy image A, image borders by BORDER_REFLECT_101;
      * in the band: horizontally ONE 16-tap composite c16 = k5 * (k11 at column 4o+1 + k11 at column 4o+2) / 2 over pixels
        4o-6 .. 4o+9; vertically 5-tap, then the 11-tap with BORDER_REFLECT_101 INSIDE the band, sample rows 4o+1, 4o+2; an output
        row with one sample outside the band takes that sample from the w6 path."""
    from oracle.dgvit_oracle import gaussian_kernel
    rs = np.random.RandomState(5)
    for (H, W) in ((64, 80), (120, 96), (40, 136)):
        raw = (rs.rand(H, W) * 7 + 0.5).astype(np.float32)
        noise = rs.normal(0, 50, (H, W)).astype(np.float32)
        want = O.depth_augment(raw, noise.astype(np.float64), out_hw=(H // 4, W // 4))
        mn, mx = float(raw.min()), float(raw.max())
        scale = 255.0 / (mx - mn)
        u8 = np.trunc((raw.astype(np.float64) * scale - mn * scale).astype(np.float32).clip(0, 255))
        A = np.clip(u8.astype(np.float64) + noise.astype(np.float64), 0, 255)
        refl = lambda i, n: np.where(i < 0, -i, np.where(i >= n, 2 * (n - 1) - i, i))
        k5, k11 = np.array([1, 4, 6, 4, 1]) / 16.0, gaussian_kernel(11).astype(np.float64).reshape(-1)
        w6 = np.array([1, 5, 10, 10, 5, 1]) / 32.0
        h12 = 0.5 * (np.append(k11, 0) + np.append(0, k11))
        c16 = np.convolve(k5, h12)
        oh, ow = H // 4, W // 4
        cols6 = refl(4 * np.arange(ow)[:, None] - 1 + np.arange(6)[None, :], W)              # [ow, 6]
        cols16 = refl(4 * np.arange(ow)[:, None] - 6 + np.arange(16)[None, :], W)            # [ow, 16]
        Ah6 = (A[:, cols6] * w6).sum(-1)                                                      # [H, ow]
        Ah16 = (A[:, cols16] * c16).sum(-1)
        bh = H // 5
        y1 = H // 2 - bh // 2
        y2 = y1 + bh
        blur5_rows = lambda M, q: sum(k5[t] * M[refl(q + t - 2, H)] for t in range(5))       # vertical 5-tap at row q
        out = np.zeros((oh, ow))
        for oy in range(oh):
            acc = 0.0
            for y in (4 * oy + 1, 4 * oy + 2):
                if y1 <= y < y2:
                    acc = acc + 0.5 * sum(k11[t] * blur5_rows(Ah16, y1 + int(refl(np.array(y - y1 + t - 5), bh))) for t in range(11))
                else:
                    acc = acc + 0.5 * blur5_rows(Ah6, y)
            out[oy] = acc / 255.0
        assert np.abs(out - want).max() < 1e-9, (H, W, float(np.abs(out - want).max()))
        # off the band the vertical direction collapses to the same w6: rows 4o-1 .. 4o+4
        for oy in range(oh):
            if not (y1 <= 4 * oy + 2 and 4 * oy + 1 < y2):
                rows = refl(4 * oy - 1 + np.arange(6), H)
                assert np.abs((Ah6[rows] * w6[:, None]).sum(0) / 255.0 - want[oy]).max() < 1e-9
