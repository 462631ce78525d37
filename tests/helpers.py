"""Shared helpers of the parity tests (seeded inputs come from oracle.init_params)."""
import os

import numpy as np
import torch

from oracle import dgvit_oracle as O
from oracle.init_params import reference_init, reference_sac_init, synthetic_batch, synthetic_noise  # noqa: F401

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SEED = 3407


def golden(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


def unpack_mask(bits, shape):
    n = int(np.prod(shape))
    return torch.from_numpy(np.unpackbits(bits)[:n].reshape(shape).astype(np.float32))


def relerr(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def load_params(module, params):
    """copy an oracle param dict into a dgvit_b200 module (in place, keeps arena aliasing)."""
    sd = module.state_dict()
    with torch.no_grad():
        for k, v in params.items():
            sd[k].copy_(v)
