"""Weight initialisers - reference graphsage/inits.py:15-25 (glorot, zeros)."""
import math

import torch

_GEN = {}


def _generator(device):
    key = str(device)
    if key not in _GEN:
        g = torch.Generator(device=device)
        g.manual_seed(123)              # the reference seeds TF with 123 (supervised_train.py:22)
        _GEN[key] = g
    return _GEN[key]


def manual_seed(seed, device="cuda"):
    _generator(torch.device(device)).manual_seed(int(seed))


def glorot(shape, name=None, device="cuda"):
    """U(-r, r), r = sqrt(6 / (fan_in + fan_out)), fp32 - reference graphsage/inits.py:15-19."""
    device = torch.device(device)
    r = math.sqrt(6.0 / (shape[0] + shape[1]))
    w = torch.empty(tuple(shape), dtype=torch.float32, device=device)
    w.uniform_(-r, r, generator=_generator(device))
    return w


def zeros(shape, name=None, device="cuda"):
    """reference graphsage/inits.py:22-25."""
    return torch.zeros(tuple(shape), dtype=torch.float32, device=torch.device(device))
