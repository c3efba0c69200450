"""Potentials -- mirror of reference src/potentials.py (C ABI ff_potential)."""
import torch

from . import _lib as L


def _potential(x, Z, harmonic):
    x = x.detach().contiguous()
    B, n, _ = x.shape
    V = torch.empty(B, dtype=x.dtype, device=x.device)
    L.check(L.lib().ff_potential(L.ptr(x), B, n, float(Z), int(harmonic), L.ptr(V), L.stream()))
    return V


class SPPotential(object):
    pass


class HO(SPPotential):
    def V(self, x):                                 # potentials.py:13-14
        return _potential(x, 0.0, True)


class PairPotential(object):
    pass


class CoulombPairPotential(PairPotential):
    def __init__(self, Z):
        self.Z = Z

    def v(self, rij):                               # potentials.py:45-46
        return self.Z / rij

    def V(self, x):                                 # potentials.py:33-39
        return _potential(x, self.Z, False)
