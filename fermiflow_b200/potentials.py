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
    """Single-particle potential: subclasses give V(x) -> (batch,) (potentials.py:5-14).  `HO` is evaluated inside
    the fused E_loc sweep; any other subclass is evaluated through its own V (torch ops on the device) and added
    to the sweep's kinetic energy."""


class HO(SPPotential):
    def V(self, x):                                 # potentials.py:13-14
        return _potential(x, 0.0, True)


class PairPotential(object):
    """Pair potential V = sum_{i<j} v(|r_i - r_j|) (potentials.py:18-39): subclasses give v(rij).
    `CoulombPairPotential` is evaluated inside the fused E_loc sweep; any other subclass goes through rij / V below
    (torch ops on the device) and is added to the sweep's kinetic energy."""

    def rij(self, x):                               # potentials.py:23-31
        n = x.shape[-2]
        row, col = torch.triu_indices(n, n, offset=1, device=x.device)
        return (x[:, row] - x[:, col]).norm(dim=-1)

    def V(self, x):                                 # potentials.py:33-39
        return self.v(self.rij(x)).sum(dim=-1)


class CoulombPairPotential(PairPotential):
    def __init__(self, Z):
        self.Z = Z

    def v(self, rij):                               # potentials.py:45-46
        return self.Z / rij

    def V(self, x):                                 # potentials.py:33-39
        return _potential(x, self.Z, False)
