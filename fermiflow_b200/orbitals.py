"""Single-particle orbitals -- mirror of reference src/orbitals.py.

Orbitals are identified by their index in HO2D.orbitals (orbitals.py:89: shells n = 0..7,
nx = 0..n, ny = n - nx); the CUDA kernels evaluate them by Hermite recursion.  Each entry
is an `Orbital` that still behaves like the reference's python callable (evaluates
phi(x) with torch ops) for inspection, but carries `.index` for the kernels.
"""
import math
import random

import torch


class Orbital:
    def __init__(self, index, nx, ny):
        self.index, self.nx, self.ny = index, nx, ny

    @staticmethod
    def _hermite(n, x):
        hm, h = torch.zeros_like(x), torch.ones_like(x)
        for k in range(n):
            hm, h = h, math.sqrt(2.0 / (k + 1)) * x * h - math.sqrt(k / (k + 1.0)) * hm
        return h

    def __call__(self, x):                      # orbitals.py:84-87
        return (1.0 / math.sqrt(math.pi)) * torch.exp(-0.5 * (x ** 2).sum(dim=-1)) \
            * self._hermite(self.nx, x[..., 0]) * self._hermite(self.ny, x[..., 1])

    def __repr__(self):
        return "Orbital(%d: nx=%d, ny=%d)" % (self.index, self.nx, self.ny)


class Orbitals(object):
    def fermion_states_random(self, n):         # orbitals.py:9-12
        orbitals, Es = zip(*random.sample(tuple(zip(self.orbitals, self.Es)), k=n))
        return orbitals, Es

    def subsets(self, k, Pmax, Ps):
        """All index subsets of size k with total price <= Pmax, sorted by price, ties in
        lexicographic order (orbitals.py:14-32; prices ascending)."""
        out, N = [], len(Ps)

        def rec(start, chosen, total):
            need = k - len(chosen)
            if need == 0:
                out.append((tuple(chosen), total))
                return
            for nxt in range(start, N - need + 1):
                if sum(Ps[nxt:nxt + need]) <= Pmax - total:
                    rec(nxt + 1, chosen + [nxt], total + Ps[nxt])
        rec(0, [], 0)
        out.sort(key=lambda it: it[1])
        indices, totals = zip(*out)
        return indices, totals

    def fermion_states(self, nup, ndown, deltaE):   # orbitals.py:34-64
        if ndown != 0:
            raise ValueError("Only the polarized case (i.e., ndown = 0) is allowed "
                             "in the present implementation.")
        E0 = sum(self.Es[:nup])
        indices, Es = self.subsets(nup, E0 + deltaE, self.Es)
        states = tuple((tuple(self.orbitals[idx] for idx in subset), ()) for subset in indices)
        return states, Es


class HO2D(Orbitals):
    """2D isotropic harmonic oscillator, h = -1/2 laplacian + 1/2 r^2 (orbitals.py:66-90)."""

    def __init__(self):
        quanta = [(nx, n - nx) for n in range(8) for nx in range(n + 1)]
        self.orbitals = [Orbital(i, nx, ny) for i, (nx, ny) in enumerate(quanta)]
        self.Es = [n + 1 for n in range(8) for nx in range(n + 1)]
        self.E_indices = lambda n: tuple(range(n * (n + 1) // 2, (n + 1) * (n + 2) // 2))


_INDEX_CACHE = {}


def orbital_indices(orbitals, device):
    """int32 device vector of HO2D indices for a tuple of Orbital objects.  The vectors are cached per (indices,
    device): building one is a blocking host-to-device copy, and the hot path asks for the same occupations at
    every iteration.  (Read-only: callers must not write into them.)"""
    try:
        idx = tuple(int(o.index) for o in orbitals)
    except AttributeError:
        raise TypeError("orbitals must come from fermiflow_b200.orbitals.HO2D (python callables "
                        "cannot be evaluated by the CUDA kernels)")
    key = (idx, str(torch.device(device)))
    t = _INDEX_CACHE.get(key)
    if t is None:
        if len(_INDEX_CACHE) > 4096:
            _INDEX_CACHE.clear()
        t = _INDEX_CACHE[key] = torch.tensor(idx, dtype=torch.int32, device=device)
    return t
