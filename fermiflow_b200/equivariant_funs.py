"""Permutation-equivariant backflow velocity field -- mirror of reference
src/equivariant_funs.py:4-102, evaluated by the CUDA kernel `backflow_kernel`
(C ABI ff_backflow).  dim = 2 only (the orbitals of this path are 2D)."""
import ctypes as C

import torch

from . import _lib as L


class Backflow(torch.nn.Module):
    def __init__(self, eta, mu=None):
        super().__init__()
        self.eta = eta
        self.mu = mu

    def _model(self, n, t_span=(0.0, 1.0), nsteps=1, n_up=None):
        eta = self.eta.kernel_params()
        mu = self.mu.kernel_params() if self.mu is not None else None
        n_up = n if n_up is None else n_up
        self._keep = (eta, mu)
        return L.make_model(n_up, n - n_up, eta, mu, t_span, nsteps)

    def _run(self, x, want_v, want_div):
        if x.dim() != 3 or x.shape[-1] != 2:
            raise ValueError("Backflow expects x of shape (batch, n, 2)")
        x = x.detach().contiguous()
        B, n, _ = x.shape
        m = self._model(n)
        v = torch.empty_like(x) if want_v else None
        div = torch.empty(B, dtype=x.dtype, device=x.device) if want_div else None
        L.check(L.lib().ff_backflow(C.byref(m), L.ptr(x), B, L.ptr(v), L.ptr(div), L.stream()))
        return v, div

    def forward(self, x):                     # equivariant_funs.py:80-89
        return self._run(x, True, False)[0]

    def divergence(self, x):                  # equivariant_funs.py:91-102
        return self._run(x, False, True)[1]
