"""Derivative helpers -- the role of reference src/utils.py on the CUDA path."""
import ctypes as C

import torch

from . import _lib as L


class ElocResult:
    __slots__ = ("z", "delta_logp", "logp", "grad", "lap", "kinetic", "potential", "eloc", "stash", "model")


def eloc_sweep(cnf, x, orb, walker_state, n_up, Z, harmonic, stash=False):
    """One forward-mode sweep (C ABI ff_eloc): log p, grad log p, laplacian log p and the
    local energy at x -- replaces utils.py:44-65 y_grad_laplacian applied to VMC.logp."""
    x = x.detach().contiguous()
    B, n, _ = x.shape
    m = cnf._model(n, n_up=n_up)
    r = ElocResult()
    dev, dt = x.device, x.dtype
    r.z = torch.empty_like(x)
    r.grad = torch.empty_like(x)
    for k in ("delta_logp", "logp", "lap", "kinetic", "potential", "eloc"):
        setattr(r, k, torch.empty(B, dtype=dt, device=dev))
    r.stash = None
    if stash:
        from .flow import _Stash
        r.stash = _Stash(m, B, dev)
    r.model = m
    L.check(L.lib().ff_eloc(C.byref(m), L.ptr(x), B, L.ptr(orb, torch.int32),
                            L.ptr(walker_state, torch.int32) if walker_state is not None else None,
                            float(Z), int(harmonic), L.ptr(r.z), L.ptr(r.delta_logp), L.ptr(r.logp),
                            L.ptr(r.grad), L.ptr(r.lap), L.ptr(r.kinetic), L.ptr(r.potential), L.ptr(r.eloc),
                            L.ptr(r.stash.y) if stash else None, L.ptr(r.stash.c) if stash else None, L.stream()))
    return r
