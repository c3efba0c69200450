"""Derivative helpers -- the role of reference src/utils.py on the CUDA path."""
import ctypes as C

import torch

from . import _lib as L


class ElocResult:
    __slots__ = ("z", "delta_logp", "logp", "grad", "lap", "kinetic", "potential", "eloc", "stash", "model")

    def without_stash(self):
        """The per-walker results only (what `model.last` keeps): the adjoint stash belongs to the autograd graph."""
        r = ElocResult()
        for k in self.__slots__:
            setattr(r, k, getattr(self, k))
        r.stash = None
        return r


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


def y_grad_laplacian(f, x):
    """utils.py:44-65: batch-wise value, gradient and Laplacian of the scalar function f at x.

    x: (batch, ...), f(x): (batch,); returns y (batch,), grad_y like x, laplacian_y (batch,).

    * f = a bound `GSVMC.logp` / `BetaVMC.logp`: ONE forward-mode sweep through the flow (C ABI ff_eloc) instead
      of the reference's 1 + 2N adjoint solves -- the second derivatives of the flow are not available through
      autograd here (`CNF.delta_logp` is once-differentiable);
    * f = functools.partial(FreeFermion.log_prob, orbitals_up, orbitals_down): one fused Slater launch
      (ff_free_fermion_logp_lap);
    * anything else (e.g. `lambda x: LogAbsSlaterDet.apply(orbitals, x)`, reference tests/test_slater.py:65-127):
      the reference's own loop of dim + 1 autograd passes; the Slater primitives of this package are twice
      differentiable (Hessian-vector product kernel ff_slater_hvp)."""
    import functools
    from .VMC import GSVMC, BetaVMC
    from .base_dist import FreeFermion
    owner, func = getattr(f, "__self__", None), getattr(f, "__func__", None)
    if isinstance(owner, GSVMC) and func is GSVMC.logp:
        r = owner.local_energy(x)
        return r.logp, r.grad, r.lap
    if isinstance(owner, BetaVMC) and func is BetaVMC.logp:
        r = eloc_sweep(owner.cnf, x, owner._state_table(x.device), owner.state_indices, owner.nup, owner._Z,
                       owner._harmonic)
        return r.logp, r.grad, r.lap
    if isinstance(f, functools.partial) and isinstance(getattr(f.func, "__self__", None), FreeFermion) \
            and getattr(f.func, "__func__", None) is FreeFermion.log_prob and len(f.args) == 2 and not f.keywords:
        return f.func.__self__.log_prob_grad_laplacian(f.args[0], f.args[1], x)
    if not x.requires_grad:
        x = x.detach().requires_grad_(True)
    x_flatten = x.flatten(start_dim=1)
    y = f(x_flatten.view_as(x))
    batch, dim = x_flatten.shape
    ones = torch.ones(batch, device=x.device, dtype=y.dtype)
    grad_y_flatten, = torch.autograd.grad(y, x_flatten, grad_outputs=ones, create_graph=True)
    laplacian_y = sum(torch.autograd.grad(grad_y_flatten[:, i], x_flatten, grad_outputs=ones, retain_graph=True)[0][:, i]
                      for i in range(dim))
    return y, grad_y_flatten.view_as(x), laplacian_y
