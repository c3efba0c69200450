"""log|Slater determinant| primitives -- mirror of reference src/slater.py, computed by the
CUDA kernel `slater_kernel` (C ABI ff_slater_logabsdet): Hermite-recursion orbitals,
pivoted Gauss-Jordan inverse, log|det|, and the Jacobi-formula gradient in one launch.
"""
from collections import Counter

import torch

from . import _lib as L
from .orbitals import orbital_indices


def _run(x, orb, walker_state, want_grad, want_lap=False):
    shape = x.shape
    n = shape[-2]
    xf = x.detach().reshape(-1, n, 2).contiguous()
    B = xf.shape[0]
    out = torch.empty(B, dtype=xf.dtype, device=xf.device)
    grad = torch.empty_like(xf) if want_grad else None
    lap = torch.empty(B, dtype=xf.dtype, device=xf.device) if want_lap else None
    L.check(L.lib().ff_slater_logabsdet(L.ptr(xf), B, n, L.ptr(orb, torch.int32),
                                        L.ptr(walker_state, torch.int32) if walker_state is not None else None,
                                        L.ptr(out), L.ptr(grad), L.ptr(lap), L.stream()))
    return (out.reshape(shape[:-2]), grad.reshape(shape) if want_grad else None,
            lap.reshape(shape[:-2]) if want_lap else None)


class _SlaterGradient(torch.autograd.Function):
    """d(scale * (log|det_up| + log|det_dn|))/dx as a differentiable function of x: the forward hands out the
    gradient the forward kernel already produced, the backward is the Hessian-vector product kernel
    (C ABI ff_slater_hvp).  This is what makes the backward of the Slater primitives differentiable again, like the
    reference's (slater.py:40-60 builds it from torch ops for utils.py:44-65)."""

    @staticmethod
    def forward(ctx, x, dlog, orb, walker_state, n_up, n_dn, scale):
        ctx.save_for_backward(x)
        ctx.meta = (orb, walker_state, n_up, n_dn, scale)
        return dlog.clone()

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, v):
        x, = ctx.saved_tensors
        orb, ws, n_up, n_dn, scale = ctx.meta
        n = n_up + n_dn
        xf = x.detach().reshape(-1, n, 2).contiguous()
        vf = v.reshape(-1, n, 2).contiguous()
        hv = torch.empty_like(xf)
        L.check(L.lib().ff_slater_hvp(L.ptr(xf), xf.shape[0], n_up, n_dn, L.ptr(orb, torch.int32),
                                      L.ptr(ws, torch.int32) if ws is not None else None, float(scale),
                                      L.ptr(vf), L.ptr(hv), L.stream()))
        return hv.reshape(x.shape), None, None, None, None, None, None


class LogAbsSlaterDet(torch.autograd.Function):
    """slater.py:4-60.  forward(orbitals, x): x (*batch, n, 2) -> log|det| (*batch).  Twice differentiable."""

    @staticmethod
    def forward(ctx, orbitals, x):
        orb = orbital_indices(orbitals, x.device)
        out, grad, _ = _run(x, orb, None, ctx.needs_input_grad[1])
        if grad is not None:
            ctx.save_for_backward(x, grad)
            ctx.orb = orb
        return out

    @staticmethod
    def backward(ctx, grad_logabsdet):
        x, dlog = ctx.saved_tensors
        d = _SlaterGradient.apply(x, dlog, ctx.orb, None, x.shape[-2], 0, 1.0)
        return None, grad_logabsdet[..., None, None] * d


def logabsslaterdet(orbitals, x):                       # slater.py:62-68
    return LogAbsSlaterDet.apply(orbitals, x)


def slater_value_grad_laplacian(orbitals, x):
    """log|det|, gradient and Laplacian from one launch (what utils.y_grad_laplacian
    obtains from LogAbsSlaterDet by 2n extra autograd passes)."""
    return _run(x, orbital_indices(orbitals, x.device), None, True, True)


def states_table(states, device):
    """int32 (Nstates, n) table of HO2D indices for a tuple of states (each a tuple of
    Orbital objects)."""
    return torch.stack([orbital_indices(s, device) for s in states]).contiguous()


def walker_states_from_collection(collection, device):
    """state_indices_collection (dict / Counter index -> multiplicity, VMC.py:97) expanded
    to one int32 state index per walker, in the dict's iteration order."""
    idx = []
    for k, times in collection.items():
        idx += [int(k)] * int(times)
    return torch.tensor(idx, dtype=torch.int32, device=device)


class LogAbsSlaterDetMultStates(torch.autograd.Function):
    """slater.py:70-156: walker b uses the orbitals of states[state_of_walker[b]]."""

    @staticmethod
    def forward(ctx, states, state_indices_collection, x):
        table = states_table(states, x.device)
        ws = state_indices_collection if isinstance(state_indices_collection, torch.Tensor) \
            else walker_states_from_collection(state_indices_collection, x.device)
        if ws.numel() != x.shape[0]:
            raise ValueError("batch must equal the sum of the multiplicities in state_indices_collection")
        out, grad, _ = _run(x, table, ws, ctx.needs_input_grad[2])
        if grad is not None:
            ctx.save_for_backward(x, grad)
            ctx.table, ctx.ws = table, ws
        return out

    @staticmethod
    def backward(ctx, grad_logabsdet):
        x, dlog = ctx.saved_tensors
        d = _SlaterGradient.apply(x, dlog, ctx.table, ctx.ws, x.shape[-2], 0, 1.0)
        return None, None, grad_logabsdet[:, None, None] * d


def logabsslaterdetmultstates(states, state_indices_collection, x):   # slater.py:158-167
    return LogAbsSlaterDetMultStates.apply(states, state_indices_collection, x)


__all__ = ["LogAbsSlaterDet", "LogAbsSlaterDetMultStates", "logabsslaterdet",
           "logabsslaterdetmultstates", "slater_value_grad_laplacian", "Counter"]
