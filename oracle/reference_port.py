"""TEST INFRASTRUCTURE / CPU BASELINE ONLY -- the reference's OWN algorithm on the CPU.

oracle/fermiflow_oracle.py restates the path with the fixed-step solver the CUDA kernels
use.  This file restates how the reference itself computes one VMC iteration, so that
bench.py can time "the reference's CPU PyTorch path" on the GPU box, where
/root/reference does not exist:

  * adaptive Dormand-Prince integration, rtol 1e-6 / atol 1e-8 (nnModule.py:151-152
    defaults, via oracle/torchdiffeq_shim),
  * gradients by the continuous adjoint: the backward of a solve is another solve of the
    augmented system (state, adjoint, parameter adjoint) backwards in time, itself
    differentiable again (nnModule.py:78-149),
  * grad and Laplacian of log p by 1 + 2N nested autograd passes (utils.py:44-65),
  * Metropolis sampling of the base state with 100 whole-configuration moves
    (base_dist.py:58-70),
  * the energy gradient of VMC.py:41-61.

Pinned against the real reference in tests/test_oracle_pin.py::test_reference_port_*
(golden "default_*" arrays were produced by /root/reference itself with the same shim).
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "torchdiffeq_shim"))
from torchdiffeq import odeint  # noqa: E402  (the shim)

from . import fermiflow_oracle as O  # noqa: E402

torch.set_default_dtype(torch.float64)


class _AdjointSolve(torch.autograd.Function):
    """y(t1) for dy/dt = rhs(y, params), differentiable to any order by re-entrant
    adjoint solves.  args: rhs, t0, t1, n_state, rtol, atol, *state, *params."""

    @staticmethod
    def forward(ctx, rhs, t0, t1, n_state, rtol, atol, *tensors):
        state, params = tensors[:n_state], tensors[n_state:]
        with torch.no_grad():
            sol = odeint(lambda t, y: rhs(y, params), tuple(state), torch.tensor([t0, t1]), rtol=rtol, atol=atol)
            end = tuple(s[-1] for s in sol)
        ctx.meta = (rhs, t0, t1, n_state, rtol, atol, len(params))
        ctx.save_for_backward(*end, *params)
        return end

    @staticmethod
    def backward(ctx, *g_end):
        rhs, t0, t1, n, rtol, atol, n_par = ctx.meta
        saved = ctx.saved_tensors
        end, params = saved[:n], saved[n:]

        def augmented(aug, par):
            y, a = aug[:n], aug[n:2 * n]
            with torch.enable_grad():
                y = tuple(v if v.requires_grad else v.detach().requires_grad_(True) for v in y)
                par_l = tuple(p if p.requires_grad else p.detach().requires_grad_(True) for p in par)
                fy = rhs(y, par_l)
                inner = -sum((ai * fi).sum() for ai, fi in zip(a, fy))
                grads = torch.autograd.grad(inner, y + par_l, create_graph=True, allow_unused=True)
            gy = tuple(g if g is not None else torch.zeros_like(v) for g, v in zip(grads[:n], y))
            gp = tuple(g if g is not None else torch.zeros_like(p) for g, p in zip(grads[n:], par_l))
            return fy + gy + gp

        g_end = tuple(g if g is not None else torch.zeros_like(e) for g, e in zip(g_end, end))
        aug0 = end + g_end + tuple(torch.zeros_like(p) for p in params)
        out = _AdjointSolve.apply(augmented, t1, t0, len(aug0), rtol, atol, *aug0, *params)
        a0, pbar = out[n:2 * n], out[2 * n:]
        return (None, None, None, None, None, None, *a0, *pbar)


def solve(rhs, state, params, t0, t1, rtol=1e-6, atol=1e-8):
    return _AdjointSolve.apply(rhs, t0, t1, len(state), rtol, atol, *state, *params)


def _split(params, has_mu):
    return tuple(params[:3]), (tuple(params[3:6]) if has_mu else None)


def _rhs_v(has_mu):
    def rhs(y, par):
        eta, mu = _split(par, has_mu)
        return (O.backflow_v(y[0], eta, mu),)
    return rhs


def _rhs_v_div(has_mu):
    def rhs(y, par):
        eta, mu = _split(par, has_mu)
        return (O.backflow_v(y[0], eta, mu), -O.backflow_div(y[0], eta, mu))
    return rhs


def generate(z, params, has_mu, t_span, **tol):
    """flow.py:42-50 (params_require_grad=False: the parameters are closed over)."""
    fixed = tuple(p.detach() for p in params)
    with torch.no_grad():
        return _AdjointSolve.apply(lambda y, par: _rhs_v(has_mu)(y, fixed), t_span[0], t_span[1], 1,
                                   tol.get("rtol", 1e-6), tol.get("atol", 1e-8), z)[0]


def logp(x, orb_up, orb_dn, params, has_mu, t_span, params_require_grad, **tol):
    """VMC.py:36-39 through flow.py:52-56."""
    rtol, atol = tol.get("rtol", 1e-6), tol.get("atol", 1e-8)
    zeros = torch.zeros(x.shape[0])
    if params_require_grad:
        z, dl = _AdjointSolve.apply(_rhs_v_div(has_mu), t_span[1], t_span[0], 2, rtol, atol, x, zeros, *params)
    else:
        fixed = tuple(p.detach() for p in params)
        z, dl = _AdjointSolve.apply(lambda y, par: _rhs_v_div(has_mu)(y, fixed), t_span[1], t_span[0], 2,
                                    rtol, atol, x, zeros)
    return O.free_fermion_logp(orb_up, orb_dn, z) - dl


def y_grad_laplacian(f, x):
    """utils.py:44-65."""
    xf = x.flatten(1)
    y = f(xf.view_as(x))
    g, = torch.autograd.grad(y, xf, grad_outputs=torch.ones(x.shape[0]), create_graph=True)
    lap = sum(torch.autograd.grad(g[:, i], xf, grad_outputs=torch.ones(x.shape[0]), retain_graph=True)[0][:, i]
              for i in range(xf.shape[1]))
    return y, g.view_as(x), lap


def metropolis(orb_up, orb_dn, batch, steps=100, tau=0.1):
    """base_dist.py:58-70."""
    n = len(orb_up) + len(orb_dn)
    x = torch.randn(batch, n, 2)
    lp = O.free_fermion_logp(orb_up, orb_dn, x)
    for _ in range(steps):
        nx = x + tau * torch.randn_like(x)
        nlp = O.free_fermion_logp(orb_up, orb_dn, nx)
        acc = torch.rand_like(lp) < torch.exp(nlp - lp)
        x[acc] = nx[acc]
        lp[acc] = nlp[acc]
    return x


def vmc_iteration(nup, ndown, params, has_mu, Z, batch, t_span=(0.0, 1.0), equil=100, **tol):
    """GSVMC.forward + backward (VMC.py:41-61, FermionHO2D.py:66-72).  Returns E, E_std and
    the parameter gradients."""
    up, dn = list(range(nup)), list(range(ndown))
    z = metropolis(up, dn, batch, steps=equil)
    x = generate(z, params, has_mu, t_span, **tol).detach().requires_grad_(True)
    leaves = [p.detach().clone().requires_grad_(True) for p in params]
    logp_full = logp(x, up, dn, leaves, has_mu, t_span, True, **tol)
    lp, g, lap = y_grad_laplacian(lambda t: logp(t, up, dn, params, has_mu, t_span, False, **tol), x)
    kinetic = -0.25 * lap - 0.125 * (g ** 2).sum(dim=(-2, -1))
    eloc = (kinetic + O.potential_coulomb(x, Z) + O.potential_ho(x)).detach()
    E, E_std = eloc.mean().item(), eloc.std().item()
    gradE = (logp_full * (eloc - E)).mean()
    grads = torch.autograd.grad(gradE, leaves)
    return E, E_std, grads


def make_params(hidden, seed, scale=1e-2):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(2):
        out += [torch.randn(hidden, generator=g), torch.randn(hidden, generator=g), scale * torch.randn(hidden, generator=g)]
    return out


def time_vmc_iteration(nup, ndown, hidden, Z, walkers, seed=0, equil=100):
    torch.manual_seed(seed)
    params = make_params(hidden, 42)
    t0 = time.time()
    vmc_iteration(nup, ndown, params, True, Z, walkers, equil=equil)
    return time.time() - t0
