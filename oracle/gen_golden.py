"""TEST INFRASTRUCTURE ONLY -- makes tests/golden/*.npz by running the REAL reference
(/root/reference/src, imported read-only) in the build container.

The reference imports `torchdiffeq`, which is absent; oracle/torchdiffeq_shim stands in
(see its header).  Run:  python oracle/gen_golden.py   (about 2-4 minutes on 8 cores).
Nothing here is used at run time on the GPU box; only the .npz files travel.
"""
import contextlib
import io
import os
import sys
from collections import Counter

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("FERMIFLOW_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(HERE, "torchdiffeq_shim"), os.path.join(REF, "src")]
OUT = os.path.join(HERE, "..", "tests", "golden")

torch.set_default_dtype(torch.float64)

import NeuralODE.nnModule as nnmod            # noqa: E402
import torchdiffeq                            # noqa: E402  (the shim)
from MLP import MLP                           # noqa: E402
from base_dist import FreeFermion             # noqa: E402
from equivariant_funs import Backflow         # noqa: E402
from flow import CNF                          # noqa: E402
from orbitals import HO2D                     # noqa: E402
from potentials import HO, CoulombPairPotential  # noqa: E402
from slater import LogAbsSlaterDet, LogAbsSlaterDetMultStates  # noqa: E402
from utils import y_grad_laplacian            # noqa: E402
from VMC import GSVMC                         # noqa: E402


def _e_e_divergence_index_first(self, x):
    """equivariant_funs.py:33-47 with ONE reordering: the reference takes the norm of the
    full (n, n) difference tensor, zero diagonal included, and only then keeps the i<j
    entries (lines 44-45).  Under torch >= 2 the double backward of norm at the discarded
    zero vectors is 0 * inf = NaN, so the reference's Laplacian is NaN in this image.
    Selecting i<j before the norm (what potentials.py:31 already does) is the same
    function and differentiates cleanly.  Applied only while generating fixtures."""
    _, n, dim = x.shape
    row, col = torch.triu_indices(n, n, offset=1)
    dij = (x[:, :, None] - x[:, None])[:, row, col, :].norm(dim=-1, keepdim=True)
    return 2 * (self.eta.grad(dij) * dij + dim * self.eta(dij)).sum(dim=(-2, -1))


Backflow._e_e_divergence = _e_e_divergence_index_first


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def set_solver(**kw):
    """Route the reference's odeint call (nnModule.py:69) through fixed kwargs."""
    def odeint(f, y0, t, rtol=None, atol=None):
        args = dict(rtol=rtol, atol=atol)
        args.update(kw)
        return torchdiffeq.odeint(f, y0, t, **args)
    nnmod.odeint = odeint


def make_mlp(H, seed, scale):
    g = torch.Generator().manual_seed(seed)
    m = MLP(1, H)
    with torch.no_grad():
        m.fc1.weight.copy_(torch.randn(H, 1, generator=g))
        m.fc1.bias.copy_(torch.randn(H, generator=g))
        m.fc2.weight.copy_(scale * torch.randn(1, H, generator=g))
    return m


def mlp_arrays(m, prefix):
    return {prefix + "_w1": m.fc1.weight.detach().numpy()[:, 0].copy(),
            prefix + "_b1": m.fc1.bias.detach().numpy().copy(),
            prefix + "_w2": m.fc2.weight.detach().numpy()[0].copy()}


def np_(t):
    return t.detach().numpy().copy()


def main():
    os.makedirs(OUT, exist_ok=True)
    ho2d = HO2D()
    g = torch.Generator().manual_seed(20240607)

    # ---- 1. backflow velocity field and divergence (equivariant_funs.py) --------------
    eta, mu = make_mlp(8, 1, 0.3), make_mlp(6, 2, 0.2)
    v = Backflow(eta, mu=mu)
    x = torch.randn(6, 5, 2, generator=g)
    d = dict(x=np_(x), v=np_(v(x)), div=np_(v.divergence(x)),
             v_nomu=np_(Backflow(eta)(x)), div_nomu=np_(Backflow(eta).divergence(x)))
    d.update(mlp_arrays(eta, "eta")); d.update(mlp_arrays(mu, "mu"))
    np.savez(os.path.join(OUT, "backflow.npz"), **d)

    # ---- 2. orbitals + log|det| with gradient and laplacian (slater.py, orbitals.py) ---
    xo = torch.randn(7, 2, generator=g)
    orb_vals = torch.stack([o(xo) for o in ho2d.orbitals], -1)
    d = dict(x_orb=np_(xo), orbitals=np_(orb_vals), Es=np.array(ho2d.Es))
    for name, idx in (("gs6", list(range(6))), ("gs10", list(range(10))),
                      ("rand7", [0, 2, 3, 7, 11, 20, 35])):
        orbs = tuple(ho2d.orbitals[k] for k in idx)
        xs = torch.randn(5, len(idx), 2, generator=g, requires_grad=True)
        y, gy, ly = quiet(y_grad_laplacian, lambda t: LogAbsSlaterDet.apply(orbs, t), xs)
        d.update({name + "_idx": np.array(idx), name + "_x": np_(xs), name + "_logabsdet": np_(y),
                  name + "_grad": np_(gy), name + "_lap": np_(ly)})
    # two-spin log_prob (base_dist.py:48)
    ff = FreeFermion()
    xs = torch.randn(5, 5, 2, generator=g, requires_grad=True)
    y, gy, ly = quiet(y_grad_laplacian,
                      lambda t: ff.log_prob(ho2d.orbitals[:3], ho2d.orbitals[:2], t), xs)
    d.update(ff_x=np_(xs), ff_logp=np_(y), ff_grad=np_(gy), ff_lap=np_(ly))
    np.savez(os.path.join(OUT, "slater.npz"), **d)

    # ---- 3. finite-T states (orbitals.py:34) and multi-state determinants -------------
    d = {}
    for nup, dE in ((3, 2), (6, 2), (3, 4), (10, 2), (4, 3)):
        E0 = sum(ho2d.Es[:nup])
        idx, Es = ho2d.subsets(nup, E0 + dE, ho2d.Es)
        d["states_%d_%d" % (nup, dE)] = np.array(idx)
        d["Es_%d_%d" % (nup, dE)] = np.array(Es)
    idx, _ = ho2d.subsets(3, sum(ho2d.Es[:3]) + 2, ho2d.Es)
    states = tuple(tuple(ho2d.orbitals[k] for k in s) for s in idx)
    sidx = torch.randint(len(states), (12,), generator=g)
    coll = Counter(sorted(sidx.tolist()))
    xs = torch.randn(12, 3, 2, generator=g, requires_grad=True)
    y, gy, ly = quiet(y_grad_laplacian,
                      lambda t: LogAbsSlaterDetMultStates.apply(states, coll, t), xs)
    d.update(ms_state_idx=np.array(sorted(sidx.tolist())), ms_x=np_(xs), ms_logabsdet=np_(y),
             ms_grad=np_(gy), ms_lap=np_(ly))
    np.savez(os.path.join(OUT, "states.npz"), **d)

    # ---- 4. potentials ------------------------------------------------------------------
    xs = torch.randn(6, 5, 2, generator=g)
    np.savez(os.path.join(OUT, "potentials.npz"), x=np_(xs), ho=np_(HO().V(xs)),
             coulomb=np_(CoulombPairPotential(1.7).V(xs)), Z=1.7)

    # ---- 5. the whole path: CNF + Slater + E_loc + parameter gradient (VMC.py) ----------
    nup, ndown, batch, Z = 3, 2, 4, 2.0
    eta, mu = make_mlp(8, 11, 0.05), make_mlp(6, 12, 0.05)
    t_span = (0.0, 1.0)
    cnf = CNF(Backflow(eta, mu=mu), t_span)
    model = GSVMC(nup, ndown, ho2d, FreeFermion(), cnf, CoulombPairPotential(Z), sp_potential=HO())
    zs = 0.8 * torch.randn(batch, nup + ndown, 2, generator=g)
    wts = torch.randn(batch, generator=g) / batch
    d = dict(nup=nup, ndown=ndown, Z=Z, t_span=np.array(t_span), z0=np_(zs), weights=np_(wts))
    d.update(mlp_arrays(eta, "eta")); d.update(mlp_arrays(mu, "mu"))

    def run(tag):
        xg = quiet(cnf.generate, zs)
        xr = xg.detach().clone().requires_grad_(True)
        z_back, dl = quiet(cnf.delta_logp, xr)
        lp, gl, ll = quiet(y_grad_laplacian, model.logp, xr)
        kin = -0.25 * ll - 0.125 * (gl ** 2).sum(dim=(-2, -1))
        pot = model.pair_potential.V(xr) + model.sp_potential.V(xr)
        for p in model.parameters():
            p.grad = None
        lp_full = quiet(model.logp, xr.detach(), params_require_grad=True)
        quiet((lp_full * wts).sum().backward)
        d.update({tag + "_x": np_(xg), tag + "_zback": np_(z_back), tag + "_delta_logp": np_(dl),
                  tag + "_logp": np_(lp), tag + "_grad": np_(gl), tag + "_lap": np_(ll),
                  tag + "_eloc": np_(kin + pot),
                  tag + "_g_eta_w1": np_(eta.fc1.weight.grad)[:, 0], tag + "_g_eta_b1": np_(eta.fc1.bias.grad),
                  tag + "_g_eta_w2": np_(eta.fc2.weight.grad)[0],
                  tag + "_g_mu_w1": np_(mu.fc1.weight.grad)[:, 0], tag + "_g_mu_b1": np_(mu.fc1.bias.grad),
                  tag + "_g_mu_w2": np_(mu.fc2.weight.grad)[0]})

    set_solver()                                  # reference defaults: dopri5 rtol 1e-6 atol 1e-8
    run("default")
    set_solver(rtol=1e-11, atol=1e-13)            # same adaptive solver, tight tolerance
    run("tight")
    set_solver(method="rk4", options=dict(step_size=1.0 / 16))   # fixed-step 3/8 RK4, 16 steps
    run("rk4s16")
    np.savez(os.path.join(OUT, "pipeline.npz"), **d)
    print("golden fixtures written to", os.path.normpath(OUT))


def main_n20():
    """The headline configuration (BASELINE.json configs[2]: N = 20, 10 up / 10 down, Deta = Dmu = 50) on the REAL
    reference: `python oracle/gen_golden.py n20` -> tests/golden/pipeline_n20.npz.
      * rk4s16_*: odeint(method="rk4", 16 steps) -- the discrete algorithm the CUDA sweeps implement: x, z, delta_logp,
        log p to rounding; the reference's continuous-adjoint gradient / nested-adjoint Laplacian / E_loc on the same
        grid (they differ from the exact discrete derivative at the O(h^4) level of the integrator);
      * tight_*: the reference's adaptive dopri5 at rtol 1e-9 (gradient, Laplacian, E_loc, parameter gradients)."""
    import time
    os.makedirs(OUT, exist_ok=True)
    ho2d = HO2D()
    g = torch.Generator().manual_seed(20241017)
    nup, ndown, batch, Z = 10, 10, 2, 2.0
    eta, mu = make_mlp(50, 21, 0.02), make_mlp(50, 22, 0.02)
    t_span = (0.0, 1.0)
    cnf = CNF(Backflow(eta, mu=mu), t_span)
    model = GSVMC(nup, ndown, ho2d, FreeFermion(), cnf, CoulombPairPotential(Z), sp_potential=HO())
    # walkers of |Psi_0|^2 (the reference's own Metropolis sampler, base_dist.py:58) rather than Gaussian points:
    # E_loc of the benchmark's magnitude, no accidental near-node configurations
    torch.manual_seed(20241017)
    zs = quiet(FreeFermion().sample, ho2d.orbitals[:nup], ho2d.orbitals[:ndown], (batch,))
    wts = torch.randn(batch, generator=g) / batch
    d = dict(nup=nup, ndown=ndown, Z=Z, t_span=np.array(t_span), z0=np_(zs), weights=np_(wts))
    d.update(mlp_arrays(eta, "eta")); d.update(mlp_arrays(mu, "mu"))

    def run(tag):
        t0 = time.time()
        xg = quiet(cnf.generate, zs)
        xr = xg.detach().clone().requires_grad_(True)
        z_back, dl = quiet(cnf.delta_logp, xr)
        lp, gl, ll = quiet(y_grad_laplacian, model.logp, xr)
        kin = -0.25 * ll - 0.125 * (gl ** 2).sum(dim=(-2, -1))
        pot = model.pair_potential.V(xr) + model.sp_potential.V(xr)
        for p in model.parameters():
            p.grad = None
        lp_full = quiet(model.logp, xr.detach(), params_require_grad=True)
        quiet((lp_full * wts).sum().backward)
        d.update({tag + "_x": np_(xg), tag + "_zback": np_(z_back), tag + "_delta_logp": np_(dl),
                  tag + "_logp": np_(lp), tag + "_grad": np_(gl), tag + "_lap": np_(ll),
                  tag + "_eloc": np_(kin + pot),
                  tag + "_g_eta_w1": np_(eta.fc1.weight.grad)[:, 0], tag + "_g_eta_b1": np_(eta.fc1.bias.grad),
                  tag + "_g_eta_w2": np_(eta.fc2.weight.grad)[0],
                  tag + "_g_mu_w1": np_(mu.fc1.weight.grad)[:, 0], tag + "_g_mu_b1": np_(mu.fc1.bias.grad),
                  tag + "_g_mu_w2": np_(mu.fc2.weight.grad)[0]})
        print(tag, "done in %.0f s" % (time.time() - t0), flush=True)

    set_solver(method="rk4", options=dict(step_size=1.0 / 16))
    run("rk4s16")
    set_solver(rtol=1e-9, atol=1e-11)
    run("tight")
    np.savez(os.path.join(OUT, "pipeline_n20.npz"), **d)
    print("N = 20 golden fixture written")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "n20":
        main_n20()
    else:
        main()
