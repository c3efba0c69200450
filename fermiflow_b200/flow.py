"""Continuous normalizing flow of fermion coordinates -- mirror of reference src/flow.py
(CNF) and src/NeuralODE/nnModule.py (solve_ivp_nnmodule + adjoint), on fixed-step 3/8-rule
RK4 CUDA kernels (C ABI ff_cnf_generate / ff_cnf_delta_logp / ff_eloc / ff_logp_backward).
"""
import ctypes as C

import torch

from . import _lib as L


def _flat_params(v):
    ps = list(v.eta.parameters_in_kernel_order())
    if v.mu is not None:
        ps += list(v.mu.parameters_in_kernel_order())
    return ps


class _Stash:
    """Device buffers the backward sweep needs (what ctx.save_for_backward holds in
    nnModule.py:73)."""

    def __init__(self, model, B, device):
        ny, nc, nw = C.c_longlong(), C.c_longlong(), C.c_longlong()
        L.check(L.lib().ff_stash_sizes(C.byref(model), B, C.byref(ny), C.byref(nc)))
        L.check(L.lib().ff_backward_work_size(C.byref(model), B, C.byref(nw)))
        self.y = torch.empty(ny.value, dtype=torch.float64, device=device)
        # the per-item radial functions (f, f', f''): 16x the size of the stage inputs (22 GB for 65536 walkers at
        # n = 20).  _lib.STASH_RADIAL = False drops them: the backward sweep then recomputes them from the stage
        # inputs (Taylor tables), which costs 28 ms more per iteration at that size but 16x less memory.
        self.c = torch.empty(nc.value, dtype=torch.float64, device=device) if L.STASH_RADIAL else None
        self.n_work = nw.value


def _backward_through_flow(cnf, model, stash, B, gbar_z, gbar_delta, need_x):
    """ff_logp_backward: returns grad_x (or None) and the list of parameter gradients."""
    dev = gbar_z.device
    v = cnf.v
    params = _flat_params(v)
    sizes = [p.numel() for p in params]
    grads = list(torch.zeros(sum(sizes), dtype=torch.float64, device=dev).split(sizes))      # one fill launch
    gp = [L.ptr(g) for g in grads] + [None] * (6 - len(grads))
    grad_x = torch.empty_like(gbar_z) if need_x else None
    work = torch.empty(stash.n_work, dtype=torch.float64, device=dev)
    L.check(L.lib().ff_logp_backward(C.byref(model), B, L.ptr(stash.y), L.ptr(stash.c),
                                     L.ptr(gbar_z.contiguous()), L.ptr(gbar_delta.contiguous()),
                                     L.ptr(grad_x), *gp, L.ptr(work), L.stream()))
    return grad_x, [g.view_as(p) for g, p in zip(grads, params)]


class _DeltaLogp(torch.autograd.Function):
    """(z, delta_logp) = flow^{-1}(x) with a hand-written backward (nnModule.py:8-103)."""

    @staticmethod
    def forward(ctx, cnf, x, with_params, *params):
        x = x.detach().contiguous()
        B, n, _ = x.shape
        model = cnf._model(n)
        z = torch.empty_like(x)
        dl = torch.empty(B, dtype=x.dtype, device=x.device)
        need_bwd = any(ctx.needs_input_grad)
        stash = _Stash(model, B, x.device) if need_bwd else None
        L.check(L.lib().ff_cnf_delta_logp(C.byref(model), L.ptr(x), B, L.ptr(z), L.ptr(dl),
                                          L.ptr(stash.y) if stash else None,
                                          L.ptr(stash.c) if stash else None, L.stream()))
        ctx.cnf, ctx.model, ctx.stash, ctx.B = cnf, model, stash, B
        ctx.n_params = len(params)
        return z, dl

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gz, gdl):
        # first order only: the second derivatives of log p w.r.t. x come from the forward-mode sweep
        # (utils.y_grad_laplacian / eloc_sweep, C ABI ff_eloc), not from differentiating this adjoint again
        B = ctx.B
        dev = ctx.stash.y.device
        gz = torch.zeros(B, ctx.model.n_up + ctx.model.n_dn, 2, dtype=torch.float64, device=dev) if gz is None else gz
        gdl = torch.zeros(B, dtype=torch.float64, device=dev) if gdl is None else gdl
        gx, gparams = _backward_through_flow(ctx.cnf, ctx.model, ctx.stash, B, gz, gdl, ctx.needs_input_grad[1])
        if ctx.n_params == 0:
            gparams = []
        return (None, gx, None, *gparams)


class _VHolder(torch.nn.Module):
    """Parameter container with the reference's sub-module names (flow.py:18-37 `V_wrapper`, `F`): the kernels
    read the MLP parameters directly, so these modules only exist to give state_dict() the reference's keys."""

    def __init__(self, v):
        super().__init__()
        self.v = v


class CNF(torch.nn.Module):
    """CNF(v, t_span, nsteps): v is a Backflow; nsteps RK4 steps across t_span
    (the reference's adaptive dopri5 with rtol 1e-6 is replaced by a fixed grid).

    The backflow is registered as `v_wrapper.v` and `f.v` like the reference's CNF (flow.py:28, 37), so
    state_dict() carries the reference's keys (`v_wrapper.v.eta.fc1.weight`, `f.v.eta.fc1.weight`, ...) and its
    checkpoints load with strict=True; `cnf.v` is the same object.  Checkpoints written by round-1 builds of this
    package (`v.eta...` keys) are remapped on load."""

    def __init__(self, v, t_span, nsteps=16):
        super().__init__()
        self.v_wrapper = _VHolder(v)
        self.f = _VHolder(v)
        self.t_span = (float(t_span[0]), float(t_span[1]))
        self.nsteps = int(nsteps)

    @property
    def v(self):
        return self.v_wrapper.v

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        old = prefix + "v."
        for k in [k for k in state_dict if k.startswith(old)]:
            val = state_dict.pop(k)
            for new in ("v_wrapper.v.", "f.v."):
                state_dict.setdefault(prefix + new + k[len(old):], val)
        return super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    def _model(self, n, n_up=None):
        return self.v._model(n, self.t_span, self.nsteps, n_up=n_up)

    def generate(self, z, nframes=None):               # flow.py:42-50
        z = z.detach().contiguous()
        B, n, _ = z.shape
        if nframes is not None:
            # trajectory at torch.linspace(*t_span, nframes) (flow.py:46-49): one fixed-grid sweep per segment,
            # ceil(nsteps / (nframes - 1)) RK4 steps each; equals generate(z) at the last frame when
            # nframes - 1 divides nsteps (same step size, same arithmetic)
            nframes = int(nframes)
            if nframes < 2:
                raise ValueError("CNF.generate: nframes must be at least 2")
            seg = -(-self.nsteps // (nframes - 1))
            t0, t1 = self.t_span
            ts = [t0 + (t1 - t0) * k / (nframes - 1) for k in range(nframes)]
            ts[-1] = t1
            frames = torch.empty(nframes, *z.shape, dtype=z.dtype, device=z.device)
            frames[0] = z
            for k in range(1, nframes):
                m = self.v._model(n, (ts[k - 1], ts[k]), seg)
                L.check(L.lib().ff_cnf_generate(C.byref(m), L.ptr(frames[k - 1]), B, 0, L.ptr(frames[k]), L.stream()))
            return frames
        m = self._model(n)
        x = torch.empty_like(z)
        L.check(L.lib().ff_cnf_generate(C.byref(m), L.ptr(z), B, 0, L.ptr(x), L.stream()))
        return x

    def check_reversibility(self, basedist, batch, orbitals_up=None, orbitals_down=()):   # flow.py:58-71
        """z -> x with log p carried along, then x -> z: prints and returns the two maximal deviations.
        basedist.sample / log_prob take the orbitals when given (FreeFermion), else the reference's signature."""
        print("---- CNF REVERSIBILITY CHECK ----")
        if orbitals_up is not None:
            z = basedist.sample(orbitals_up, orbitals_down, (batch,))
            logp0 = lambda y: basedist.log_prob(orbitals_up, orbitals_down, y)
        else:
            z = basedist.sample((batch,))
            logp0 = basedist.log_prob
        x = self.generate(z)
        # forward-time integral of -div v: the same sweep on the reversed interval
        fwd = CNF(self.v, (self.t_span[1], self.t_span[0]), nsteps=self.nsteps)
        x2, dl_fwd = fwd.delta_logp(z)
        logp = logp0(z) + dl_fwd
        z_reverse, delta_logp = self.delta_logp(x)
        logp_reverse = logp0(z_reverse) - delta_logp
        dz, dlp = (z_reverse - z).abs().max(), (logp_reverse - logp).abs().max()
        print("MaxAbs of z_reverse - z:", dz)
        print("MaxAbs of logp_inverse - logp:", dlp)
        return float(dz), float(dlp), float((x2 - x).abs().max())

    def calibrate_nsteps(self, x, rtol=1e-6, atol=1e-8, min_nsteps=2, max_nsteps=512):
        """Tolerance-driven choice of the fixed grid (the reference integrates with torchdiffeq's adaptive dopri5 at
        rtol 1e-6 / atol 1e-8, nnModule.py:161-162; here every walker takes the same `nsteps` RK4 steps).  Step doubling
        on the sample `x` ([batch, n, 2]): the smallest power-of-two multiple of `min_nsteps` whose (z, delta_logp) agree
        with the result of twice as many steps in torchdiffeq's mixed norm, rms(err / (atol + rtol max(|a|, |b|))) <= 1,
        scaled by 16 / 15 (Richardson: the coarser result carries 16 / 15 of the difference at fourth order).
        Sets and returns self.nsteps; raises if max_nsteps does not reach the tolerance."""
        def norm(a, b):
            tol = atol + rtol * torch.maximum(a.abs(), b.abs())
            return float(((a - b) / tol).pow(2).mean().sqrt()) * 16.0 / 15.0
        keep, ns = self.nsteps, int(min_nsteps)
        try:
            self.nsteps = ns
            z0, d0 = self.delta_logp(x)
            while 2 * ns <= max_nsteps:
                self.nsteps = 2 * ns
                z1, d1 = self.delta_logp(x)
                if max(norm(z0, z1), norm(d0, d1)) <= 1.0:
                    keep = ns
                    return ns
                ns, z0, d0 = 2 * ns, z1, d1
            raise RuntimeError("CNF.calibrate_nsteps: %d RK4 steps do not reach rtol %g / atol %g" % (max_nsteps, rtol, atol))
        finally:
            self.nsteps = keep

    def delta_logp(self, x, params_require_grad=False):  # flow.py:52-56
        params = _flat_params(self.v) if params_require_grad else []
        return _DeltaLogp.apply(self, x, params_require_grad, *params)

    def backflow_potential(self):                        # flow.py:86-92
        return self.v.eta, self.v.mu
