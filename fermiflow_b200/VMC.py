"""Variational Monte Carlo drivers -- mirror of reference src/VMC.py (GSVMC, BetaVMC).

forward(batch) does what the reference does per iteration -- sample walkers, local energy,
and a scalar whose backward() leaves the energy gradient in the parameters' .grad -- but
as three fused CUDA stages:
  1. Metropolis sampling of the base state + flow z -> x     (ff_metropolis, ff_cnf_generate)
  2. one forward-mode sweep x -> z giving log p, its gradient and Laplacian, E_loc, and
     the stash of stage inputs                                (ff_eloc)
  3. on backward(): the exact reverse sweep + parameter-gradient reduction
                                                              (ff_logp_backward)
Walkers are independent Markov chains: under torch.distributed each rank works on its own
`batch` walkers and only the energy moments and the few-KB parameter gradient are
all-reduced (allreduce_gradients()).
"""
from collections import Counter

import torch
import torch.distributed as dist

from . import _lib as L
from .base_dist import _FreeFermionLogp, _state_table
from .flow import _backward_through_flow, _flat_params
from .orbitals import orbital_indices
from .potentials import HO, CoulombPairPotential
from .utils import eloc_sweep


def _potential_flags(pair_potential, sp_potential):
    if not isinstance(pair_potential, CoulombPairPotential):
        raise NotImplementedError("the CUDA path implements the reference's CoulombPairPotential only")
    if sp_potential is not None and not isinstance(sp_potential, HO):
        raise NotImplementedError("the CUDA path implements the reference's HO single-particle potential only")
    return float(pair_potential.Z), sp_potential is not None


class _LogpFromSweep(torch.autograd.Function):
    """log p(x; theta) of an E_loc sweep as a differentiable function of the flow
    parameters: backward = ff_logp_backward seeded with d log p0/dz and -1."""

    @staticmethod
    def forward(ctx, owner, res, orb, walker_state, n_up, n_dn, *params):
        ctx.owner, ctx.res, ctx.orb, ctx.ws, ctx.nud = owner, res, orb, walker_state, (n_up, n_dn)
        return res.logp.clone()

    @staticmethod
    def backward(ctx, g):
        res = ctx.res
        B = g.shape[0]
        z = res.z.detach().requires_grad_(True)
        with torch.enable_grad():
            lp0 = _FreeFermionLogp.apply(z, ctx.orb, ctx.ws, *ctx.nud)
        g0, = torch.autograd.grad(lp0, z, grad_outputs=g.contiguous())
        _, gparams = _backward_through_flow(ctx.owner.cnf, res.model, res.stash, B, g0, -g, False)
        return (None, None, None, None, None, None, *gparams)


def _global_mean_std(v):
    """mean and unbiased std of v over all ranks (3-number all-reduce)."""
    s = torch.stack([v.sum(), (v * v).sum(), torch.tensor(float(v.numel()), device=v.device, dtype=v.dtype)])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(s)
    tot, tot2, cnt = s.tolist()
    mean = tot / cnt
    var = max(tot2 - cnt * mean * mean, 0.0) / max(cnt - 1.0, 1.0)
    return mean, var ** 0.5, cnt


class _VMCBase(torch.nn.Module):
    def allreduce_gradients(self):
        """Sum the parameter gradients over ranks (call between backward() and step())."""
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return
        grads = [p.grad for p in self.parameters() if p.grad is not None]
        if not grads:
            return
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat)
        o = 0
        for g in grads:
            g.copy_(flat[o:o + g.numel()].view_as(g))
            o += g.numel()


class GSVMC(_VMCBase):
    """Ground-state VMC (VMC.py:4-61)."""

    def __init__(self, nup, ndown, orbitals, basedist, cnf, pair_potential, sp_potential=None):
        super().__init__()
        self.orbitals_up, self.orbitals_down = orbitals.orbitals[:nup], orbitals.orbitals[:ndown]
        self.nup, self.ndown = nup, ndown
        self.basedist = basedist
        self.cnf = cnf
        self.pair_potential = pair_potential
        self.sp_potential = sp_potential
        self._Z, self._harmonic = _potential_flags(pair_potential, sp_potential)

    def _orb(self, device):
        return orbital_indices(tuple(self.orbitals_up) + tuple(self.orbitals_down), device)

    def sample(self, sample_shape):                                   # VMC.py:31-34
        z = self.basedist.sample(self.orbitals_up, self.orbitals_down, sample_shape)
        x = self.cnf.generate(z)
        return z, x

    def logp(self, x, params_require_grad=False):                     # VMC.py:36-39
        z, delta_logp = self.cnf.delta_logp(x, params_require_grad=params_require_grad)
        return self.basedist.log_prob(self.orbitals_up, self.orbitals_down, z) - delta_logp

    def local_energy(self, x, stash=False):
        """log p, grad, laplacian, kinetic, potential, E_loc at x (VMC.py:44-55)."""
        return eloc_sweep(self.cnf, x, self._orb(x.device), None, self.nup, self._Z, self._harmonic, stash=stash)

    def forward(self, batch):                                         # VMC.py:41-61
        _, x = self.sample((batch,))
        res = self.local_energy(x, stash=True)
        Eloc = res.eloc
        self.E, self.E_std, nglobal = _global_mean_std(Eloc)
        self.last = res
        logp_full = _LogpFromSweep.apply(self, res, self._orb(x.device), None, self.nup, self.ndown,
                                         *_flat_params(self.cnf.v))
        # sum / global count == mean over all ranks once the gradients are all-reduced
        gradE = (logp_full * (Eloc - self.E)).sum() / nglobal
        return gradE


class BetaVMC(_VMCBase):
    """Finite-temperature VMC (VMC.py:63-171)."""

    def __init__(self, beta, nup, ndown, deltaE, boltzmann, orbitals, basedist, cnf,
                 pair_potential, sp_potential=None):
        super().__init__()
        self.beta = beta
        self.nup, self.ndown = nup, ndown
        self.states, self.Es_original = orbitals.fermion_states(nup, ndown, deltaE)
        self.Es_original = torch.tensor(self.Es_original, dtype=torch.float64)
        self.Nstates = len(self.states)
        self.log_state_weights = torch.nn.Parameter(
            -self.beta * (self.Es_original - self.Es_original[0])
            if boltzmann else torch.randn(self.Nstates, dtype=torch.float64))
        self.basedist = basedist
        self.cnf = cnf
        self.pair_potential = pair_potential
        self.sp_potential = sp_potential
        self._Z, self._harmonic = _potential_flags(pair_potential, sp_potential)
        self._table = None

    def _state_table(self, device):
        if self._table is None or self._table.device != device:
            self._table = _state_table(self.states, device)
        return self._table

    @property
    def state_indices_collection(self):
        """Counter {state index: multiplicity} of the last sample (VMC.py:97)."""
        return Counter({int(s): int(c) for s, c in enumerate(self.state_counts.tolist()) if c})

    def sample_states(self, batch, uniforms=None):
        """Occupation sampling (VMC.py:94-97): sorted state index per walker."""
        dev = self.log_state_weights.device
        u = torch.rand(batch, dtype=torch.float64, device=dev) if uniforms is None else uniforms.contiguous()
        state = torch.empty(batch, dtype=torch.int32, device=dev)
        counts = torch.empty(self.Nstates, dtype=torch.int32, device=dev)
        work = torch.empty(self.Nstates, dtype=torch.float64, device=dev)
        L.check(L.lib().ff_occupation_sample(L.ptr(self.log_state_weights.detach().contiguous()), self.Nstates,
                                             L.ptr(u), batch, L.ptr(state, torch.int32),
                                             L.ptr(counts, torch.int32), L.ptr(work), L.stream()))
        self.state_indices, self.state_counts = state, counts
        return state

    def sample(self, sample_shape, nframes=None):                      # VMC.py:89-108
        batch = int(sample_shape[0])
        state = self.sample_states(batch)
        z = self.basedist.sample_multstates(self._state_table(state.device), state, sample_shape)
        x = self.cnf.generate(z, nframes=nframes)
        return z, x

    def logp(self, x, params_require_grad=False):                      # VMC.py:110-118
        z, delta_logp = self.cnf.delta_logp(x, params_require_grad=params_require_grad)
        log_prob_z = _FreeFermionLogp.apply(z, self._state_table(x.device), self.state_indices, self.nup, self.ndown)
        return log_prob_z - delta_logp

    def forward(self, batch):                                          # VMC.py:120-171
        _, x = self.sample((batch,))
        table, state = self._state_table(x.device), self.state_indices
        res = eloc_sweep(self.cnf, x, table, state, self.nup, self._Z, self._harmonic, stash=True)
        self.last = res
        Eloc = res.eloc
        self.E, self.E_std, nglobal = _global_mean_std(Eloc)

        logp_all = torch.log_softmax(self.log_state_weights, dim=0)
        logp_states = logp_all[state.long()]
        Floc = Eloc + logp_states.detach() / self.beta
        self.F, self.F_std, _ = _global_mean_std(Floc)
        self.S = -_global_mean_std(logp_states.detach())[0]
        self.logp_states_all = logp_all.detach()
        self.S_analytical = -(self.logp_states_all * self.logp_states_all.exp()).sum().item()

        # sum_w logp_all[state_w] (Floc_w - F) = sum_s logp_all[s] * (sum of the weights of the walkers in s): the
        # per-state sums are one index_add_, and autograd never sees an 8000-fold duplicated gather (VMC.py:159)
        wstate = torch.zeros(self.Nstates, dtype=Eloc.dtype, device=Eloc.device).index_add_(0, state.long(), Floc - self.F)
        gradF_phi = (logp_all * wstate).sum() / nglobal

        # E_loc minus its mean over the walkers that share a state (VMC.py:163-168)
        sl = state.long()
        ssum = torch.zeros(self.Nstates, dtype=Eloc.dtype, device=Eloc.device).index_add_(0, sl, Eloc)
        cnt = self.state_counts.to(Eloc.dtype)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            both = torch.stack([ssum, cnt])
            dist.all_reduce(both)
            ssum, cnt = both[0], both[1]
        Eloc_x_mean = (ssum / cnt.clamp_min(1.0))[sl]
        logp_full = _LogpFromSweep.apply(self, res, table, state, self.nup, self.ndown, *_flat_params(self.cnf.v))
        gradF_theta = (logp_full * (Eloc - Eloc_x_mean)).sum() / nglobal
        return gradF_phi, gradF_theta
