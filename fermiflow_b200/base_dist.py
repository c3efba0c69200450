"""Base distribution |Psi_0|^2 of non-interacting fermions -- mirror of reference
src/base_dist.py (FreeFermion): log_prob via the Slater kernels, sample via the
one-thread-per-walker Metropolis kernel (C ABI ff_free_fermion_logp / ff_metropolis)."""
import torch

from . import _lib as L
from .orbitals import orbital_indices
from .slater import _SlaterGradient, walker_states_from_collection


class BaseDist(object):
    pass


class _FreeFermionLogp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, orb, walker_state, n_up, n_dn):
        shape = x.shape
        xf = x.detach().reshape(-1, n_up + n_dn, 2).contiguous()
        B = xf.shape[0]
        out = torch.empty(B, dtype=xf.dtype, device=xf.device)
        grad = torch.empty_like(xf) if ctx.needs_input_grad[0] else None
        L.check(L.lib().ff_free_fermion_logp(L.ptr(xf), B, n_up, n_dn, L.ptr(orb, torch.int32),
                                             L.ptr(walker_state, torch.int32) if walker_state is not None else None,
                                             L.ptr(out), L.ptr(grad), L.stream()))
        if grad is not None:
            ctx.save_for_backward(x, grad.reshape(shape))
            ctx.meta = (orb, walker_state, n_up, n_dn)
        return out.reshape(shape[:-2])

    @staticmethod
    def backward(ctx, g):
        # differentiable again (Hessian-vector product kernel), like the reference's log_prob (base_dist.py:48-56
        # through slater.py:40-60)
        x, dlog = ctx.saved_tensors
        d = _SlaterGradient.apply(x, dlog, *ctx.meta, 2.0)
        return g[..., None, None] * d, None, None, None, None


def _state_table(states, device):
    rows = [orbital_indices(tuple(up) + tuple(dn), device) for up, dn in states]
    return torch.stack(rows).contiguous()


def _rank_world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


class FreeFermion(BaseDist):
    """Random numbers: the Metropolis kernel draws from Philox4x32-10 with key = seed of the call and counter =
    (GLOBAL walker index, step, particle).  The seed defaults to torch.initial_seed() at the first sample() (so
    torch.manual_seed() controls the chains, as it does for the reference's torch.randn sampler) and advances
    with every call; manual_seed(s) fixes it explicitly.  Under torch.distributed rank r samples the global walkers
    [r B, (r + 1) B): ranks never share a stream, whatever the seed, and a W-rank run of B walkers each draws
    the same chains as one rank with W B walkers.  (torch seeds its default generator from the system entropy:
    without torch.manual_seed or manual_seed the chains differ from process to process, like the reference's.)"""

    def __init__(self, device=torch.device("cuda")):
        self.device = torch.device(device)
        self.seed = None
        self._calls = 0

    def manual_seed(self, seed):
        self.seed, self._calls = int(seed), 0

    # ---- single state ---------------------------------------------------------------
    def log_prob(self, orbitals_up, orbitals_down, x):             # base_dist.py:48-56
        orb = orbital_indices(tuple(orbitals_up) + tuple(orbitals_down), x.device)
        return _FreeFermionLogp.apply(x, orb, None, len(orbitals_up), len(orbitals_down))

    def log_prob_grad_laplacian(self, orbitals_up, orbitals_down, x):
        """log p0(x), grad and Laplacian in one fused kernel (C ABI ff_free_fermion_logp_lap):
        what utils.py:44-65 y_grad_laplacian(self.log_prob, x) returns through 1 + 2N autograd
        passes in the reference."""
        n_up, n_dn = len(orbitals_up), len(orbitals_down)
        orb = orbital_indices(tuple(orbitals_up) + tuple(orbitals_down), x.device)
        xf = x.detach().reshape(-1, n_up + n_dn, 2).contiguous()
        B = xf.shape[0]
        logp = torch.empty(B, dtype=xf.dtype, device=xf.device)
        grad, lap = torch.empty_like(xf), torch.empty_like(logp)
        L.check(L.lib().ff_free_fermion_logp_lap(L.ptr(xf), B, n_up, n_dn, L.ptr(orb, torch.int32), None,
                                                 L.ptr(logp), L.ptr(grad), L.ptr(lap), L.stream()))
        bs = x.shape[:-2]
        return logp.reshape(bs), grad.reshape(x.shape), lap.reshape(bs)

    def _metropolis(self, B, n_up, n_dn, orb, walker_state, steps, tau, noise=None):
        x = torch.empty(B, n_up + n_dn, 2, dtype=torch.float64, device=self.device)
        x0 = nrm = uni = None
        if noise is not None:
            x0, nrm, uni = (t.contiguous() for t in noise)
        if self.seed is None:
            self.seed = torch.initial_seed() & 0xFFFFFFFFFFFFFFFF
        seed = (self.seed + 0x9E3779B97F4A7C15 * self._calls) & 0xFFFFFFFFFFFFFFFF
        self._calls += 1
        rank, _ = _rank_world()
        L.check(L.lib().ff_metropolis(B, n_up, n_dn, L.ptr(orb, torch.int32),
                                      L.ptr(walker_state, torch.int32) if walker_state is not None else None,
                                      int(steps), float(tau), seed, rank * B, L.ptr(x0), L.ptr(nrm), L.ptr(uni),
                                      L.ptr(x), None, L.stream()))
        return x

    def sample(self, orbitals_up, orbitals_down, sample_shape, equilibrim_steps=100, tau=0.1, noise=None):
        """base_dist.py:58-70.  noise=(x0, normals, uniforms) replays given random numbers."""
        B = 1
        for s in sample_shape:
            B *= int(s)
        orb = orbital_indices(tuple(orbitals_up) + tuple(orbitals_down), self.device)
        x = self._metropolis(B, len(orbitals_up), len(orbitals_down), orb, None, equilibrim_steps, tau, noise)
        return x.reshape(*sample_shape, len(orbitals_up) + len(orbitals_down), 2)

    # ---- several states (finite temperature) ----------------------------------------
    def log_prob_multstates(self, states, state_indices_collection, x, method=2):   # base_dist.py:72-101
        if len(x.shape[:-2]) != 1:
            raise ValueError("FreeFermion.log_prob_multstates: x is required to have "
                             "only one batch dimension.")
        table = states if isinstance(states, torch.Tensor) else _state_table(states, x.device)
        ws = state_indices_collection if isinstance(state_indices_collection, torch.Tensor) \
            else walker_states_from_collection(state_indices_collection, x.device)
        n_up = len(states[0][0]) if not isinstance(states, torch.Tensor) else x.shape[-2]
        return _FreeFermionLogp.apply(x, table, ws, n_up, x.shape[-2] - n_up)

    def sample_multstates(self, states, state_indices_collection, sample_shape,
                          equilibrim_steps=100, tau=0.1, cpu=False, method=2, noise=None):  # base_dist.py:103-134
        if len(sample_shape) != 1:
            raise ValueError("FreeFermion.sample_multstates: sample_shape is "
                             "required to have only one batch dimension.")
        if cpu:
            raise RuntimeError("fermiflow_b200 has no CPU path (cpu=True)")
        table = states if isinstance(states, torch.Tensor) else _state_table(states, self.device)
        ws = state_indices_collection if isinstance(state_indices_collection, torch.Tensor) \
            else walker_states_from_collection(state_indices_collection, self.device)
        n_up = len(states[0][0]) if not isinstance(states, torch.Tensor) else table.shape[1]
        n_dn = table.shape[1] - n_up
        return self._metropolis(int(sample_shape[0]), n_up, n_dn, table, ws, equilibrim_steps, tau, noise)
