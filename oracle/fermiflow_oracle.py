"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the FermiFlow per-walker VMC hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this file.  The product (fermiflow_b200/) never does.

It restates, in torch-CPU float64, what the reference computes on this path, with the
ODE solved by the fixed-step 3/8-rule RK4 that `torchdiffeq.odeint(method="rk4")`
implements (see oracle/torchdiffeq_shim) instead of the adaptive default, and with the
gradient / Laplacian of log p obtained by plain nested autograd through the unrolled
steps (what /root/reference/src/utils.py:44 y_grad_laplacian does through the
reference's adjoint).  Every function cites the reference lines it follows.

Pinned by tests/test_oracle_*.py against
  * outputs of the real reference run in the build container (tests/golden/*.npz,
    made by oracle/gen_golden.py), and
  * the reference's own known-answer tests (HO eigen-energies, antisymmetry,
    equivariance, divergence-vs-autograd).
"""
import math

import numpy as np
import torch

torch.set_default_dtype(torch.float64)

# --------------------------------------------------------------------------------------
# MLP.py:4-45 -- single hidden layer, sigmoid, scalar output, D_in = 1 on this path
# --------------------------------------------------------------------------------------

def mlp_value(p, d):
    """MLP.py:30-32.  p = (w1[H], b1[H], w2[H]); d any shape; returns same shape."""
    w1, b1, w2 = p
    return (torch.sigmoid(d[..., None] * w1 + b1) * w2).sum(-1)


def mlp_grad(p, d):
    """MLP.py:37-45 hand-coded d(MLP)/d(input)."""
    w1, b1, w2 = p
    s = torch.sigmoid(d[..., None] * w1 + b1)
    return (w2 * s * (1.0 - s) * w1).sum(-1)


# --------------------------------------------------------------------------------------
# equivariant_funs.py:4-102 -- backflow velocity field and its divergence
# --------------------------------------------------------------------------------------

def _pairs(n):
    iu = torch.triu_indices(n, n, offset=1)
    return iu[0], iu[1]


def backflow_v(x, eta, mu=None):
    """equivariant_funs.py:17-31 (_e_e), 49-62 (_e_n), 80-89 (forward).

    x: (batch, n, 2).  The reference adds eye(n) on the diagonal and subtracts
    eta(|ones|) afterwards, which cancels exactly; here the i != j terms are summed
    directly.
    """
    b, n, dim = x.shape
    i, j = _pairs(n)
    rij = x[:, i] - x[:, j]                       # (b, P, dim)
    dij = rij.norm(dim=-1)
    g = mlp_value(eta, dij)[..., None] * rij       # eta(d) * r_ij
    v = torch.zeros_like(x)
    v = v.index_add(1, i, g).index_add(1, j, -g)
    if mu is not None:
        v = v + mlp_value(mu, x.norm(dim=-1))[..., None] * x
    return v


def backflow_div(x, eta, mu=None):
    """equivariant_funs.py:33-47 (_e_e_divergence), 64-78 (_e_n_divergence), 91-102."""
    b, n, dim = x.shape
    i, j = _pairs(n)
    dij = (x[:, i] - x[:, j]).norm(dim=-1)
    div = 2.0 * (mlp_grad(eta, dij) * dij + dim * mlp_value(eta, dij)).sum(-1)
    if mu is not None:
        di = x.norm(dim=-1)
        div = div + (mlp_grad(mu, di) * di + dim * mlp_value(mu, di)).sum(-1)
    return div


# --------------------------------------------------------------------------------------
# torchdiffeq fixed-grid rk4 (3/8 rule); flow.py:42-56 generate / delta_logp
# --------------------------------------------------------------------------------------

def rk4_38(f, y, t0, t1, nsteps):
    """y is a tuple of tensors; f(y) -> tuple.  nsteps equal steps from t0 to t1."""
    h = (t1 - t0) / nsteps
    for _ in range(nsteps):
        k1 = f(y)
        k2 = f(tuple(a + h * b / 3 for a, b in zip(y, k1)))
        k3 = f(tuple(a + h * (c - b / 3) for a, b, c in zip(y, k1, k2)))
        k4 = f(tuple(a + h * (b - c + d) for a, b, c, d in zip(y, k1, k2, k3)))
        y = tuple(a + (b + 3 * (c + d) + e) * h * 0.125
                  for a, b, c, d, e in zip(y, k1, k2, k3, k4))
    return y


def cnf_generate(z, eta, mu, t_span, nsteps):
    """flow.py:42-50: x = flow of dx/dt = v(x) from t_span[0] to t_span[1]."""
    return rk4_38(lambda y: (backflow_v(y[0], eta, mu),), (z,), t_span[0], t_span[1], nsteps)[0]


def cnf_delta_logp(x, eta, mu, t_span, nsteps):
    """flow.py:52-56: integrate (v, -div v) from t_span[1] back to t_span[0]."""
    f = lambda y: (backflow_v(y[0], eta, mu), -backflow_div(y[0], eta, mu))  # noqa: E731
    z, dl = rk4_38(f, (x, torch.zeros(x.shape[0])), t_span[1], t_span[0], nsteps)
    return z, dl


# --------------------------------------------------------------------------------------
# orbitals.py:66-90 HO2D orbitals; slater.py:4-68 log|det|; base_dist.py:48-56 log_prob
# --------------------------------------------------------------------------------------

HO2D_QUANTA = [(nx, n - nx) for n in range(8) for nx in range(n + 1)]   # orbitals.py:89
HO2D_ENERGIES = [n + 1 for n in range(8) for nx in range(n + 1)]        # orbitals.py:90


def hermite_fn(n, x):
    """Normalised Hermite polynomial h_n(x) (orbitals.py:74-83 lists n = 0..7
    explicitly); three-term recursion h_{k+1} = sqrt(2/(k+1)) x h_k - sqrt(k/(k+1)) h_{k-1}."""
    h_prev, h = torch.zeros_like(x), torch.ones_like(x)
    for k in range(n):
        h_prev, h = h, math.sqrt(2.0 / (k + 1)) * x * h - math.sqrt(k / (k + 1.0)) * h_prev
    return h


def ho2d_orbital(k, x):
    """orbitals.py:84-87: pi^-1/2 exp(-r^2/2) h_nx(x) h_ny(y), k indexes HO2D_QUANTA."""
    nx, ny = HO2D_QUANTA[k]
    return (1.0 / math.sqrt(math.pi)) * torch.exp(-0.5 * (x ** 2).sum(-1)) \
        * hermite_fn(nx, x[..., 0]) * hermite_fn(ny, x[..., 1])


def logabs_slater(orb_idx, x):
    """slater.py:12-38 forward: log|det phi_j(r_i)|.  orb_idx: list of orbital indices
    (same for every walker) or LongTensor (batch, n) (per-walker occupation, the
    LogAbsSlaterDetMultStates case slater.py:70-118)."""
    n = x.shape[-2]
    if n == 0:
        return torch.zeros(x.shape[:-2])
    if isinstance(orb_idx, torch.Tensor) and orb_idx.dim() == 2:
        cols = []
        allv = torch.stack([ho2d_orbital(k, x) for k in range(len(HO2D_QUANTA))], -1)  # (b,n,36)
        for jcol in range(n):
            cols.append(torch.gather(allv, -1, orb_idx[:, None, jcol, None].expand(-1, n, 1))[..., 0])
        D = torch.stack(cols, -1)
    else:
        D = torch.stack([ho2d_orbital(int(k), x) for k in orb_idx], -1)
    return torch.linalg.slogdet(D)[1]


def free_fermion_logp(orb_up, orb_dn, x):
    """base_dist.py:48-56: 2 (log|det_up| + log|det_down|)."""
    nup = orb_up.shape[-1] if isinstance(orb_up, torch.Tensor) else len(orb_up)
    return 2.0 * (logabs_slater(orb_up, x[..., :nup, :]) + logabs_slater(orb_dn, x[..., nup:, :]))


def metropolis_sample(orb_up, orb_dn, x0, normals, uniforms, tau=0.1):
    """base_dist.py:58-70 with the random numbers supplied: x0 (b,n,2) initial normals,
    normals (steps,b,n,2), uniforms (steps,b)."""
    x = x0.clone()
    logp = free_fermion_logp(orb_up, orb_dn, x)
    for eps, u in zip(normals, uniforms):
        new_x = x + tau * eps
        new_logp = free_fermion_logp(orb_up, orb_dn, new_x)
        accept = u < torch.exp(new_logp - logp)
        x[accept] = new_x[accept]
        logp[accept] = new_logp[accept]
    return x


# --------------------------------------------------------------------------------------
# potentials.py
# --------------------------------------------------------------------------------------

def potential_ho(x):
    """potentials.py:13-14."""
    return 0.5 * (x ** 2).sum(dim=(-2, -1))


def potential_coulomb(x, Z):
    """potentials.py:23-46: Z sum_{i<j} 1/|r_i - r_j|."""
    i, j = _pairs(x.shape[-2])
    return (Z / (x[:, i] - x[:, j]).norm(dim=-1)).sum(-1)


# --------------------------------------------------------------------------------------
# VMC.py:36-61 -- log p, local energy, energy gradient
# --------------------------------------------------------------------------------------

def logp(x, orb_up, orb_dn, eta, mu, t_span, nsteps):
    """VMC.py:36-39."""
    z, dl = cnf_delta_logp(x, eta, mu, t_span, nsteps)
    return free_fermion_logp(orb_up, orb_dn, z) - dl


def logp_grad_laplacian(x, *model):
    """utils.py:44-65 y_grad_laplacian applied to VMC.py:36 logp."""
    x = x.detach().clone().requires_grad_(True)
    xf = x.flatten(1)
    y = logp(xf.view_as(x), *model)
    g, = torch.autograd.grad(y.sum(), xf, create_graph=True)
    lap = torch.zeros(x.shape[0])
    for c in range(xf.shape[1]):
        lap = lap + torch.autograd.grad(g[:, c].sum(), xf, retain_graph=True)[0][:, c]
    return y.detach(), g.detach().view_as(x), lap


def free_fermion_grad_laplacian(orb_up, orb_dn, x):
    """utils.py:44-65 y_grad_laplacian applied to base_dist.py:48-56 FreeFermion.log_prob:
    log p0, its gradient and Laplacian through 1 + 2N autograd passes (the reference's path for
    the BASELINE.json "batched log|det| + exact Laplacian" microbenchmark)."""
    x = x.detach().clone().requires_grad_(True)
    xf = x.flatten(1)
    y = free_fermion_logp(orb_up, orb_dn, xf.view_as(x))
    g, = torch.autograd.grad(y.sum(), xf, create_graph=True)
    lap = torch.zeros(x.shape[0])
    for c in range(xf.shape[1]):
        lap = lap + torch.autograd.grad(g[:, c].sum(), xf, retain_graph=True)[0][:, c]
    return y.detach(), g.detach().view_as(x), lap


def local_energy(x, orb_up, orb_dn, eta, mu, t_span, nsteps, Z, harmonic=True):
    """VMC.py:46-55: E_loc = -1/4 lap - 1/8 |grad|^2 + V."""
    lp, g, lap = logp_grad_laplacian(x, orb_up, orb_dn, eta, mu, t_span, nsteps)
    kin = -0.25 * lap - 0.125 * (g ** 2).sum(dim=(-2, -1))
    pot = potential_coulomb(x, Z)
    if harmonic:
        pot = pot + potential_ho(x)
    return dict(logp=lp, grad=g, lap=lap, kinetic=kin, potential=pot, eloc=kin + pot)


def weighted_logp_param_grad(x, weights, orb_up, orb_dn, eta, mu, t_span, nsteps):
    """d/dtheta sum_b weights_b log p(x_b; theta): VMC.py:57-59 with weights =
    (E_loc - E)/batch.  Returns the gradients in the order (eta..., mu...)."""
    leaves = [p.detach().clone().requires_grad_(True) for p in eta]
    eta_l = tuple(leaves[:3])
    mu_l = None
    if mu is not None:
        ml = [p.detach().clone().requires_grad_(True) for p in mu]
        leaves += ml
        mu_l = tuple(ml)
    loss = (logp(x.detach(), orb_up, orb_dn, eta_l, mu_l, t_span, nsteps) * weights).sum()
    return [g.detach() for g in torch.autograd.grad(loss, leaves)]


# --------------------------------------------------------------------------------------
# Finite temperature: orbitals.py:16-64 state enumeration, VMC.py:94-101 occupations
# --------------------------------------------------------------------------------------

def fermion_states(nup, ndown, deltaE):
    """orbitals.py:34-64 (polarised case): every nup-subset of the 36 HO2D orbitals whose
    energy is <= E0 + deltaE, ordered by (energy, lexicographic index tuple)."""
    if ndown != 0:
        raise ValueError("Only the polarized case (i.e., ndown = 0) is allowed "
                         "in the present implementation.")
    Es = HO2D_ENERGIES
    emax = sum(Es[:nup]) + deltaE
    found = _subsets_pruned(nup, emax, Es)
    found.sort(key=lambda ec: ec[0])           # stable: lexicographic order kept inside a level
    return [c for _, c in found], [e for e, _ in found]


def _subsets_pruned(k, pmax, prices):
    """orbitals.py:16-32 restated as a depth-first search with the same pruning bound
    (prices are sorted ascending, so the cheapest completion is the next k-i items)."""
    out, n = [], len(prices)

    def rec(start, chosen, total):
        need = k - len(chosen)
        if need == 0:
            out.append((total, tuple(chosen)))
            return
        for nxt in range(start, n - need + 1):
            if sum(prices[nxt:nxt + need]) > pmax - total:
                continue
            rec(nxt + 1, chosen + [nxt], total + prices[nxt])
    rec(0, [], 0)
    return out


def boltzmann_logits(beta, Es):
    """VMC.py:78-80 with boltzmann=True."""
    Es = torch.as_tensor(Es, dtype=torch.float64)
    return -beta * (Es - Es[0])


def categorical_from_uniforms(logits, u):
    """VMC.py:94-97: Categorical(logits).sample, restated as inverse-CDF on supplied
    uniforms (what torch.multinomial does on CPU: first index whose normalised
    cumulative probability is >= u), followed by the sort of VMC.py:97."""
    p = torch.softmax(torch.as_tensor(logits, dtype=torch.float64), -1)
    cdf = torch.cumsum(p, -1)
    cdf = cdf / cdf[-1]
    idx = torch.searchsorted(cdf, torch.as_tensor(u, dtype=torch.float64), right=False)
    idx = idx.clamp_(max=len(p) - 1)
    return torch.sort(idx)[0]


def beta_vmc_estimators(Eloc, state_idx, log_state_weights, beta):
    """VMC.py:139-171 given the local energies: state_idx (batch,) sorted state index per walker
    (VMC.py:97, 144-145), log_state_weights the (Nstates,) logits (a leaf tensor if its gradient is wanted).
    Returns E, E_std, F, F_std, S, S_analytical, gradF_phi (scalar with graph, VMC.py:162) and theta_weights,
    the per-walker factors of logp_full in gradF_theta (VMC.py:164-169: (Eloc - mean of Eloc over the walkers
    of the same state) / batch)."""
    Eloc = Eloc.detach()
    logp_all = torch.log_softmax(log_state_weights, -1)          # Categorical(logits).log_prob
    logp_states = logp_all[state_idx]
    Floc = Eloc + logp_states.detach() / beta
    F = Floc.mean().item()
    out = dict(E=Eloc.mean().item(), E_std=Eloc.std().item(), F=F, F_std=Floc.std().item(),
               S=-logp_states.detach().mean().item(),
               S_analytical=-(logp_all.detach() * logp_all.detach().exp()).sum().item())
    out["gradF_phi"] = (logp_states * (Floc - F)).mean()
    x_mean = torch.empty_like(Eloc)
    for s in torch.unique(state_idx):
        sel = state_idx == s
        x_mean[sel] = Eloc[sel].mean()
    out["theta_weights"] = (Eloc - x_mean) / Eloc.shape[0]
    return out


# --------------------------------------------------------------------------------------
# Philox4x32-10 counter RNG (the generator the CUDA sampler uses), in numpy
# --------------------------------------------------------------------------------------

def philox4x32_10(counter, key):
    """counter: (..., 4) uint32, key: (..., 2) uint32 -> (..., 4) uint32.
    Salmon et al., SC'11, the published Philox-4x32 with 10 rounds."""
    c = np.array(counter, dtype=np.uint64, copy=True)
    k0 = np.array(key[..., 0], dtype=np.uint64, copy=True)
    k1 = np.array(key[..., 1], dtype=np.uint64, copy=True)
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    W0, W1 = np.uint64(0x9E3779B9), np.uint64(0xBB67AE85)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = M0 * c[..., 0]
        p1 = M1 * c[..., 2]
        n0 = (p1 >> np.uint64(32)) ^ c[..., 1] ^ k0
        n1 = p1 & mask
        n2 = (p0 >> np.uint64(32)) ^ c[..., 3] ^ k1
        n3 = p0 & mask
        c = np.stack([n0, n1, n2, n3], -1)
        k0 = (k0 + W0) & mask
        k1 = (k1 + W1) & mask
    return c.astype(np.uint32)


def u01_from_bits(hi, lo):
    """53-bit uniform in (0, 1): ((hi << 21 ^ lo >> 11) + 0.5) * 2^-53."""
    x = (np.asarray(hi, np.uint64) << np.uint64(21)) ^ (np.asarray(lo, np.uint64) >> np.uint64(11))
    return (x.astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)
