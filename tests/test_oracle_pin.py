"""Pins the CPU oracle (oracle/fermiflow_oracle.py) to the real reference: every array in
tests/golden/*.npz was produced by /root/reference/src itself (oracle/gen_golden.py)."""
import numpy as np
import torch

from oracle import fermiflow_oracle as O

torch.set_default_dtype(torch.float64)
T = torch.from_numpy


def mlp(g, name):
    return tuple(T(g[name + "_" + k]) for k in ("w1", "b1", "w2"))


def close(a, b, rtol, atol=0.0):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    err = np.abs(a - b).max()
    assert err <= atol + rtol * np.abs(b).max(), (err, np.abs(b).max())


def test_backflow_matches_reference(golden):
    g = golden("backflow")
    x, eta, mu = T(g["x"]), mlp(g, "eta"), mlp(g, "mu")
    close(O.backflow_v(x, eta, mu), g["v"], 1e-13)
    close(O.backflow_div(x, eta, mu), g["div"], 1e-13)
    close(O.backflow_v(x, eta), g["v_nomu"], 1e-13)
    close(O.backflow_div(x, eta), g["div_nomu"], 1e-13)


def test_orbitals_match_reference(golden):
    g = golden("slater")
    x = T(g["x_orb"])
    vals = torch.stack([O.ho2d_orbital(k, x) for k in range(36)], -1)
    close(vals, g["orbitals"], 1e-13)
    assert list(g["Es"]) == O.HO2D_ENERGIES


def _grad_lap(f, x):
    x = x.clone().requires_grad_(True)
    xf = x.flatten(1)
    y = f(xf.view_as(x))
    gr, = torch.autograd.grad(y.sum(), xf, create_graph=True)
    lap = sum(torch.autograd.grad(gr[:, c].sum(), xf, retain_graph=True)[0][:, c] for c in range(xf.shape[1]))
    return y.detach(), gr.detach().view_as(x), lap


def test_logabs_slater_matches_reference(golden):
    g = golden("slater")
    for name in ("gs6", "gs10", "rand7"):
        idx, x = list(g[name + "_idx"]), T(g[name + "_x"])
        y, gr, lap = _grad_lap(lambda t: O.logabs_slater(idx, t), x)
        close(y, g[name + "_logabsdet"], 1e-12)
        close(gr, g[name + "_grad"], 1e-10)
        close(lap, g[name + "_lap"], 1e-9)
    y, gr, lap = _grad_lap(lambda t: O.free_fermion_logp([0, 1, 2], [0, 1], t), T(g["ff_x"]))
    close(y, g["ff_logp"], 1e-12)
    close(gr, g["ff_grad"], 1e-10)
    close(lap, g["ff_lap"], 1e-9)


def test_states_match_reference(golden):
    g = golden("states")
    for nup, dE in ((3, 2), (6, 2), (3, 4), (10, 2), (4, 3)):
        st, Es = O.fermion_states(nup, 0, dE)
        assert np.array_equal(np.array(st), g["states_%d_%d" % (nup, dE)])
        assert np.array_equal(np.array(Es), g["Es_%d_%d" % (nup, dE)])
    st, _ = O.fermion_states(3, 0, 2)
    occ = torch.tensor(st)[T(g["ms_state_idx"])]
    y, gr, lap = _grad_lap(lambda t: O.logabs_slater(occ, t), T(g["ms_x"]))
    close(y, g["ms_logabsdet"], 1e-12)
    close(gr, g["ms_grad"], 1e-10)
    close(lap, g["ms_lap"], 1e-9)


def test_potentials_match_reference(golden):
    g = golden("potentials")
    x = T(g["x"])
    close(O.potential_ho(x), g["ho"], 1e-14)
    close(O.potential_coulomb(x, float(g["Z"])), g["coulomb"], 1e-14)


def _model(g):
    return (list(range(int(g["nup"]))), list(range(int(g["ndown"]))), mlp(g, "eta"), mlp(g, "mu"),
            tuple(g["t_span"]))


def test_pipeline_same_discrete_solver(golden):
    """Reference run with odeint(method='rk4', step 1/16): the forward quantities are the
    same discrete algorithm as the oracle -> rounding-level agreement."""
    g = golden("pipeline")
    up, dn, eta, mu, ts = _model(g)
    x = O.cnf_generate(T(g["z0"]), eta, mu, ts, 16)
    close(x, g["rk4s16_x"], 1e-13)
    z, dl = O.cnf_delta_logp(T(g["rk4s16_x"]), eta, mu, ts, 16)
    close(z, g["rk4s16_zback"], 1e-13)
    close(dl, g["rk4s16_delta_logp"], 1e-12, 1e-15)
    close(O.logp(T(g["rk4s16_x"]), up, dn, eta, mu, ts, 16), g["rk4s16_logp"], 1e-13)


def test_pipeline_converges_to_tight_reference(golden):
    """Reference with its adaptive solver at rtol 1e-11 (adjoint gradients, nested adjoint
    Laplacian) vs the oracle's exact derivatives of the 64-step discrete flow: both
    approximate the same continuous quantities."""
    g = golden("pipeline")
    up, dn, eta, mu, ts = _model(g)
    x = T(g["tight_x"])
    r = O.local_energy(x, up, dn, eta, mu, ts, 64, float(g["Z"]))
    close(r["logp"], g["tight_logp"], 1e-10)
    close(r["grad"], g["tight_grad"], 1e-8)
    close(r["lap"], g["tight_lap"], 1e-7)
    close(r["eloc"], g["tight_eloc"], 1e-7)
    gr = O.weighted_logp_param_grad(x, T(g["weights"]), up, dn, eta, mu, ts, 64)
    for a, k in zip(gr, ("eta_w1", "eta_b1", "eta_w2", "mu_w1", "mu_b1", "mu_w2")):
        close(a, g["tight_g_" + k], 1e-8, 1e-12)


def test_pipeline_default_tolerance_consistent(golden):
    """The reference at its default rtol=1e-6 agrees with the 16-step oracle to that level."""
    g = golden("pipeline")
    up, dn, eta, mu, ts = _model(g)
    r = O.local_energy(T(g["default_x"]), up, dn, eta, mu, ts, 16, float(g["Z"]))
    close(r["logp"], g["default_logp"], 1e-5)
    close(r["eloc"], g["default_eloc"], 1e-4)


def test_reference_port_matches_reference_default(golden):
    """oracle/reference_port.py (the reference's own adaptive + adjoint algorithm, used as the
    CPU baseline of bench.py) against the real reference's default-tolerance run."""
    from oracle import reference_port as R
    g = golden("pipeline")
    up, dn, eta, mu, ts = _model(g)
    params = list(eta) + list(mu)
    x = R.generate(T(g["z0"]), params, True, ts)
    close(x, g["default_x"], 1e-12)
    xr = T(g["default_x"]).clone().requires_grad_(True)
    lp, gr, lap = R.y_grad_laplacian(lambda t: R.logp(t, up, dn, params, True, ts, False), xr)
    close(lp.detach(), g["default_logp"], 1e-11)
    close(gr.detach(), g["default_grad"], 1e-9)
    close(lap.detach(), g["default_lap"], 1e-8)
    leaves = [p.clone().requires_grad_(True) for p in params]
    lpf = R.logp(T(g["default_x"]), up, dn, leaves, True, ts, True)
    grads = torch.autograd.grad((lpf * T(g["weights"])).sum(), leaves)
    for a, k in zip(grads, ("eta_w1", "eta_b1", "eta_w2", "mu_w1", "mu_b1", "mu_w2")):
        close(a, g["default_g_" + k], 1e-9, 1e-13)


def test_headline_config_n20_pinned(golden):
    """The benchmark configuration (N = 20, 10 up / 10 down, Deta = Dmu = 50, 16 RK4 steps) on the real reference
    (oracle/gen_golden.py n20): forward quantities of the same discrete flow to rounding; the reference's
    adjoint-based gradient / Laplacian / E_loc on the same 16-step grid agree with the oracle's exact discrete
    derivatives to 3e-12 / 9e-13 / 8e-14 (measured); the adaptive reference at rtol 1e-9 against the 64-step oracle
    to ~1e-11."""
    g = golden("pipeline_n20")
    up, dn, eta, mu, ts = _model(g)
    assert len(up) == 10 and len(dn) == 10 and len(eta[0]) == 50
    x = O.cnf_generate(T(g["z0"]), eta, mu, ts, 16)
    close(x, g["rk4s16_x"], 1e-13)
    xr = T(g["rk4s16_x"])
    z, dl = O.cnf_delta_logp(xr, eta, mu, ts, 16)
    close(z, g["rk4s16_zback"], 1e-13)
    close(dl, g["rk4s16_delta_logp"], 1e-12, 1e-15)
    r = O.local_energy(xr, up, dn, eta, mu, ts, 16, float(g["Z"]))
    close(r["logp"], g["rk4s16_logp"], 1e-13)
    close(r["grad"], g["rk4s16_grad"], 1e-10)
    close(r["lap"], g["rk4s16_lap"], 1e-10)
    close(r["eloc"], g["rk4s16_eloc"], 1e-11)
    r = O.local_energy(T(g["tight_x"]), up, dn, eta, mu, ts, 64, float(g["Z"]))
    close(r["logp"], g["tight_logp"], 1e-10)
    close(r["grad"], g["tight_grad"], 1e-9)
    close(r["lap"], g["tight_lap"], 1e-9)
    close(r["eloc"], g["tight_eloc"], 1e-10)
