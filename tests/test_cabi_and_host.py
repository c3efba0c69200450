"""CPU-only checks: the C-ABI library loads and exports every symbol include/*.h declares,
the host-side mirror behaves like the reference's host logic, and nothing silently falls
back to the CPU."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
torch.set_default_dtype(torch.float64)


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "fermiflow_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(ff_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    path = g.build()
    lib = ctypes.CDLL(path)
    names = _declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "missing symbol " + n
    lib.ff_version.restype = ctypes.c_int
    assert lib.ff_version() >= 100


def test_ctypes_signatures_cover_header():
    from fermiflow_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared_symbols()


def test_no_cpu_fallback():
    from fermiflow_b200 import MLP, Backflow, CNF
    cnf = CNF(Backflow(MLP(1, 4)), (0.0, 1.0), nsteps=2)
    with pytest.raises(RuntimeError, match="CUDA"):
        cnf.generate(torch.randn(2, 3, 2))
    with pytest.raises(RuntimeError, match="CUDA"):
        Backflow(MLP(1, 4))(torch.randn(2, 3, 2))


def test_argument_errors_reported():
    """Error behaviour of the C ABI without touching a device."""
    from fermiflow_b200 import _lib as L
    lib = L.lib()
    assert lib.ff_cnf_generate(None, None, 0, 0, None, None) != 0
    assert b"model" in lib.ff_last_error()
    m = L.FFModel()
    m.n_up, m.n_dn, m.H_eta, m.nsteps = 2, 0, 0, 4
    assert lib.ff_cnf_generate(ctypes.byref(m), None, 0, 0, None, None) != 0
    assert b"hidden" in lib.ff_last_error()
    assert lib.ff_potential(None, 5, 3, 1.0, 1, None, None) != 0


def test_fermion_states_match_reference(golden):
    from fermiflow_b200 import HO2D
    ho, g = HO2D(), golden("states")
    for nup, dE in ((3, 2), (6, 2), (3, 4), (10, 2), (4, 3)):
        states, Es = ho.fermion_states(nup, 0, dE)
        idx = np.array([[o.index for o in up] for up, _ in states])
        assert np.array_equal(idx, g["states_%d_%d" % (nup, dE)])
        assert np.array_equal(np.array(Es), g["Es_%d_%d" % (nup, dE)])
    with pytest.raises(ValueError):
        ho.fermion_states(3, 1, 2)


def test_orbital_callables_match_reference(golden):
    from fermiflow_b200 import HO2D
    g = golden("slater")
    x = torch.from_numpy(g["x_orb"])
    vals = torch.stack([o(x) for o in HO2D().orbitals], -1)
    assert np.abs(vals.numpy() - g["orbitals"]).max() < 1e-13
    assert HO2D().Es == list(g["Es"])


def test_mlp_mirror_matches_reference_mlp(golden):
    """MLP.forward / grad (MLP.py:30-45) and parameter naming (state_dict compatibility)."""
    from fermiflow_b200 import MLP
    m = MLP(1, 7)
    assert sorted(m.state_dict()) == ["fc1.bias", "fc1.weight", "fc2.weight"]
    x = torch.randn(11, 1, requires_grad=True)
    y = m(x)
    gx, = torch.autograd.grad(y.sum(), x)
    assert torch.allclose(gx, m.grad(x))
    m.init_zeros()
    assert float(m(x).abs().max()) == 0.0
    with pytest.raises(ValueError):
        MLP(2, 3).kernel_params()


def test_vmc_rejects_unsupported_potentials():
    from fermiflow_b200 import MLP, Backflow, CNF, HO2D, FreeFermion, GSVMC

    class Other:
        pass
    cnf = CNF(Backflow(MLP(1, 4)), (0.0, 1.0))
    with pytest.raises(NotImplementedError):
        GSVMC(2, 0, HO2D(), FreeFermion("cpu"), cnf, Other())
