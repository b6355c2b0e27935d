"""CPU tests of the known-answer tooling (oracle/ed.py): the sparse-Lanczos restatement of the
notebooks' dense ED helpers must reproduce the energies recorded from those helpers (tests/golden/known_answers.npz,
Tutorial_1DTFIM.ipynb#cell6/#cell8, Tutorial_1DJ1J2.ipynb#cell6/#cell8) and the free-fermion formula."""
import numpy as np
import pytest

from oracle import ed
from oracle import rnnwf_oracle as O


def test_tfim_chain_matches_notebook_ed_and_free_fermions(golden):
    ka = golden("known_answers")
    for N in (4, 6, 8):
        e, psi = ed.tfim_ground_state(np.ones(N), 1.0)
        assert abs(e - float(ka[f"tfim_N{N}"])) < 1e-8
        assert abs(e - O.tfim1d_exact_energy(N)) < 1e-9
        assert abs(np.linalg.norm(psi) - 1) < 1e-10 and (psi > -1e-12).all()
    e10, _ = ed.tfim_ground_state(np.ones(10), 1.0)
    assert abs(e10 - float(ka["tfim_N10_recorded"])) < 1e-7
    e12, _ = ed.tfim_ground_state(np.ones(12), 0.7)                  # sparse path (dim 4096), inhomogeneous check below
    assert abs(e12 - O.tfim1d_exact_energy(12, Bx=0.7)) < 1e-8


def test_j1j2_matches_notebook_ed(golden):
    ka = golden("known_answers")
    for N in (4, 6, 8):
        for marshall in (False, True):       # the Marshall rotation is unitary: same spectrum
            e, psi = ed.j1j2_ground_state(N, 1.0, 0.2, marshall)
            assert abs(e - float(ka[f"j1j2_N{N}_J2_0.2"])) < 1e-8
    e, psi = ed.j1j2_ground_state(10, 1.0, 0.2, True)
    assert abs(e - float(ka["j1j2_N10_J2_0.2_recorded"])) < 1e-6
    # with the Marshall sign the J2 = 0 ground state is non-negative in the zero-magnetisation sector
    _, psi0 = ed.j1j2_ground_state(8, 1.0, 0.0, True)
    psi0 = psi0 if psi0.sum() > 0 else -psi0
    assert (psi0 > -1e-10).all()


def test_tfim_lattice_and_energy_sum_rule():
    Jz = np.ones((3, 3))
    e, psi = ed.tfim_ground_state(Jz, 3.0)
    # dense cross-check with an independent construction
    N, dim = 9, 512
    idx = np.arange(dim)
    bits = (idx[:, None] >> np.arange(N - 1, -1, -1)[None, :]) & 1
    s = (2 * bits - 1).reshape(dim, 3, 3)
    Hm = np.diag((-(s[:, :-1, :] * s[:, 1:, :]).sum(axis=(1, 2)) - (s[:, :, :-1] * s[:, :, 1:]).sum(axis=(1, 2))).astype(float))
    for i in range(N):
        Hm[idx, idx ^ (1 << i)] += -3.0
    assert abs(e - np.linalg.eigvalsh(Hm)[0]) < 1e-9
    # E0 = -sum_bonds <sz sz> - Bx sum <sx> from the exact observables
    m, c = ed.sz_moments(psi)
    zz = sum(c[i * 3 + j, (i + 1) * 3 + j] for i in range(2) for j in range(3)) + sum(c[i * 3 + j, i * 3 + j + 1] for i in range(3) for j in range(2))
    assert abs(-zz - 3.0 * ed.sigma_x(psi).sum() - e) < 1e-9
    assert np.abs(m).max() < 1e-8                                     # Z2 symmetric ground state


def test_renyi2_known_states():
    N = 6
    prod = np.zeros(1 << N); prod[5] = 1.0
    assert abs(ed.renyi2(prod, 3)) < 1e-12                            # product state
    ghz = np.zeros(1 << N); ghz[0] = ghz[-1] = 2 ** -0.5
    assert abs(ed.renyi2(ghz, 3) - np.log(2)) < 1e-12                 # GHZ: one bit of entanglement
    assert abs(ed.renyi2(ghz, 1) - np.log(2)) < 1e-12
