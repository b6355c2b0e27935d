"""GPU tests of the reference-facing host API (rnnwavefunctions_b200/training.py and the reference-layout shim
modules): local-energy functions with the reference's signatures against the golden vectors, and the run_*
drivers converging to exact-diagonalisation energies (the reference's only end-to-end check,
Tutorial_1DTFIM.ipynb#cell18 / Tutorial_1DJ1J2.ipynb#cell18)."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

from oracle import rnnwf_oracle as O

pytestmark = pytest.mark.gpu

from rnnwavefunctions_b200 import training as TR  # noqa: E402
from rnnwavefunctions_b200.wavefunction import (ComplexRNNwavefunction, RNNwavefunction1D, RNNwavefunction2D,  # noqa: E402
                                                RNNwavefunction2DFlat, Session)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ising_local_energies_reference_signature(golden):
    g = golden("tfim1d")
    units = [int(u) for u in g["a_units"]]
    samples, Jz, Bx = g["a_samples"], g["a_Jz"], float(g["a_Bx"])
    ns, N = samples.shape
    wf = RNNwavefunction1D(N, units=units, seed=1)
    wf.params.copy_(torch.tensor(g["a_params"].astype(np.float32), device=wf.device))
    queue = np.zeros((N + 1, ns, N), np.int32)
    lps = np.zeros((N + 1) * ns)
    e = TR.Ising_local_energies(Jz, Bx, samples, queue, wf, None, lps, Session())
    assert isinstance(e, np.ndarray) and e.dtype == np.float64 and e.shape == (ns,)
    np.testing.assert_allclose(e, g["a_eloc"], rtol=1e-5)
    assert np.array_equal(queue[0], samples)
    # the wave-function methods return device tensors; Session.run materialises them like sess.run
    lp = Session().run(wf.log_probability(samples, 2))
    np.testing.assert_allclose(lp, g["a_logprobs"][:ns], rtol=1e-5)
    s = Session().run(wf.sample(33, 2))
    assert s.shape == (33, N) and s.dtype == np.int64


def test_ising2d_local_energies_both_models(golden):
    g = golden("tfim2d")
    Jz, Bx = g["Jz"], float(g["Bx"])
    wf = RNNwavefunction2DFlat(4, 4, units=[int(g["flat_units"][0])])
    wf.params.copy_(torch.tensor(g["flat_params"], device=wf.device))
    e = TR.Ising2D_local_energies(Jz, Bx, 4, 4, g["flat_samples"], None, wf, None, None, None)
    np.testing.assert_allclose(e, g["flat_eloc"], rtol=1e-11)
    wf2 = RNNwavefunction2D(4, 4, units=[int(g["md_units"][0])])
    wf2.params.copy_(torch.tensor(g["md_params"], device=wf2.device))
    e2 = TR.Ising2D_local_energies(Jz, Bx, 4, 4, g["md_samples"], None, wf2, None, None, None)
    np.testing.assert_allclose(e2, g["md_eloc"], rtol=1e-11)
    s = wf2.sample(20, 2)
    assert tuple(s.shape) == (20, 4, 4)
    assert wf2.rnn.state_size == int(g["md_units"][0])


def test_j1j2_host_functions(golden):
    g = golden("j1j2")
    sig = g["sigmas"]
    N = sig.shape[1]
    sigmaH = np.zeros((2 * N + 1, N), np.int32)
    me = np.zeros(2 * N + 1, np.float32)
    o = 0
    for s in sig:                                             # keyword form: periodic / Marshall_sign independent
        num = TR.J1J2MatrixElements(g["J1"], g["J2"], g["Bz"], s, sigmaH, me, periodic=False, Marshall_sign=True)
        assert np.array_equal(sigmaH[:num], g["p0m1_sigmaH"][o:o + num])
        assert np.array_equal(me[:num], g["p0m1_elements"][o:o + num])
        o += num
    # J1J2Slices: reference_compat reproduces the reference's positional mis-binding (Marshall_sign -> periodic)
    ns = len(sig)
    sigmas = np.zeros(((2 * N + 1) * ns, N), np.int32)
    H = np.zeros((2 * N + 1) * ns, np.float32)
    slices, total = TR.J1J2Slices(g["J1"], g["J2"], g["Bz"], sig, sigmas, H, sigmaH, me, True, reference_compat=True)
    assert total == g["p1m0_num"].sum()
    assert np.array_equal(sigmas[:total], g["p1m0_sigmaH"]) and np.array_equal(H[:total], g["p1m0_elements"])
    slices, total = TR.J1J2Slices(g["J1"], g["J2"], g["Bz"], sig, sigmas, H, sigmaH, me, True)
    assert np.array_equal(sigmas[:total], g["p0m1_sigmaH"]) and np.array_equal(H[:total], g["p0m1_elements"])
    assert [s.stop - s.start for s in slices] == list(g["p0m1_num"])


def test_run_1dtfim_converges_to_exact(tmp_path, golden):
    exact = float(golden("known_answers")["tfim_N10_recorded"])           # -12.38148999965476 (Tutorial_1DTFIM.ipynb#cell8)
    E, V = TR.run_1DTFIM(numsteps=700, systemsize=10, num_units=10, Bx=1, num_layers=1, numsamples=500, learningrate=1e-2, seed=111,
                         checkpoint_dir=str(tmp_path), verbose=False)
    assert len(E) == 701 and len(V) == 701
    last = np.mean(E[-100:])
    assert abs(last - exact) < 0.02, (last, exact)
    assert last > exact - 0.01                                          # variational
    assert np.mean(V[-100:]) < 0.1
    files = sorted(os.listdir(tmp_path))
    assert "meanEnergy_N10_samp500_Jz1.0_Bx1_GRURNN_OBC_TFIM_units_10.npy" in files
    assert any(f.endswith(".npz") for f in files)
    # resume picks the run up from the checkpoint written at it=500 (the reference's restore path is commented out, :172-183)
    E2, _ = TR.run_1DTFIM(numsteps=520, systemsize=10, num_units=10, Bx=1, num_layers=1, numsamples=500, learningrate=1e-2, seed=111,
                          checkpoint_dir=str(tmp_path), verbose=False, resume=True)
    assert len(E2) == 521 and np.allclose(E2[:501], E[:501])
    # the checkpoint is written after update 500, so the resumed trajectory is the uninterrupted one (same Philox draws,
    # same Adam state): iterations 501..520 must reproduce the first run, which a skipped update would break
    np.testing.assert_allclose(E2[501:521], E[501:521], rtol=1e-9, atol=1e-9)


def test_run_1dtfim_tensor_core_path_converges(tmp_path):
    # 50 units: rnnwf_tfim_eloc runs the (pipelined) tcgen05 3xFP16 chain kernel; the optimisation must reach the free-fermion ground state
    from rnnwavefunctions_b200 import ops
    N = 16
    model = ops.make_model(num_layers=2, units=50, n_sites=N)
    assert ops.tfim_chain_mode(model) == 3                                # 3: pipelined 3xFP16 tcgen05 kernel (gru_tc16p.cuh)
    exact = O.tfim1d_exact_energy(N, 1.0, 1.0)
    E, V = TR.run_1DTFIM(numsteps=400, systemsize=N, num_units=50, Bx=1, num_layers=2, numsamples=500, learningrate=5e-3, seed=7,
                         save=False, verbose=False)
    last = np.mean(E[-50:])
    assert abs(last - exact) < 0.05, (last, exact)
    assert last > exact - 0.02
    assert np.mean(V[-50:]) < 0.2


def test_run_1dtfim_parity_and_multilayer(tmp_path, golden):
    exact = float(golden("known_answers")["tfim_N8"])
    E, V = TR.run_1DTFIM(numsteps=500, systemsize=8, num_units=8, Bx=1, num_layers=2, numsamples=400, learningrate=1e-2, seed=3,
                         parity_symmetric=True, save=False, verbose=False)
    assert abs(np.mean(E[-50:]) - exact) < 0.03


def ed_tfim2d(Nx, Ny, Bx):
    N = Nx * Ny
    dim = 1 << N
    idx = np.arange(dim)
    bits = (idx[:, None] >> np.arange(N)[None, :]) & 1
    s = (1 - 2 * bits).reshape(dim, Nx, Ny)
    diag = -(s[:, :-1, :] * s[:, 1:, :]).sum(axis=(1, 2)) - (s[:, :, :-1] * s[:, :, 1:]).sum(axis=(1, 2))
    Hm = np.diag(diag.astype(np.float64))
    for i in range(N):
        Hm[idx, idx ^ (1 << i)] += -Bx
    return np.linalg.eigvalsh(Hm)[0]


def test_run_2dtfim_both_models(tmp_path):
    exact = ed_tfim2d(3, 3, 3.0)
    E1, _ = TR.run_2DTFIM_1DRNN(numsteps=900, systemsize_x=3, systemsize_y=3, Bx=3, num_units=12, num_layers=1, numsamples=400,
                                learningrate=1e-2, seed=333, save=False, verbose=False)
    assert abs(np.mean(E1[-50:]) - exact) < 0.15, (np.mean(E1[-50:]), exact)      # decaying lr 1/((1/lr)+it/10): slow tail
    assert np.mean(E1[-50:]) > exact - 0.02
    E2, _ = TR.run_2DTFIM_2DRNN(numsteps=500, systemsize_x=3, systemsize_y=3, Bx=3, num_units=12, numsamples=400,
                                learningrate=5e-3, seed=111, save=False, verbose=False)
    assert abs(np.mean(E2[-50:]) - exact) < 0.15, (np.mean(E2[-50:]), exact)
    assert np.mean(E2[-50:]) > exact - 0.02


def test_run_j1j2_converges(tmp_path, golden):
    exact = float(golden("known_answers")["j1j2_N6_J2_0.2"])
    E, V = TR.run_J1J2(numsteps=1500, systemsize=6, J1_=1.0, J2_=0.2, Marshall_sign=True, num_units=10, num_layers=1,
                       numsamples=300, learningrate=5e-3, seed=111, save=False, verbose=False)
    last = np.mean(np.real(E[-100:]))
    assert isinstance(E[0], complex)
    assert abs(last - exact) < 0.05, (last, exact)


def test_reference_layout_shims_import():
    for d, mod, names in [("1DTFIM", "TrainingRNN_1DTFIM", ["Ising_local_energies", "run_1DTFIM"]),
                          ("2DTFIM_1DRNN", "Training1DRNN_2DTFIM", ["Ising2D_local_energies", "run_2DTFIM"]),
                          ("2DTFIM_2DRNN", "Training2DRNN_2DTFIM", ["Ising2D_local_energies", "run_2DTFIM"]),
                          ("J1J2", "TrainingRNN_J1J2", ["J1J2MatrixElements", "J1J2Slices", "run_J1J2"])]:
        sys.path.insert(0, os.path.join(ROOT, d))
        try:
            for m in (mod, "RNNwavefunction", "ComplexRNNwavefunction", "MDRNNcell", "RNNwavefunction_paritysym"):
                sys.modules.pop(m, None)
            M = importlib.import_module(mod)
            for n in names:
                assert callable(getattr(M, n))
            wfmod = "ComplexRNNwavefunction" if d == "J1J2" else "RNNwavefunction"
            assert hasattr(importlib.import_module(wfmod), "RNNwavefunction")
        finally:
            sys.path.pop(0)


def test_session_run_materialises_samples_and_tensors_like_sess_run():
    """`samples = sess.run(samples_)` (1DTFIM/TrainingRNN_1DTFIM.py:203): int64 host array equal to the device tensor, through the
    pinned one-byte-per-site path; other tensors and nested fetches come back as NumPy arrays too."""
    from rnnwavefunctions_b200.wavefunction import RNNwavefunction1D, RNNwavefunction2D, Session
    sess = Session()
    wf = RNNwavefunction1D(12, units=[10], seed=3)
    t = wf.sample(300, 2)
    h = sess.run(t)
    assert h.dtype == np.int64 and h.shape == (300, 12) and np.array_equal(h, t.cpu().numpy())
    h2 = sess.run(wf.sample(200, 2))                      # the staging buffer is reused: earlier results must not change
    assert np.array_equal(h, t.cpu().numpy()) and h2.shape == (200, 12)
    lp = wf.log_probability(h)
    a, (b,) = sess.run([lp, (lp * 2,)])
    assert np.array_equal(a, lp.cpu().numpy()) and np.array_equal(b, 2 * a)
    wf2 = RNNwavefunction2D(3, 4, units=[6], seed=3)
    t2 = wf2.sample(50, 2)
    assert np.array_equal(sess.run(t2), t2.cpu().numpy()) and sess.run(t2).shape == (50, 3, 4)
