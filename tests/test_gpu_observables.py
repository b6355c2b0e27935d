"""GPU tests of rnnwavefunctions_b200/observables.py and rnnwf_tfim_flip_ratios.
The exact values come from enumerating the RNN state itself: psi(sigma) = sqrt(P(sigma)) over all 2^N configurations with the
oracle's log_probability, then oracle/ed.py's state-vector observables.  Per-sample amplitude ratios are deterministic and held to
the 1e-5 gate; sample averages to 5 standard errors."""
import numpy as np
import pytest
import torch

from oracle import ed
from oracle import rnnwf_oracle as O

pytestmark = pytest.mark.gpu

from rnnwavefunctions_b200 import observables as OBS, ops  # noqa: E402
from rnnwavefunctions_b200.wavefunction import ComplexRNNwavefunction, RNNwavefunction1D, RNNwavefunction2D, RNNwavefunctionParity  # noqa: E402


def _exact_state(wf, N):
    p = O.unflatten(wf.params.cpu().numpy(), wf.shapes, np.float32)
    conf = O.all_configs(N)
    lp = O.log_probability(p, conf)
    return p, conf, np.exp(0.5 * lp)


@pytest.mark.parametrize("units,N", [([50], 10), ([50, 50, 50], 9), ([12, 12], 8)])
def test_flip_ratios_match_oracle(units, N):
    wf = RNNwavefunction1D(N, units=units, seed=7, device="cuda:0")
    wf.params.mul_(2.0)                                    # away from the near-uniform random-init state
    p, conf, psi = _exact_state(wf, N)
    s = wf.sample(300)
    s_h = np.asarray(s.cpu() if isinstance(s, torch.Tensor) else s).astype(np.int64)
    Jz = np.ones(N)
    e, lp, ratios = ops.tfim_flip_ratios(wf.model, wf.params, wf._u8(s_h), Jz, 1.0)
    lp0 = O.log_probability(p, s_h)
    want = np.empty((len(s_h), N))
    for k in range(N):
        f = s_h.copy()
        f[:, k] ^= 1
        want[:, k] = np.exp(0.5 * (O.log_probability(p, f) - lp0))
    np.testing.assert_allclose(ratios.cpu().numpy(), want, rtol=1e-5)
    np.testing.assert_allclose(lp.cpu().numpy(), lp0, rtol=1e-5)
    # the ratios are the terms of the reference's local energy (1DTFIM/TrainingRNN_1DTFIM.py:70-74)
    e_ref = O.ising_local_energies(Jz, 1.0, s_h, lambda c: O.log_probability(p, c))
    np.testing.assert_allclose(e.cpu().numpy(), e_ref, rtol=1e-5)
    np.testing.assert_allclose(O.tfim1d_diag(Jz, s_h) - ratios.sum(1).cpu().numpy(), e_ref, rtol=1e-5)


def test_observables_against_exact_enumeration():
    N = 8
    wf = RNNwavefunction1D(N, units=[20], seed=3, device="cuda:0")
    wf.params.mul_(2.5)
    _, _, psi = _exact_state(wf, N)
    assert abs((psi ** 2).sum() - 1) < 1e-5                # the autoregressive model is normalised
    ns = 40000
    s = wf.sample(ns)
    m, c = OBS.sz_moments(s, device="cuda:0")
    m_ex, c_ex = ed.sz_moments(psi)
    assert np.abs(m.cpu().numpy() - m_ex).max() < 5 / np.sqrt(ns)
    assert np.abs(c.cpu().numpy() - c_ex).max() < 5 / np.sqrt(ns)
    np.testing.assert_allclose(OBS.sz_connected(s, device="cuda:0").cpu().numpy(), c.cpu().numpy() - np.outer(m.cpu().numpy(), m.cpu().numpy()), atol=1e-12)
    sx, err = OBS.sigma_x(wf, s, return_error=True)
    assert (np.abs(sx.cpu().numpy() - ed.sigma_x(psi)) < 5 * err.cpu().numpy() + 1e-6).all()
    for n_A in (1, 4, 6):
        s2, e2 = OBS.renyi2_entropy(wf, s, n_A, return_error=True)
        assert abs(s2 - ed.renyi2(psi, n_A)) < 5 * e2 + 1e-6, (n_A, s2, ed.renyi2(psi, n_A), e2)


def test_sigma_x_parity_model_and_2d_rnn():
    # parity-symmetric model: ratios of the symmetrised amplitude
    N = 8
    wf = RNNwavefunctionParity(N, units=[16], seed=5, device="cuda:0")
    wf.params.mul_(2.0)
    p = O.unflatten(wf.params.cpu().numpy(), wf.shapes, np.float32)
    conf = O.all_configs(N)
    psi = np.exp(0.5 * O.log_probability_parity(p, conf))
    s = wf.sample(20000)
    sx, err = OBS.sigma_x(wf, s, return_error=True)
    assert (np.abs(sx.cpu().numpy() - ed.sigma_x(psi)) < 5 * err.cpu().numpy() + 1e-6).all()
    # 2-D RNN (MDRNN, float64): site order x-major as the reference's slots (Training2DRNN_2DTFIM.py:55-61)
    wf2 = RNNwavefunction2D(3, 3, units=[10], seed=11, device="cuda:0")
    wf2.params.mul_(1.5)
    p2 = O.unflatten(wf2.params.cpu().numpy(), wf2.shapes, np.float64)
    conf2 = O.all_configs(9).reshape(-1, 3, 3)
    psi2 = np.exp(0.5 * O.mdrnn_log_probability(p2, conf2))
    s2 = wf2.sample(20000)
    sx2, err2 = OBS.sigma_x(wf2, s2, return_error=True)
    assert (np.abs(sx2.cpu().numpy() - ed.sigma_x(psi2)) < 5 * err2.cpu().numpy() + 1e-6).all()
    r2, e2 = OBS.renyi2_entropy(wf2, s2, 4, return_error=True)
    assert abs(r2 - ed.renyi2(psi2, 4)) < 5 * e2 + 1e-6


def test_renyi2_complex_rnn():
    N = 8
    wf = ComplexRNNwavefunction(N, units=[12], seed=9, device="cuda:0")
    wf.params.mul_(2.0)
    p = O.unflatten(wf.params.cpu().numpy(), wf.shapes, np.float32)
    conf = O.all_configs(N)
    la = O.crnn_log_amplitude(p, conf)
    psi = np.where(np.isfinite(la.real), np.exp(la), 0.0)
    assert abs((np.abs(psi) ** 2).sum() - 1) < 1e-5
    s = wf.sample(40000)
    r2, e2 = OBS.renyi2_entropy(wf, s, 4, return_error=True)
    assert abs(r2 - ed.renyi2(psi, 4)) < 5 * e2 + 1e-6, (r2, ed.renyi2(psi, 4), e2)
