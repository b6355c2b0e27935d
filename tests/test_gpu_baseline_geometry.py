"""Parity at the geometry of every BASELINE.json configuration: the CUDA path through the C ABI
against the oracle's full recompute / CPU autograd on a few samples each, at the sizes the benchmark times.

    cfg1  1-D TFIM N=20, 1 x GRU(50), 500 samples           (1DTFIM/run_1dTFIM.py as shipped)
    cfg2  1-D TFIM N=1000, 3 x GRU(50), +- parity symmetry  (1DTFIM/RNNwavefunction.py, RNNwavefunction_paritysym.py)
    cfg3  2-D TFIM 12x12, 1-D GRU(100), float64              (2DTFIM_1DRNN/RNNwavefunction.py:86-130)
    cfg4  2-D TFIM 12x12, MDRNN(100), float64                (2DTFIM_2DRNN/RNNwavefunction.py:120-200)
    cfg5  J1-J2 N=100, J2=0.2, cRNN GRU(50), Marshall sign   (J1J2/TrainingRNN_J1J2.py:255-279)

Tolerances: 1e-5 relative for FP32 log-probabilities / local energies (north star), 1e-4 relative L2 for FP32 gradients,
1e-10 / 1e-8 for the float64 models."""
import os

import numpy as np
import pytest
import torch

from oracle import rnnwf_oracle as O
from oracle import torch_grad as TG

pytestmark = pytest.mark.gpu

from rnnwavefunctions_b200 import ops, params as P  # noqa: E402

CHEADS = ("wf_dense_ampl", "wf_dense_phase")


def dev():
    return torch.device("cuda:0")


def u8(samples):
    return torch.as_tensor(np.asarray(samples).reshape(len(samples), -1).astype(np.uint8), device=dev())


def p64(p):
    return {k: v.astype(np.float64) for k, v in p.items()}


# ------------------------------------------------------------------------------------------------------------------
# cfg2: N = 1000, 3 x GRU(50)
# ------------------------------------------------------------------------------------------------------------------
N2, L2, H2 = 1000, 3, 50


@pytest.fixture(scope="module")
def cfg2():
    model = ops.make_model(num_layers=L2, units=H2, n_sites=N2)
    flat_h = P.init_flat(P.gru_shapes([H2] * L2), 111, np.float32)       # the benchmark's weights (bench.py)
    flat = torch.tensor(flat_h, device=dev())
    p = O.unflatten(flat_h, O.gru_param_shapes([H2] * L2), np.float32)
    samples = ops.sample(model, flat, 640, seed=21)
    return model, flat, p, samples


def test_cfg2_default_chain_kernel_is_the_pipelined_tensor_core_one(cfg2):
    model = cfg2[0]
    assert "RNNWF_CHAIN" not in os.environ
    assert ops.tfim_chain_mode(model) == 3


@pytest.mark.parametrize("ns", [300, 640])
def test_cfg2_default_kernel_vs_cuda_core_chains_multi_tile(cfg2, ns):
    """The shipped default (tc16p: several 128-row tiles, partly filled last tile at 300, more work items than CTAs, longest-first
    atomic scheduling) against the FFMA engine on the same samples."""
    model, flat, _, samples = cfg2
    s = samples[:ns]
    e, lp = ops.tfim_eloc(model, flat, s, np.ones(N2), 1.0)
    os.environ["RNNWF_CHAIN"] = "ffma"
    try:
        assert ops.tfim_chain_mode(model) == 0
        e0, lp0 = ops.tfim_eloc(model, flat, s, np.ones(N2), 1.0)
    finally:
        os.environ.pop("RNNWF_CHAIN", None)
    np.testing.assert_allclose(e.cpu().numpy(), e0.cpu().numpy(), rtol=1e-5)
    np.testing.assert_allclose(lp.cpu().numpy(), lp0.cpu().numpy(), rtol=1e-5)
    # rows must not depend on which tile / work item they were evaluated in
    e2, _ = ops.tfim_eloc(model, flat, s[128:128 + 100].contiguous(), np.ones(N2), 1.0)
    np.testing.assert_allclose(e2.cpu().numpy(), e.cpu().numpy()[128:228], rtol=1e-6)


def test_cfg2_parity_symmetric_eloc_and_logprob_vs_oracle(cfg2):
    """1DTFIM/RNNwavefunction_paritysym.py:125-145 at N=1000: exp(lp) underflows there, the oracle and the kernel both
    use the log-add-exp form."""
    model, flat, p, samples = cfg2
    s = samples[:2]
    sh = s.cpu().numpy().astype(np.int64)
    ref_lp = O.log_probability_parity(p, sh)
    ref_e = O.ising_local_energies(np.ones(N2), 1.0, sh, lambda c: O.log_probability_parity(p, c))
    e, lp = ops.tfim_eloc(model, flat, s, np.ones(N2), 1.0, flags=ops.PARITY_SYM)
    np.testing.assert_allclose(lp.cpu().numpy(), ref_lp, rtol=1e-5)
    np.testing.assert_allclose(e.cpu().numpy(), ref_e, rtol=1e-5)
    np.testing.assert_allclose(ops.logpsi(model, flat, s, flags=ops.PARITY_SYM).cpu().numpy(), ref_lp, rtol=1e-5)
    # parity symmetry itself: the mirrored configuration has the same amplitude
    lpm = ops.logpsi(model, flat, torch.flip(s, dims=[1]).contiguous(), flags=ops.PARITY_SYM)
    np.testing.assert_allclose(lpm.cpu().numpy(), lp.cpu().numpy(), rtol=1e-6)


@pytest.mark.parametrize("parity", [False, True])
def test_cfg2_gradient_vs_cpu_autograd(cfg2, parity):
    """rnnwf_vmc_grad (stash pass, backward recurrence, tcgen05 weight-gradient reduction with its FP64 flushes every 32 blocks)
    against float64 CPU autograd on 3 samples of 1000 sites."""
    model, flat, p, samples = cfg2
    s = samples[:3]
    w = np.array([0.7, -1.1, 0.4])
    ref = TG.gru_vmc_grad(p64(p), s.cpu().numpy().astype(np.int64), w, parity=parity)
    got = ops.vmc_grad(model, flat, s, torch.tensor(w, device=dev()), flags=ops.PARITY_SYM if parity else 0).cpu().numpy()
    err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    assert err < 1e-4, err


def test_cfg2_gradient_many_tiles_is_the_sum_of_its_parts(cfg2):
    """Gradient over 384 samples (three tiles, 3000 (tile, site) blocks = many TMEM flushes) equals the sum over three separate
    128-sample calls, and a 3-sample piece of it equals the oracle-checked value above."""
    model, flat, _, samples = cfg2
    rng = np.random.default_rng(2)
    w = torch.tensor(rng.normal(size=384), device=dev())
    g = ops.vmc_grad(model, flat, samples[:384], w)
    parts = sum(ops.vmc_grad(model, flat, samples[i:i + 128].contiguous(), w[i:i + 128].contiguous()) for i in (0, 128, 256))
    assert float((g - parts).norm() / g.norm()) < 1e-6


# ------------------------------------------------------------------------------------------------------------------
# trained-magnitude weights through the 3 x FP16 split (weak #4): saturated gates, |w| >> 1
# ------------------------------------------------------------------------------------------------------------------
def test_trained_weights_eloc_vs_oracle():
    from rnnwavefunctions_b200 import training as TR
    from rnnwavefunctions_b200.vmc import TFIM, VMC
    from rnnwavefunctions_b200.wavefunction import RNNwavefunction1D
    N, units = 24, [50, 50]
    wf = RNNwavefunction1D(N, units=units, seed=5, device="cuda:0")
    assert ops.tfim_chain_mode(wf.model) == 3
    opt = VMC(wf, TFIM(np.ones(N), 1.0), 500)
    for it in range(600):
        opt.step(1e-2)
    flat_h = wf.params.cpu().numpy()
    p = O.unflatten(flat_h, O.gru_param_shapes(units), np.float32)
    assert np.abs(flat_h).max() > 1.5                                   # well away from the glorot range (|w| < 0.25)
    s = opt.draw()[:40]
    sh = s.cpu().numpy().astype(np.int64)
    ref = O.ising_local_energies(np.ones(N), 1.0, sh, lambda c: O.log_probability(p, c))
    e, lp = ops.tfim_eloc(wf.model, wf.params, s, np.ones(N), 1.0)
    np.testing.assert_allclose(e.cpu().numpy(), ref, rtol=1e-5)
    np.testing.assert_allclose(lp.cpu().numpy(), O.log_probability(p, sh), rtol=1e-5)
    # and with the weights blown up further (x4: gates saturate, lo halves of small products go subnormal in FP16)
    big = {k: (v * 4).astype(np.float32) if k.endswith("kernel") else v for k, v in p.items()}
    flat_big = torch.tensor(O.flatten(big), device=dev())
    sb = ops.sample(wf.model, flat_big, 24, seed=3)
    sbh = sb.cpu().numpy().astype(np.int64)
    refb = O.ising_local_energies(np.ones(N), 1.0, sbh, lambda c: O.log_probability(big, c))
    eb, lpb = ops.tfim_eloc(wf.model, flat_big, sb, np.ones(N), 1.0)
    np.testing.assert_allclose(lpb.cpu().numpy(), O.log_probability(big, sbh), rtol=1e-5)
    np.testing.assert_allclose(eb.cpu().numpy(), refb, rtol=1e-5, atol=1e-5)
    del TR


# ------------------------------------------------------------------------------------------------------------------
# cfg3: 12 x 12, 1-D GRU(100), float64
# ------------------------------------------------------------------------------------------------------------------
def test_cfg3_gru100_f64_logprob_eloc_gradient_vs_oracle():
    Nx = Ny = 12
    N, Hh = Nx * Ny, 100
    p = O.randomize_biases(O.init_gru_params([Hh], seed=333, dtype=np.float64, scale=1.0), seed=334)
    model = ops.make_model(dtype=ops.F64, num_layers=1, units=Hh, n_sites=N, nx=Nx, ny=Ny)
    flat = torch.tensor(O.flatten(p), device=dev())
    assert flat.numel() == ops.param_count(model) == 31202              # P of cfg3
    s = ops.sample(model, flat, 150, seed=4)
    sh = s.cpu().numpy().astype(np.int64)
    ref_lp = O.log_probability(p, sh)
    np.testing.assert_allclose(ops.logpsi(model, flat, s).cpu().numpy(), ref_lp, rtol=1e-11)
    Jz = np.random.default_rng(6).uniform(0.5, 1.5, size=(Nx, Ny))
    k = 6
    ref_e = O.ising2d_local_energies(Jz, 2.0, Nx, Ny, sh[:k], lambda c: O.log_probability(p, c), flat=True)
    e, lp = ops.tfim_eloc(model, flat, s[:k], Jz, 2.0)
    np.testing.assert_allclose(e.cpu().numpy(), ref_e, rtol=1e-10)
    np.testing.assert_allclose(lp.cpu().numpy(), ref_lp[:k], rtol=1e-11)
    # whole batch (two row tiles): rows independent of their tile
    e_all, _ = ops.tfim_eloc(model, flat, s, Jz, 2.0)
    np.testing.assert_allclose(e_all.cpu().numpy()[:k], ref_e, rtol=1e-10)
    w = np.random.default_rng(7).normal(size=10)
    ref_g = TG.gru_vmc_grad(p, sh[:10], w)
    got = ops.vmc_grad(model, flat, s[:10], torch.tensor(w, device=dev())).cpu().numpy()
    np.testing.assert_allclose(got, ref_g, rtol=1e-8, atol=1e-10 * np.abs(ref_g).max())


# ------------------------------------------------------------------------------------------------------------------
# cfg4: 12 x 12, MDRNN(100), float64
# ------------------------------------------------------------------------------------------------------------------
def test_cfg4_mdrnn100_eloc_and_gradient_at_12x12_vs_oracle():
    Nx = Ny = 12
    Hh = 100
    p = O.randomize_biases(O.init_mdrnn_params(Hh, seed=111, dtype=np.float64, scale=0.5), seed=112)
    model = ops.make_model(cell=ops.CELL_MDRNN, dtype=ops.F64, num_layers=1, units=Hh, n_sites=Nx * Ny, nx=Nx, ny=Ny)
    flat = torch.tensor(O.flatten(p), device=dev())
    assert flat.numel() == ops.param_count(model) == 20702
    s = O.mdrnn_sample(p, 4, Nx, Ny, seed=8)
    Jz = np.ones((Nx, Ny))
    ref = O.ising2d_local_energies(Jz, 2.0, Nx, Ny, s[:2], lambda c: O.mdrnn_log_probability(p, c), flat=False)
    e, lp = ops.tfim_eloc(model, flat, u8(s[:2]), Jz, 2.0)
    np.testing.assert_allclose(e.cpu().numpy(), ref, rtol=1e-10)
    np.testing.assert_allclose(lp.cpu().numpy(), O.mdrnn_log_probability(p, s[:2]), rtol=1e-11)
    w = np.array([0.3, -0.9, 1.2, 0.1])
    ref_g = TG.mdrnn_vmc_grad(p, s, w)
    got = ops.vmc_grad(model, flat, u8(s), torch.tensor(w, device=dev())).cpu().numpy()
    np.testing.assert_allclose(got, ref_g, rtol=1e-8, atol=1e-10 * np.abs(ref_g).max())


# ------------------------------------------------------------------------------------------------------------------
# cfg5: J1-J2 N = 100, J2 = 0.2, cRNN GRU(50), Marshall sign
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("marshall", [True, False])
def test_cfg5_crnn_n100_eloc_logamp_gradient_vs_oracle(marshall):
    N, Hh = 100, 50
    p = O.randomize_biases(O.init_gru_params([Hh], seed=111, dtype=np.float32, heads=CHEADS, scale=1.5), seed=112)
    model = ops.make_model(cell=ops.CELL_GRU, head=ops.HEAD_COMPLEX, dtype=ops.F32, num_layers=1, units=Hh, n_sites=N)
    flat = torch.tensor(O.flatten(p), device=dev())
    assert flat.numel() == ops.param_count(model) == 8204               # P of cfg5
    s = ops.sample(model, flat, 200, seed=2)
    sh = s.cpu().numpy().astype(np.int64)
    assert (sh.sum(axis=1) == N // 2).all()                              # zero magnetisation
    la = ops.logpsi(model, flat, s).cpu().numpy()
    ref_la = O.crnn_log_amplitude(p, sh)
    np.testing.assert_allclose(la.real, ref_la.real, rtol=1e-5)
    np.testing.assert_allclose(la.imag, ref_la.imag, rtol=1e-5, atol=1e-5)
    J1, J2, Bz = np.ones(N), 0.2 * np.ones(N), np.zeros(N)
    k = 5
    ref_e = O.j1j2_local_energies(J1, J2, Bz, sh[:k], lambda c: O.crnn_log_amplitude(p, c), marshall_sign=marshall)
    e, la2 = ops.j1j2_eloc(model, flat, s[:k], J1, J2, Bz, marshall_sign=marshall)
    e = e.cpu().numpy()
    assert np.abs(e - ref_e).max() < 2e-5 * max(1.0, np.abs(ref_e).max())      # reference combine is complex64
    np.testing.assert_allclose(la2.cpu().numpy().real, ref_la[:k].real, rtol=1e-5)
    # rows of a multi-tile launch agree with the few-sample launch
    e_all, _ = ops.j1j2_eloc(model, flat, s, J1, J2, Bz, marshall_sign=marshall)
    assert np.abs(e_all.cpu().numpy()[:k] - e).max() < 1e-6 * max(1.0, np.abs(e).max())
    if marshall:
        rng = np.random.default_rng(3)
        ew = rng.normal(size=8) + 1j * rng.normal(size=8)
        w = 2.0 * (ew - ew.mean()) / 8
        ref_g = TG.crnn_vmc_grad(p64(p), sh[:8], w)
        got = ops.vmc_grad(model, flat, s[:8], torch.tensor(w, device=dev())).cpu().numpy()
        err = np.linalg.norm(got - ref_g) / np.linalg.norm(ref_g)
        assert err < 1e-4, err


# ------------------------------------------------------------------------------------------------------------------
# cfg1: the shipped run script (N = 20, 50 units, 500 samples) converges to the free-fermion / DMRG energy
# ------------------------------------------------------------------------------------------------------------------
def test_cfg1_run_1dtfim_n20_converges_to_dmrg_energy():
    """1DTFIM/run_1dTFIM.py: run_1DTFIM(numsteps, systemsize=20, num_units=50, Bx=1, num_layers=1, numsamples=500, lr 5e-3, seed 111);
    Tutorial_1DTFIM.ipynb#cell24 gives E = -25.1077971081 (SURVEY 8d gate v).  The reference trains 10^4 steps; 3000 steps get the
    energy within statistical error of the exact value plus the variational bias of a 50-unit GRU (a few 1e-4 relative)."""
    from rnnwavefunctions_b200 import training as TR
    exact = -25.1077971081
    assert abs(O.tfim1d_exact_energy(20, 1.0, 1.0) - exact) < 1e-6      # free fermions agree with the DMRG table
    E, V = TR.run_1DTFIM(numsteps=3000, systemsize=20, num_units=50, Bx=1, num_layers=1, numsamples=500, learningrate=5e-3, seed=111,
                         save=False, verbose=False)
    last = np.asarray(E[-100:])
    mean = last.mean()
    sigma = np.sqrt(np.mean(V[-100:]) / 500 / 100)                      # standard error of the mean of the last 100 iterations
    print(f"cfg1: E = {mean:.6f} +- {sigma:.6f}, exact {exact:.6f}, var {np.mean(V[-100:]):.5f}")
    assert mean > exact - 3 * sigma                                     # variational: not below the ground state beyond noise
    assert mean - exact < 3 * sigma + 2e-4 * abs(exact), (mean, sigma, exact)
    assert np.mean(V[-100:]) < 0.05
