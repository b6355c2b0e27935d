"""CPU tests: pin the oracle against the reference's golden vectors and known answers."""
import math

import numpy as np
import pytest

from oracle import rnnwf_oracle as O


def _gru_from_flat(flat, units, dtype, heads=("wf_dense",)):
    return O.unflatten(flat, O.gru_param_shapes([int(u) for u in units], 2, heads), dtype)


def test_param_counts_match_notebooks():
    # Tutorial_1DTFIM.ipynb#cell15: 422 parameters; Tutorial_1DJ1J2.ipynb#cell15: 444
    assert O.num_params(O.gru_param_shapes([10])) == 422
    assert O.num_params(O.gru_param_shapes([10], heads=("wf_dense_ampl", "wf_dense_phase"))) == 444
    assert O.num_params(O.gru_param_shapes([50, 50, 50])) == 38502
    assert O.num_params(O.gru_param_shapes([100])) == 31202            # cfg3
    assert O.num_params(O.mdrnn_param_shapes(100)) == 20702            # cfg4


def test_gru_cell_matches_torch_grucell():
    torch = pytest.importorskip("torch")
    p = O.randomize_biases(O.init_gru_params([7], seed=1, dtype=np.float64, scale=2.0))
    base = "RNNwavefunction/multi_rnn_cell/cell_0/cudnn_compatible_gru_cell/"
    cell = torch.nn.GRUCell(2, 7).double()
    Kg, bg = p[base + "gates/kernel"], p[base + "gates/bias"]
    with torch.no_grad():
        # torch gate order r,z,n ; TF columns [r|u]
        cell.weight_ih.copy_(torch.tensor(np.concatenate([Kg[:2].T, p[base + "candidate/input_projection/kernel"].T])))
        cell.weight_hh.copy_(torch.tensor(np.concatenate([Kg[2:].T, p[base + "candidate/hidden_projection/kernel"].T])))
        cell.bias_ih.copy_(torch.tensor(np.concatenate([bg, p[base + "candidate/input_projection/bias"]])))
        cell.bias_hh.copy_(torch.tensor(np.concatenate([0 * bg, p[base + "candidate/hidden_projection/bias"]])))
    rng = np.random.default_rng(0)
    x = rng.normal(size=(5, 2))
    h = rng.uniform(-1, 1, size=(5, 7))
    ours = O.gru_cell(p, 0, x, h)
    ref = cell(torch.tensor(x), torch.tensor(h)).detach().numpy()
    np.testing.assert_allclose(ours, ref, rtol=1e-12, atol=1e-12)


def test_tfim1d_golden(golden):
    g = golden("tfim1d")
    for tag in "abc":
        p = _gru_from_flat(g[f"{tag}_params"], g[f"{tag}_units"], np.float32)
        samples, Jz, Bx = g[f"{tag}_samples"], g[f"{tag}_Jz"], float(g[f"{tag}_Bx"])
        assert np.array_equal(O.tfim1d_diag(Jz, samples), g[f"{tag}_diag"])          # bit-exact
        if Bx != 0:
            assert np.array_equal(O.tfim1d_queue(samples), g[f"{tag}_queue"])        # bit-exact
        e = O.ising_local_energies(Jz, Bx, samples, lambda c: O.log_probability(p, c))
        np.testing.assert_allclose(e, g[f"{tag}_eloc"], rtol=1e-13, atol=0)
    big = np.unpackbits(g["big_samples"], axis=1)[:, :1000].astype(np.int32)
    assert np.array_equal(O.tfim1d_diag(np.ones(1000), big), g["big_diag"])


def test_tfim2d_golden(golden):
    g = golden("tfim2d")
    Jz, Bx = g["Jz"], float(g["Bx"])
    p = _gru_from_flat(g["flat_params"], g["flat_units"], np.float64)
    e = O.ising2d_local_energies(Jz, Bx, 4, 4, g["flat_samples"], lambda c: O.log_probability(p, c), flat=True)
    np.testing.assert_allclose(e, g["flat_eloc"], rtol=1e-13)
    np.testing.assert_allclose(O.log_probability(p, g["flat_queue"].reshape(-1, 16)), g["flat_logprobs"], rtol=1e-13)
    p2 = O.unflatten(g["md_params"], O.mdrnn_param_shapes(int(g["md_units"][0])), np.float64)
    e2 = O.ising2d_local_energies(Jz, Bx, 4, 4, g["md_samples"], lambda c: O.mdrnn_log_probability(p2, c), flat=False)
    np.testing.assert_allclose(e2, g["md_eloc"], rtol=1e-13)
    q = O.tfim1d_queue(g["md_samples"].reshape(5, 16)).reshape(17, 5, 4, 4)
    assert np.array_equal(q, g["md_queue"])


def test_j1j2_golden(golden):
    g = golden("j1j2")
    for periodic in (False, True):
        for marshall in (False, True):
            tag = f"p{int(periodic)}m{int(marshall)}"
            rows, els, nums = [], [], []
            for s in g["sigmas"]:
                c, h = O.j1j2_matrix_elements(g["J1"], g["J2"], g["Bz"], s, periodic, marshall)
                rows.append(c), els.append(h), nums.append(len(h))
            assert np.array_equal(np.asarray(nums), g[f"{tag}_num"])
            assert np.array_equal(np.concatenate(rows), g[f"{tag}_sigmaH"])
            assert np.array_equal(np.concatenate(els), g[f"{tag}_elements"])        # bit-exact float32
    p = _gru_from_flat(g["full_params"], g["full_units"], np.float32, ("wf_dense_ampl", "wf_dense_phase"))
    N = g["full_samples"].shape[1]
    e = O.j1j2_local_energies(np.ones(N), float(g["full_J2"]) * np.ones(N), np.zeros(N), g["full_samples"],
                              lambda c: O.crnn_log_amplitude(p, c))
    np.testing.assert_allclose(e, g["full_eloc"], rtol=1e-6)
    assert (g["full_samples"].sum(1) == N // 2).all()


def test_known_energies(golden):
    g = golden("known_answers")
    for N in (4, 6, 8):
        assert abs(O.tfim1d_exact_energy(N) - float(g[f"tfim_N{N}"])) < 1e-10
    assert abs(O.tfim1d_exact_energy(10) - float(g["tfim_N10_recorded"])) < 1e-9
    # DMRG table, Tutorial_1DTFIM.ipynb#cell24
    assert abs(O.tfim1d_exact_energy(20) - (-25.1077971081)) < 1e-7
    assert abs(O.tfim1d_exact_energy(1000) - (-1272.8762945220)) < 2e-6


def test_normalisation_bruteforce():
    N = 8
    cfg = O.all_configs(N)
    p = O.randomize_biases(O.init_gru_params([5, 4], seed=4, dtype=np.float32, scale=3.0))
    assert abs(np.exp(O.log_probability(p, cfg)).sum() - 1) < 1e-5
    assert abs(np.exp(O.log_probability_parity(p, cfg)).sum() - 1) < 1e-5
    lp = O.log_probability_parity(p, cfg)
    np.testing.assert_allclose(lp, O.log_probability_parity(p, cfg[:, ::-1]), rtol=1e-12)
    np.testing.assert_allclose(lp, O.log_probability_parity(p, cfg, reference_exact=True), rtol=1e-12)
    pm = O.init_mdrnn_params(5, seed=2, dtype=np.float64, scale=2.0)
    cfg2 = O.all_configs(9).reshape(-1, 3, 3)
    assert abs(np.exp(O.mdrnn_log_probability(pm, cfg2)).sum() - 1) < 1e-10
    pc = O.randomize_biases(O.init_gru_params([6], seed=9, dtype=np.float32, heads=("wf_dense_ampl", "wf_dense_phase"), scale=3.0))
    la = O.crnn_log_amplitude(pc, cfg)
    sector = cfg.sum(1) == N // 2
    assert abs(np.exp(2 * la.real[sector]).sum() - 1) < 1e-5
    assert np.all(np.isneginf(la.real[~sector]) | (la.real[~sector] < -60))


def test_sampling_statistics():
    N = 6
    p = O.randomize_biases(O.init_gru_params([5], seed=8, dtype=np.float32, scale=3.0))
    s = O.sample(p, 20000, N, seed=5)
    cfg = O.all_configs(N)
    pe = np.exp(O.log_probability(p, cfg))
    idx = (s * (1 << np.arange(N - 1, -1, -1))).sum(1)
    cnt = np.bincount(idx, minlength=2 ** N)
    chi2 = ((cnt - 20000 * pe) ** 2 / (20000 * pe)).sum()
    assert chi2 < 2 ** N + 6 * math.sqrt(2 * 2 ** N)
    # sample ids make the stream independent of batch splitting
    s2 = O.sample(p, 100, N, seed=5, sample_offset=50)
    assert np.array_equal(s2[:50], s[50:100])
    pc = O.randomize_biases(O.init_gru_params([6], seed=9, dtype=np.float32, heads=("wf_dense_ampl", "wf_dense_phase"), scale=2.0))
    sc = O.crnn_sample(pc, 500, 8, seed=3)
    assert (sc.sum(1) == 4).all()
    sm = O.mdrnn_sample(O.init_mdrnn_params(5, seed=2), 16, 3, 3, seed=1)
    assert sm.shape == (16, 3, 3) and set(np.unique(sm)) <= {0, 1}


def test_philox_known_answer():
    # Random123 known-answer vectors for philox4x32-10
    out = O.philox4x32(0, 0, 0, 0, 0, 0)
    assert [int(x) for x in out] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    out = O.philox4x32(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff)
    assert [int(x) for x in out] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    out = O.philox4x32(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0)
    assert [int(x) for x in out] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_adam_tf1_first_step():
    th, m, v, t = O.adam_tf1(np.array([1.0]), np.array([0.5]), np.zeros(1), np.zeros(1), 0, 1e-2)
    # first step: lr_t*m/(sqrt(v)+eps) = lr*sqrt(1-b2)/(1-b1) * 0.05/(sqrt(0.00025)+1e-8)
    assert abs(th[0] - (1 - 1e-2 * math.sqrt(0.001) / 0.1 * 0.05 / (math.sqrt(0.00025) + 1e-8))) < 1e-15


def test_autograd_oracle_matches_numpy_forward():
    torch = pytest.importorskip("torch")
    from oracle import torch_grad as TG
    p = O.randomize_biases(O.init_gru_params([5, 5], seed=4, dtype=np.float64, scale=2.0))
    s = np.random.default_rng(0).integers(0, 2, size=(7, 9))
    lp = TG.gru_logprob_t(TG._t(p), [5, 5], s).detach().numpy()
    np.testing.assert_allclose(lp, O.log_probability(p, s), rtol=1e-12)
    pc = O.randomize_biases(O.init_gru_params([6], seed=9, dtype=np.float64, heads=("wf_dense_ampl", "wf_dense_phase"), scale=2.0))
    sc = O.crnn_sample(pc, 9, 8, seed=3)
    re, im = TG.crnn_logamp_t(TG._t(pc), [6], sc)
    la = O.crnn_log_amplitude(pc, sc)
    np.testing.assert_allclose(re.detach().numpy(), la.real, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(im.detach().numpy(), la.imag, rtol=1e-10, atol=1e-12)
    pm = O.init_mdrnn_params(5, seed=2, dtype=np.float64, scale=2.0)
    sm = np.random.default_rng(1).integers(0, 2, size=(6, 3, 4))
    np.testing.assert_allclose(TG.mdrnn_logprob_t(TG._t(pm), sm).detach().numpy(), O.mdrnn_log_probability(pm, sm), rtol=1e-12)
    # finite-difference check of the autograd gradient on one parameter
    w = np.random.default_rng(2).normal(size=7)
    g = TG.gru_vmc_grad(p, s, w)
    flat = O.flatten(p)
    shapes = O.gru_param_shapes([5, 5])
    i = 37
    e = 1e-6
    fp, fm = flat.copy(), flat.copy()
    fp[i] += e
    fm[i] -= e
    fd = ((w * O.log_probability(O.unflatten(fp, shapes), s)).sum() - (w * O.log_probability(O.unflatten(fm, shapes), s)).sum()) / (2 * e)
    assert abs(fd - g[i]) < 1e-6 * max(1, abs(fd))
