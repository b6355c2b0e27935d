"""Property tests at the BASELINE geometry (cfg2: N=1000, 3 x GRU(50)), where the oracle is too slow for whole batches:
size-independent identities of the path checked through the C ABI on the GPU."""
import os

import numpy as np
import pytest
import torch

from oracle import rnnwf_oracle as O

pytestmark = pytest.mark.gpu

from rnnwavefunctions_b200 import ops, params as P  # noqa: E402

N, L, H = 1000, 3, 50


@pytest.fixture(scope="module")
def setup():
    dev = torch.device("cuda:0")
    model = ops.make_model(num_layers=L, units=H, n_sites=N)
    flat = torch.tensor(P.init_flat(P.gru_shapes([H] * L), 111, np.float32), device=dev)
    samples = ops.sample(model, flat, 384, seed=5)
    return model, flat, samples


def test_sampler_is_independent_of_sharding(setup):
    model, flat, samples = setup
    a = ops.sample(model, flat, 128, seed=5, sample_offset=0)
    b = ops.sample(model, flat, 256, seed=5, sample_offset=128)
    assert torch.equal(torch.cat([a, b]), samples)                      # global Philox ids: any split gives the same rows
    assert set(torch.unique(samples).tolist()) <= {0, 1}


def test_bx_zero_is_the_diagonal_bit_exact(setup):
    model, flat, samples = setup
    Jz = np.random.default_rng(0).uniform(0.5, 1.5, N)
    e, _ = ops.tfim_eloc(model, flat, samples, Jz, 0.0)
    d = ops.tfim_diag(model, samples, Jz)
    assert torch.equal(e, d)
    ref = O.tfim1d_diag(Jz, samples.cpu().numpy().astype(np.int64))
    assert np.array_equal(d.cpu().numpy(), ref)                         # reference accumulation order, f64


def test_tensor_core_and_cuda_core_chains_agree_at_full_size(setup):
    model, flat, samples = setup
    res = {}
    for chain in ("ffma", "tc16p"):
        os.environ["RNNWF_CHAIN"] = chain
        try:
            e, lp = ops.tfim_eloc(model, flat, samples[:256], np.ones(N), 1.0)
            res[chain] = (e.cpu().numpy(), lp.cpu().numpy())
        finally:
            os.environ.pop("RNNWF_CHAIN", None)
    np.testing.assert_allclose(res["tc16p"][0], res["ffma"][0], rtol=1e-5)
    np.testing.assert_allclose(res["tc16p"][1], res["ffma"][1], rtol=1e-5)
    assert np.all(res["tc16p"][0] > -1272.8762953418 - 300)               # local energies scatter around a variational energy
    assert res["tc16p"][0].mean() > -1272.8762953418                      # above the exact ground state (free fermions / DMRG table)


def test_local_energy_of_three_samples_matches_the_full_recompute_oracle(setup):
    model, flat, samples = setup
    p = O.unflatten(flat.cpu().numpy(), O.gru_param_shapes([H] * L), np.float32)
    s = samples[:3].cpu().numpy().astype(np.int64)
    ref = O.ising_local_energies(np.ones(N), 1.0, s, lambda c: O.log_probability(p, c))
    e, lp = ops.tfim_eloc(model, flat, samples[:3], np.ones(N), 1.0)
    np.testing.assert_allclose(e.cpu().numpy(), ref, rtol=1e-5)
    np.testing.assert_allclose(lp.cpu().numpy(), O.log_probability(p, s), rtol=1e-5)
    np.testing.assert_allclose(ops.logpsi(model, flat, samples[:3]).cpu().numpy(), O.log_probability(p, s), rtol=1e-5)


def test_gradient_is_linear_in_the_weights_and_blind_to_constants(setup):
    model, flat, samples = setup
    s = samples[:200]
    rng = np.random.default_rng(1)
    w1 = torch.tensor(rng.normal(size=200), device=flat.device)
    w2 = torch.tensor(rng.normal(size=200), device=flat.device)
    g1, g2, g12 = (ops.vmc_grad(model, flat, s, w) for w in (w1, w2, w1 + w2))
    err = (g1 + g2 - g12).norm() / g12.norm()
    assert err < 1e-5, float(err)
    # sum_i d log P(sigma_i) over ALL configurations vanishes; over samples it is the score: E[score] -> 0 as 1/sqrt(ns)
    ones = torch.ones(200, dtype=torch.float64, device=flat.device) / 200
    score = ops.vmc_grad(model, flat, s, ones)
    assert score.norm() < 0.5 * g1.norm()
