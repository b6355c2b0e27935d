"""Tensor-core (tcgen05; pipelined 3xFP16 "tc16p", 3xFP16 "tc16", 3xTF32 "tc32") chain kernels vs the CUDA-core chain kernel and the oracle (1e-5 relative gate)."""
import os

import numpy as np
import pytest
import torch

from oracle import rnnwf_oracle as O

pytestmark = pytest.mark.gpu

from rnnwavefunctions_b200 import ops  # noqa: E402


def dev():
    return torch.device("cuda:0")


def u8(samples):
    return torch.as_tensor(np.asarray(samples).astype(np.uint8), device=dev())


def eloc_with(chain, model, flat, s, Jz, Bx, flags=0):
    old = os.environ.get("RNNWF_CHAIN")
    os.environ["RNNWF_CHAIN"] = chain
    try:
        e, lp = ops.tfim_eloc(model, flat, s, Jz, Bx, flags=flags)
        return e.cpu().numpy(), lp.cpu().numpy()
    finally:
        if old is None:
            os.environ.pop("RNNWF_CHAIN", None)
        else:
            os.environ["RNNWF_CHAIN"] = old


@pytest.mark.parametrize("chain", ["tc16p", "tc16", "tc32"])
@pytest.mark.parametrize("L,N,ns,parity", [(1, 20, 150, False), (3, 37, 300, False), (2, 24, 70, True), (3, 130, 260, False)])
def test_tc_chain_matches_ffma_and_oracle(L, N, ns, parity, chain):
    units = [50] * L
    p = O.randomize_biases(O.init_gru_params(units, seed=L, dtype=np.float32, scale=2.0), seed=L + 1)
    model = ops.make_model(num_layers=L, units=50, n_sites=N)
    flat = torch.tensor(O.flatten(p), device=dev())
    s = O.sample(p, ns, N, seed=3)
    Jz = np.random.default_rng(5).uniform(0.5, 1.5, size=N)
    flags = ops.PARITY_SYM if parity else 0
    e_tc, lp_tc = eloc_with(chain, model, flat, u8(s), Jz, 0.9, flags)
    e_ff, lp_ff = eloc_with("ffma", model, flat, u8(s), Jz, 0.9, flags)
    np.testing.assert_allclose(e_tc, e_ff, rtol=2e-5)      # two FP32-grade evaluations; each is held to 1e-5 against the oracle below
    np.testing.assert_allclose(lp_tc, lp_ff, rtol=1e-5)
    if N <= 40:
        lpf = (lambda c: O.log_probability_parity(p, c)) if parity else (lambda c: O.log_probability(p, c))
        ref = O.ising_local_energies(Jz, 0.9, s, lpf)
        np.testing.assert_allclose(e_tc, ref, rtol=1e-5)


@pytest.mark.parametrize("L,N,ns,marshall,j2", [(1, 12, 150, False, 0.2), (2, 16, 260, True, 0.5), (1, 10, 40, True, 0.0)])
def test_tc_j1j2_exchange_chains_match_ffma_and_oracle(L, N, ns, marshall, j2):
    heads = ("wf_dense_ampl", "wf_dense_phase")
    units = [50] * L
    p = O.randomize_biases(O.init_gru_params(units, seed=L + 3, dtype=np.float32, heads=heads, scale=1.5), seed=L + 4)
    model = ops.make_model(head=ops.HEAD_COMPLEX, num_layers=L, units=50, n_sites=N)
    flat = torch.tensor(O.flatten(p), device=dev())
    s = O.crnn_sample(p, ns, N, seed=3)
    rng = np.random.default_rng(4)
    J1, J2, Bz = rng.uniform(0.5, 1.5, N), j2 * np.ones(N), rng.uniform(-0.2, 0.2, N)
    if j2:
        J2[2] = 0.0
    out = {}
    for chain in ("tc16p", "tc16", "ffma"):
        os.environ["RNNWF_CHAIN"] = chain
        try:
            e, la = ops.j1j2_eloc(model, flat, u8(s), J1, J2, Bz, marshall_sign=marshall)
            out[chain] = (e.cpu().numpy(), la.cpu().numpy())
        finally:
            os.environ.pop("RNNWF_CHAIN", None)
    ref = O.j1j2_local_energies(J1, J2, Bz, s, lambda c: O.crnn_log_amplitude(p, c), marshall_sign=marshall)
    scale = max(1.0, np.abs(ref).max())
    la_ref = O.crnn_log_amplitude(p, s)
    for chain in ("tc16p", "tc16"):
        assert np.abs(out[chain][0] - out["ffma"][0]).max() < 2e-5 * scale
        assert np.abs(out[chain][0] - ref).max() < 3e-5 * scale              # the reference combine is complex64
        np.testing.assert_allclose(out[chain][1].real, la_ref.real, rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(out[chain][1].imag, la_ref.imag, rtol=1e-5, atol=2e-5)


@pytest.mark.parametrize("ns", [1, 3, 129, 257])
def test_ragged_sample_counts_local_energies_and_gradient(ns):
    """Sample counts that do not fill a tile (1, 3) or spill one row into the next 128-row work item (129, 257): local energies on
    the pipelined tensor-core kernel (3-layer specialised copies) against the full-recompute oracle, and the VMC gradient (stash
    pass on the tensor-core base kernel, weight-gradient reduction on tcgen05 3xTF32) against CPU autograd."""
    from oracle import torch_grad as TG
    L, N = 3, 12
    p = O.randomize_biases(O.init_gru_params([50] * L, seed=7, dtype=np.float32, scale=2.0), seed=8)
    model = ops.make_model(num_layers=L, units=50, n_sites=N)
    flat = torch.tensor(O.flatten(p), device=dev())
    s = O.sample(p, ns, N, seed=11)
    Jz = np.random.default_rng(2).uniform(0.5, 1.5, size=N)
    e, lp = ops.tfim_eloc(model, flat, u8(s), Jz, 1.1)
    ref = O.ising_local_energies(Jz, 1.1, s, lambda c: O.log_probability(p, c))
    np.testing.assert_allclose(e.cpu().numpy(), ref, rtol=1e-5)
    np.testing.assert_allclose(lp.cpu().numpy(), O.log_probability(p, s), rtol=1e-5)
    w = np.random.default_rng(ns).normal(size=ns) / ns
    got = ops.vmc_grad(model, flat, u8(s), torch.tensor(w, device=dev())).cpu().numpy()
    want = TG.gru_vmc_grad({k: v.astype(np.float64) for k, v in p.items()}, s, w)
    err = np.linalg.norm(got - want) / np.linalg.norm(want)
    assert err < 1e-4, err


@pytest.mark.parametrize("L", [1, 2, 3])
@pytest.mark.parametrize("N", [2, 3, 4, 5])
def test_shortest_chains(L, N):
    """Chains of 0..4 sites: the triangles at both ends of the anti-diagonal order have no full diagonal between them."""
    p = O.randomize_biases(O.init_gru_params([50] * L, seed=3 + L, dtype=np.float32, scale=2.0), seed=N)
    model = ops.make_model(num_layers=L, units=50, n_sites=N)
    flat = torch.tensor(O.flatten(p), device=dev())
    s = O.all_configs(N)
    Jz = np.ones(N)
    e, lp = ops.tfim_eloc(model, flat, u8(s), Jz, 0.7)
    ref = O.ising_local_energies(Jz, 0.7, s, lambda c: O.log_probability(p, c))
    np.testing.assert_allclose(e.cpu().numpy(), ref, rtol=1e-5)
    np.testing.assert_allclose(np.exp(lp.cpu().numpy()).sum(), 1.0, rtol=1e-5)      # all 2^N configurations: normalised
