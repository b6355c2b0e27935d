"""Tensor-core (tcgen05, 3xTF32) chain kernel vs the CUDA-core chain kernel and the oracle (1e-5 relative gate)."""
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


@pytest.mark.parametrize("chain", ["tc16", "tc32"])
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
