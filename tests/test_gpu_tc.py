"""Tensor-core (tcgen05, pipelined 3xFP16 "tc16p") chain kernel vs the CUDA-core chain kernel and the oracle (1e-5 relative gate).
The earlier generations ("tc16": unpipelined 3xFP16, "tc32": 3xTF32) are only compiled with -DRNNWF_LEGACY (A/B builds); their cases
run when such a library is loaded (RNNWF_LIB) and are skipped for the product build."""
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


MODE = {"ffma": 0, "tc32": 1, "tc16": 2, "tc16p": 3}


def has_chain(chain, model):
    """Is this chain-kernel generation compiled into the loaded library?"""
    old = os.environ.get("RNNWF_CHAIN")
    os.environ["RNNWF_CHAIN"] = chain
    try:
        probe = ops.make_model(num_layers=model.num_layers, units=model.units, n_sites=model.n_sites)
        return ops.tfim_chain_mode(probe) == MODE[chain]
    finally:
        if old is None:
            os.environ.pop("RNNWF_CHAIN", None)
        else:
            os.environ["RNNWF_CHAIN"] = old


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
    if not has_chain(chain, model):
        pytest.skip(f"{chain} is a legacy generation (-DRNNWF_LEGACY builds only)")
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
    gens = [c for c in ("tc16p", "tc16") if has_chain(c, model)]
    assert "tc16p" in gens
    for chain in gens + ["ffma"]:
        os.environ["RNNWF_CHAIN"] = chain
        try:
            e, la = ops.j1j2_eloc(model, flat, u8(s), J1, J2, Bz, marshall_sign=marshall)
            out[chain] = (e.cpu().numpy(), la.cpu().numpy())
        finally:
            os.environ.pop("RNNWF_CHAIN", None)
    ref = O.j1j2_local_energies(J1, J2, Bz, s, lambda c: O.crnn_log_amplitude(p, c), marshall_sign=marshall)
    scale = max(1.0, np.abs(ref).max())
    la_ref = O.crnn_log_amplitude(p, s)
    for chain in gens:
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


@pytest.mark.parametrize("H,L,N,ns", [(32, 1, 20, 140), (40, 3, 24, 150), (26, 2, 16, 130)])
def test_narrower_stacks_run_zero_padded_on_the_tensor_core_kernel(H, L, N, ns):
    """26 <= num_units < 50 (`num_units` is a kwarg of every run_* driver, 1DTFIM/TrainingRNN_1DTFIM.py:79): the tcgen05 kernel runs the
    stack zero-padded to 50 units; results must equal the CUDA-core engine's and the oracle's, for the TFIM and the J1-J2 chains."""
    units = [H] * L
    p = O.randomize_biases(O.init_gru_params(units, seed=H, dtype=np.float32, scale=2.0), seed=H + 1)
    model = ops.make_model(num_layers=L, units=H, n_sites=N)
    assert ops.tfim_chain_mode(model) == 3
    assert ops.tfim_chain_mode(ops.make_model(num_layers=L, units=20, n_sites=N)) == 0       # too narrow to pay for 50-unit tensor work
    assert ops.tfim_chain_mode(ops.make_model(num_layers=L, units=64, n_sites=N)) == 0       # wider than the kernel's 50-unit blocks
    flat = torch.tensor(O.flatten(p), device=dev())
    s = O.sample(p, ns, N, seed=3)
    Jz = np.random.default_rng(5).uniform(0.5, 1.5, size=N)
    e_tc, lp_tc = eloc_with("tc16p", model, flat, u8(s), Jz, 0.9)
    e_ff, lp_ff = eloc_with("ffma", model, flat, u8(s), Jz, 0.9)
    ref = O.ising_local_energies(Jz, 0.9, s, lambda c: O.log_probability(p, c))
    np.testing.assert_allclose(e_tc, ref, rtol=1e-5)
    np.testing.assert_allclose(e_tc, e_ff, rtol=2e-5)
    np.testing.assert_allclose(lp_tc, O.log_probability(p, s), rtol=1e-5)
    np.testing.assert_allclose(ops.logpsi(model, flat, u8(s)).cpu().numpy(), O.log_probability(p, s), rtol=1e-5)
    # parity-symmetric model and the gradient (CUDA-core path for this width) are unaffected
    e_par, _ = eloc_with("tc16p", model, flat, u8(s), Jz, 0.9, ops.PARITY_SYM)
    np.testing.assert_allclose(e_par, O.ising_local_energies(Jz, 0.9, s, lambda c: O.log_probability_parity(p, c)), rtol=1e-5)
    from oracle import torch_grad as TG
    w = np.random.default_rng(1).normal(size=ns) / ns
    got = ops.vmc_grad(model, flat, u8(s), torch.tensor(w, device=dev())).cpu().numpy()
    want = TG.gru_vmc_grad({k: v.astype(np.float64) for k, v in p.items()}, s, w)
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-4
    if L == 1 and N % 2 == 0:
        heads = ("wf_dense_ampl", "wf_dense_phase")
        pc = O.randomize_biases(O.init_gru_params(units, seed=H + 2, dtype=np.float32, heads=heads, scale=1.5), seed=H + 3)
        mc = ops.make_model(head=ops.HEAD_COMPLEX, num_layers=L, units=H, n_sites=N)
        fc = torch.tensor(O.flatten(pc), device=dev())
        sc = O.crnn_sample(pc, 100, N, seed=3)
        J1, J2, Bz = np.ones(N), 0.2 * np.ones(N), np.zeros(N)
        e, la = ops.j1j2_eloc(mc, fc, u8(sc), J1, J2, Bz, marshall_sign=True)
        refc = O.j1j2_local_energies(J1, J2, Bz, sc, lambda c: O.crnn_log_amplitude(pc, c), marshall_sign=True)
        assert np.abs(e.cpu().numpy() - refc).max() < 3e-5 * max(1.0, np.abs(refc).max())


@pytest.mark.parametrize("L", [2, 3])
def test_repeated_calls_are_bitwise_identical(L):
    """Every kernel on the path is deterministic (no atomics in the arithmetic, fixed reduction orders), so identical calls must agree
    bit for bit.  This is the regression test of a hand-off race of the pipelined chain kernel: the head partial sums of the second
    row thread of a sample were read behind a barrier that counts arrivals, which a warp running one step ahead could complete early
    (seen as 32 wrong log-probability terms at site N - 2 in about one of ten stash passes of the gradient at 10^4 samples)."""
    N, ns = 96, 10000
    from rnnwavefunctions_b200 import params as P
    model = ops.make_model(num_layers=L, units=50, n_sites=N)
    flat = torch.tensor(P.init_flat(P.gru_shapes([50] * L), 7 + L, np.float32), device=dev())
    s = ops.sample(model, flat, ns, seed=5)
    w = torch.randn(ns, dtype=torch.float64, device=dev(), generator=torch.Generator(device=dev()).manual_seed(3)) / ns
    g0 = ops.vmc_grad(model, flat, s, w).clone()
    e0, lp0 = (t.clone() for t in ops.tfim_eloc(model, flat, s, np.ones(N), 1.0))
    for _ in range(12):
        assert torch.equal(ops.vmc_grad(model, flat, s, w), g0)
    for _ in range(3):
        e, lp = ops.tfim_eloc(model, flat, s, np.ones(N), 1.0)
        assert torch.equal(lp, lp0) and torch.equal(e, e0)


def sample_with(sampler, model, flat, ns, **kw):
    old = os.environ.get("RNNWF_SAMPLER")
    if sampler:
        os.environ["RNNWF_SAMPLER"] = sampler
    else:
        os.environ.pop("RNNWF_SAMPLER", None)
    try:
        return ops.sample(model, flat, ns, **kw).cpu().numpy().astype(np.int64)
    finally:
        if old is None:
            os.environ.pop("RNNWF_SAMPLER", None)
        else:
            os.environ["RNNWF_SAMPLER"] = old


@pytest.mark.parametrize("L,N,ns", [(1, 20, 700), (2, 33, 300), (3, 24, 700), (3, 130, 200)])
def test_tensor_core_sampler_draws_follow_the_oracle_conditionals(L, N, ns):
    """tc16p::chain_kernel<.., SAMPLE> (1DTFIM/RNNwavefunction.py:35-74): every draw is sigma_n = (u_n >= P(0 | sigma_<n)) with the
    Philox uniform of (global sample id, site) unless u is within 1e-5 of the threshold; rows do not depend on the sharding; the
    CUDA-core sampler gives (nearly always) the same rows."""
    units = [50] * L
    p = O.randomize_biases(O.init_gru_params(units, seed=10 + L, dtype=np.float32, scale=2.0), seed=L + 1)
    model = ops.make_model(num_layers=L, units=50, n_sites=N)
    flat = torch.tensor(O.flatten(p), device=dev())
    seed, off = 4321, 777
    s = sample_with(None, model, flat, ns, seed=seed, sample_offset=off)
    assert s.shape == (ns, N) and set(np.unique(s)) <= {0, 1}
    probs = O.gru_conditionals(p, s)
    ids = np.arange(ns, dtype=np.uint64) + np.uint64(off)
    bad = 0
    for n in range(N):
        u = O.philox_uniform(seed, ids, n)
        want = (u >= probs[:, n, 0]).astype(np.int64)
        near = np.abs(u - probs[:, n, 0]) < 1e-5
        bad += int(((want != s[:, n]) & ~near).sum())
    assert bad == 0
    s2 = sample_with(None, model, flat, min(150, ns - 100), seed=seed, sample_offset=off + 100)   # another split, a partly filled tile
    assert np.array_equal(s2, s[100:250])
    sf = sample_with("ffma", model, flat, ns, seed=seed, sample_offset=off)
    assert (sf == s).all(axis=1).mean() > 0.99


def test_batches_beyond_the_workspace_budget_are_evaluated_in_slices():
    """ops.sample_chunk: with a workspace budget smaller than one call needs, E_loc / log psi are evaluated slice by slice (identical
    numbers: samples are independent) and the gradient is summed over slices (the same per-sample terms in another summation order)."""
    N, L, ns = 64, 2, 3000
    from rnnwavefunctions_b200 import params as P
    model = ops.make_model(num_layers=L, units=50, n_sites=N)
    flat = torch.tensor(P.init_flat(P.gru_shapes([50] * L), 3, np.float32), device=dev())
    s = ops.sample(model, flat, ns, seed=9)
    w = torch.randn(ns, dtype=torch.float64, device=dev(), generator=torch.Generator(device=dev()).manual_seed(1)) / ns
    e0, lp0 = ops.tfim_eloc(model, flat, s, np.ones(N), 1.0)
    ep0, _ = ops.tfim_eloc(model, flat, s, np.ones(N), 1.0, ops.PARITY_SYM)
    g0 = ops.vmc_grad(model, flat, s, w).clone()
    l0 = ops.logpsi(model, flat, s).clone()
    need = ops.workspace_bytes(model, ops.OP_VMC_GRAD, ns)
    old = os.environ.get("RNNWF_WS_BUDGET_GB")
    os.environ["RNNWF_WS_BUDGET_GB"] = str(0.1 * need / 2 ** 30)
    try:
        ops.release_workspace()
        assert ops.sample_chunk(model, ops.OP_VMC_GRAD, ns, 0, dev()) < ns
        assert ops.sample_chunk(model, ops.OP_TFIM_ELOC, ns, ops.PARITY_SYM, dev()) < ns
        e1, lp1 = ops.tfim_eloc(model, flat, s, np.ones(N), 1.0)
        ep1, _ = ops.tfim_eloc(model, flat, s, np.ones(N), 1.0, ops.PARITY_SYM)
        g1 = ops.vmc_grad(model, flat, s, w)
        l1 = ops.logpsi(model, flat, s)
    finally:
        if old is None:
            os.environ.pop("RNNWF_WS_BUDGET_GB", None)
        else:
            os.environ["RNNWF_WS_BUDGET_GB"] = old
        ops.release_workspace()
    assert torch.equal(e1, e0) and torch.equal(lp1, lp0) and torch.equal(ep1, ep0) and torch.equal(l1, l0)
    assert ((g1 - g0).norm() / g0.norm()).item() < 1e-5     # the FP32 accumulators of the weight-gradient reduction group the samples by tile


@pytest.mark.parametrize("L,N,ns,parity", [(1, 20, 150, False), (3, 37, 300, True), (2, 64, 200, False), (3, 130, 260, False)])
def test_log_probability_on_the_tensor_core_base_pass(L, N, ns, parity):
    """rnnwf_logpsi for the stacks the tcgen05 kernel covers = its base pass without the stash (1DTFIM/RNNwavefunction.py:76-118,
    RNNwavefunction_paritysym.py:125-145): against the oracle (1e-5) and the CUDA-core forward kernel (RNNWF_LOGPSI=ffma)."""
    units = [50] * L
    p = O.randomize_biases(O.init_gru_params(units, seed=20 + L, dtype=np.float32, scale=2.0), seed=L + 2)
    model = ops.make_model(num_layers=L, units=50, n_sites=N)
    flat = torch.tensor(O.flatten(p), device=dev())
    s = O.sample(p, ns, N, seed=4)
    flags = ops.PARITY_SYM if parity else 0
    got = ops.logpsi(model, flat, u8(s), flags).cpu().numpy()
    os.environ["RNNWF_LOGPSI"] = "ffma"
    try:
        ff = ops.logpsi(model, flat, u8(s), flags).cpu().numpy()
    finally:
        os.environ.pop("RNNWF_LOGPSI", None)
    ref = O.log_probability_parity(p, s) if parity else O.log_probability(p, s)
    np.testing.assert_allclose(got, ref, rtol=1e-5)
    np.testing.assert_allclose(got, ff, rtol=1e-5)
