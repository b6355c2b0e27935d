"""GPU parity tests (run with -m gpu on a B200): CUDA path through the C ABI vs the CPU oracle and the
reference-generated golden vectors.  Tolerances: bit-exact for integer/enumeration/diagonal work;
1e-5 relative for FP32 log-probabilities and local energies (BASELINE.json north_star)."""
import math

import numpy as np
import pytest
import torch

from oracle import rnnwf_oracle as O

pytestmark = pytest.mark.gpu

from rnnwavefunctions_b200 import ops  # noqa: E402


def dev():
    return torch.device("cuda:0")


def gru_setup(units, N, dtype=np.float32, heads=("wf_dense",), seed=1, scale=2.0, nx=0, ny=0):
    p = O.randomize_biases(O.init_gru_params(units, seed=seed, dtype=dtype, heads=heads, scale=scale), seed=seed + 1)
    model = ops.make_model(cell=ops.CELL_GRU, head=ops.HEAD_COMPLEX if len(heads) == 2 else ops.HEAD_PROB,
                           dtype=ops.F32 if dtype == np.float32 else ops.F64, num_layers=len(units), units=units[0],
                           n_sites=N, nx=nx, ny=ny)
    flat = torch.tensor(O.flatten(p), device=dev())
    assert flat.numel() == ops.param_count(model)
    return p, model, flat


def u8(samples):
    return torch.as_tensor(np.asarray(samples).reshape(len(samples), -1).astype(np.uint8), device=dev())


@pytest.mark.parametrize("units,N,ns", [([50], 20, 37), ([5, 5], 12, 300), ([7], 9, 11), ([10, 10, 10], 33, 70)])
def test_logprob_matches_oracle_f32(units, N, ns):
    p, model, flat = gru_setup(units, N)
    rng = np.random.default_rng(0)
    s = rng.integers(0, 2, size=(ns, N))
    got = ops.logpsi(model, flat, u8(s)).cpu().numpy()
    ref = O.log_probability(p, s)
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-6)
    got_par = ops.logpsi(model, flat, u8(s), flags=ops.PARITY_SYM).cpu().numpy()
    np.testing.assert_allclose(got_par, O.log_probability_parity(p, s), rtol=1e-5, atol=1e-6)


def test_logprob_f64():
    p, model, flat = gru_setup([6], 16, dtype=np.float64, nx=4, ny=4)
    rng = np.random.default_rng(1)
    s = rng.integers(0, 2, size=(50, 16))
    got = ops.logpsi(model, flat, u8(s)).cpu().numpy()
    np.testing.assert_allclose(got, O.log_probability(p, s), rtol=1e-12)


def test_logprob_benchmark_shape():
    # cfg2 geometry: N=1000, 3 x GRU(50); two tiles of rows
    p, model, flat = gru_setup([50, 50, 50], 1000, scale=1.0, seed=111)
    rng = np.random.default_rng(2)
    s = rng.integers(0, 2, size=(150, 1000))
    got = ops.logpsi(model, flat, u8(s)).cpu().numpy()
    ref = O.log_probability(p, s)
    np.testing.assert_allclose(got, ref, rtol=1e-5)


def test_normalisation_on_device():
    N = 10
    p, model, flat = gru_setup([8, 8], N, scale=3.0)
    cfg = O.all_configs(N)
    lp = ops.logpsi(model, flat, u8(cfg)).cpu().numpy()
    assert abs(np.exp(lp).sum() - 1) < 1e-5


@pytest.mark.parametrize("units,N", [([50], 20), ([6, 6], 14)])
def test_sampler_consistent_with_oracle(units, N):
    p, model, flat = gru_setup(units, N, scale=3.0)
    ns, seed, off = 700, 1234, 1000
    s = ops.sample(model, flat, ns, seed=seed, sample_offset=off).cpu().numpy().astype(np.int64)
    assert s.shape == (ns, N) and set(np.unique(s)) <= {0, 1}
    # teacher-forced check of every draw: sigma_n == (u_n >= p_n[0]) unless u is within 1e-5 of the threshold
    probs = O.gru_conditionals(p, s)
    ids = np.arange(ns, dtype=np.uint64) + np.uint64(off)
    bad = 0
    for n in range(N):
        u = O.philox_uniform(seed, ids, n)
        want = (u >= probs[:, n, 0]).astype(np.int64)
        near = np.abs(u - probs[:, n, 0]) < 1e-5
        bad += int(((want != s[:, n]) & ~near).sum())
    assert bad == 0
    # and the whole-sample agreement with the oracle sampler is (nearly) exact
    so = O.sample(p, ns, N, seed=seed, sample_offset=off)
    assert (so == s).all(axis=1).mean() > 0.995
    # sharding invariance: a different batch split gives the same rows
    s2 = ops.sample(model, flat, 100, seed=seed, sample_offset=off + 50).cpu().numpy()
    assert np.array_equal(s2, s[50:150])


def test_sampler_statistics():
    N = 6
    p, model, flat = gru_setup([5], N, scale=3.0)
    ns = 40000
    s = ops.sample(model, flat, ns, seed=7).cpu().numpy().astype(np.int64)
    pe = np.exp(O.log_probability(p, O.all_configs(N)))
    cnt = np.bincount((s * (1 << np.arange(N - 1, -1, -1))).sum(1), minlength=2 ** N)
    chi2 = ((cnt - ns * pe) ** 2 / (ns * pe)).sum()
    assert chi2 < 2 ** N + 6 * math.sqrt(2 * 2 ** N)


def test_tfim1d_golden(golden):
    g = golden("tfim1d")
    for tag in "abc":
        units = [int(u) for u in g[f"{tag}_units"]]
        samples, Jz, Bx = g[f"{tag}_samples"], g[f"{tag}_Jz"], float(g[f"{tag}_Bx"])
        N = samples.shape[1]
        model = ops.make_model(num_layers=len(units), units=units[0], n_sites=N)
        flat = torch.tensor(g[f"{tag}_params"].astype(np.float32), device=dev())
        su8 = u8(samples)
        diag = ops.tfim_diag(model, su8, Jz).cpu().numpy()
        assert np.array_equal(diag, g[f"{tag}_diag"])                       # bit-exact f64
        if Bx != 0:
            assert np.array_equal(ops.tfim_enumerate(su8).cpu().numpy(), g[f"{tag}_queue"])   # bit-exact
        eloc, logp = ops.tfim_eloc(model, flat, su8, Jz, Bx)
        np.testing.assert_allclose(eloc.cpu().numpy(), g[f"{tag}_eloc"], rtol=1e-5)
        np.testing.assert_allclose(logp.cpu().numpy(), g[f"{tag}_logprobs"][:len(samples)], rtol=1e-5)
    big = np.unpackbits(g["big_samples"], axis=1)[:, :1000]
    model = ops.make_model(num_layers=1, units=4, n_sites=1000)
    assert np.array_equal(ops.tfim_diag(model, u8(big), np.ones(1000)).cpu().numpy(), g["big_diag"])


def test_tfim2d_flat_golden(golden):
    g = golden("tfim2d")
    model = ops.make_model(dtype=ops.F64, num_layers=1, units=int(g["flat_units"][0]), n_sites=16, nx=4, ny=4)
    flat = torch.tensor(g["flat_params"], device=dev())
    su8 = u8(g["flat_samples"])
    eloc, logp = ops.tfim_eloc(model, flat, su8, g["Jz"], float(g["Bx"]))
    np.testing.assert_allclose(eloc.cpu().numpy(), g["flat_eloc"], rtol=1e-11)
    assert np.array_equal(ops.tfim_enumerate(su8).cpu().numpy(), g["flat_queue"])
    # diagonal part alone must be bit-exact (NumPy pairwise summation order reproduced on device)
    p = O.unflatten(g["flat_params"], O.gru_param_shapes([int(g["flat_units"][0])]), np.float64)
    d_ref = O.tfim2d_diag(g["Jz"], g["flat_samples"].reshape(-1, 4, 4))
    assert np.array_equal(ops.tfim_diag(model, su8, g["Jz"]).cpu().numpy(), d_ref)


@pytest.mark.parametrize("nx,ny", [(3, 9), (12, 12), (9, 17)])
def test_tfim2d_diag_bit_exact_pairwise(nx, ny):
    rng = np.random.default_rng(nx * 100 + ny)
    s = rng.integers(0, 2, size=(33, nx, ny))
    Jz = rng.uniform(0.5, 1.5, size=(nx, ny))
    model = ops.make_model(dtype=ops.F64, num_layers=1, units=4, n_sites=nx * ny, nx=nx, ny=ny)
    assert np.array_equal(ops.tfim_diag(model, u8(s), Jz).cpu().numpy(), O.tfim2d_diag(Jz, s))


@pytest.mark.parametrize("units,N,ns,Bx", [([50], 20, 150, 1.0), ([6, 6, 6], 17, 40, 0.6)])
def test_eloc_prefix_reuse_equals_full_recompute(units, N, ns, Bx):
    p, model, flat = gru_setup(units, N, scale=2.5)
    s = O.sample(p, ns, N, seed=3)
    Jz = np.random.default_rng(5).uniform(0.5, 1.5, size=N)
    ref = O.ising_local_energies(Jz, Bx, s, lambda c: O.log_probability(p, c))
    eloc, logp = ops.tfim_eloc(model, flat, u8(s), Jz, Bx)
    np.testing.assert_allclose(eloc.cpu().numpy(), ref, rtol=1e-5)
    refp = O.ising_local_energies(Jz, Bx, s, lambda c: O.log_probability_parity(p, c))
    elocp, logpp = ops.tfim_eloc(model, flat, u8(s), Jz, Bx, flags=ops.PARITY_SYM)
    np.testing.assert_allclose(elocp.cpu().numpy(), refp, rtol=1e-5)
    np.testing.assert_allclose(logpp.cpu().numpy(), O.log_probability_parity(p, s), rtol=1e-5)


def test_eloc_benchmark_shape_small_batch():
    # cfg2 geometry with a handful of samples: prefix reuse (N(N+1)/2 steps) vs the reference algorithm's
    # full recompute of all (N+1) configurations per sample
    N = 1000
    p, model, flat = gru_setup([50, 50, 50], N, scale=1.0, seed=111)
    s = O.sample(p, 3, N, seed=11)
    Jz = np.ones(N)
    ref = O.ising_local_energies(Jz, 1.0, s, lambda c: O.log_probability(p, c))
    eloc, _ = ops.tfim_eloc(model, flat, u8(s), Jz, 1.0)
    np.testing.assert_allclose(eloc.cpu().numpy(), ref, rtol=1e-5)


def test_adam_and_moments():
    rng = np.random.default_rng(0)
    n = 1000
    th, g = rng.normal(size=n), rng.normal(size=n)
    m, v = np.zeros(n), np.zeros(n)
    model = ops.make_model(dtype=ops.F64, units=4, n_sites=4)
    tth, tm, tv = (torch.tensor(a, device=dev()) for a in (th, m, v))
    t = 0
    for step in range(3):
        th, m, v, t = O.adam_tf1(th, g * (step + 1), m, v, t, 1e-2)
        ops.adam_step(model, tth, tm, tv, torch.tensor(g * (step + 1), device=dev()), t, 1e-2)
    np.testing.assert_allclose(tth.cpu().numpy(), th, rtol=1e-12)
    e = torch.tensor(rng.normal(size=777), device=dev())
    st = ops.energy_moments(e).cpu().numpy()
    np.testing.assert_allclose(st, [e.sum().item(), (e * e).sum().item(), 777], rtol=1e-12)


@pytest.mark.parametrize("units,N,ns,parity", [([50], 20, 100, False), ([6, 6, 6], 11, 170, False), ([7, 7], 10, 90, True)])
def test_vmc_gradient_matches_autograd(units, N, ns, parity):
    from oracle import torch_grad as TG
    p, model, flat = gru_setup(units, N, scale=2.0)
    s = O.sample(p, ns, N, seed=5)
    rng = np.random.default_rng(3)
    e = rng.normal(size=ns)
    w = (e - e.mean()) / ns
    p64 = {k: v.astype(np.float64) for k, v in p.items()}
    ref = TG.gru_vmc_grad(p64, s, w, parity=parity)
    got = ops.vmc_grad(model, flat, u8(s), torch.tensor(w, device=dev()), flags=ops.PARITY_SYM if parity else 0).cpu().numpy()
    err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    assert err < 1e-4, err


def test_vmc_gradient_f64_exact():
    from oracle import torch_grad as TG
    p, model, flat = gru_setup([5, 5], 9, dtype=np.float64, nx=3, ny=3)
    s = O.sample(p, 40, 9, seed=2)
    w = np.random.default_rng(0).normal(size=40)
    ref = TG.gru_vmc_grad(p, s, w)
    got = ops.vmc_grad(model, flat, u8(s), torch.tensor(w, device=dev())).cpu().numpy()
    np.testing.assert_allclose(got, ref, rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("H,nx,ny,ns", [(10, 3, 4, 70), (50, 4, 4, 130), (100, 2, 2, 5), (7, 3, 3, 40), (64, 3, 5, 65), (4, 2, 1, 9)])
def test_f64_one_layer_eloc_dmma_and_thread_tile_kernels_agree_with_oracle(H, nx, ny, ns):
    """float64 one-layer GRU local energies: the DMMA chain kernel (gru_f64mma.cuh; even widths with two spare columns in the last block
    of 8 units) and the thread-tile engine (every other width: 7 is odd, 64 fills its blocks) against the oracle's full recompute,
    incl. ragged sample counts (partly filled 64-row tiles), the shortest lattices and a weight ring that does not divide the width."""
    import os
    N = nx * ny
    p, model, flat = gru_setup([H], N, dtype=np.float64, nx=nx, ny=ny, scale=2.0 if H < 60 else 1.0)
    s = O.sample(p, ns, N, seed=3)
    Jz = np.random.default_rng(5).uniform(0.5, 1.5, size=(nx, ny))
    ref = O.ising2d_local_energies(Jz, 1.7, nx, ny, s, lambda c: O.log_probability(p, c), flat=True)
    eloc, logp = ops.tfim_eloc(model, flat, u8(s), Jz, 1.7)
    np.testing.assert_allclose(eloc.cpu().numpy(), ref, rtol=1e-10)
    np.testing.assert_allclose(logp.cpu().numpy(), O.log_probability(p, s), rtol=1e-11)
    os.environ["RNNWF_CHAIN"] = "ffma"
    try:
        e2, _ = ops.tfim_eloc(model, flat, u8(s), Jz, 1.7)
    finally:
        os.environ.pop("RNNWF_CHAIN", None)
    np.testing.assert_allclose(e2.cpu().numpy(), ref, rtol=1e-10)
    # parity symmetry through the same kernels (rows of both directions)
    refp = O.ising2d_local_energies(Jz, 1.7, nx, ny, s, lambda c: O.log_probability_parity(p, c), flat=True)
    ep, _ = ops.tfim_eloc(model, flat, u8(s), Jz, 1.7, flags=ops.PARITY_SYM)
    np.testing.assert_allclose(ep.cpu().numpy(), refp, rtol=1e-10)


@pytest.mark.parametrize("units,dtype", [([50, 30, 40], "f32"), ([6, 9], "f32"), ([12, 7], "f64")])
def test_unequal_layer_widths_run_zero_padded(units, dtype):
    """`units` may be any list (MultiRNNCell([cell(units[n]) ...]), 1DTFIM/RNNwavefunction.py:32): the kernels see the stack zero-padded to
    its widest layer; log-probability, sampler, local energies, gradient and one optimiser step against the oracle at the real widths."""
    from oracle import torch_grad as TG
    from rnnwavefunctions_b200.vmc import TFIM, VMC
    from rnnwavefunctions_b200.wavefunction import RNNwavefunction1D, RNNwavefunction2DFlat
    N, ns = 12, 200
    if dtype == "f32":
        wf, npd, tol = RNNwavefunction1D(N, units=units, seed=5), np.float32, 1e-5
        H = TFIM(np.ones(N), 1.0)
    else:
        wf, npd, tol = RNNwavefunction2DFlat(3, 4, units=units, seed=5), np.float64, 1e-10
        H = TFIM(np.ones((3, 4)), 1.0)
    assert wf.num_params == sum(int(np.prod(s)) for s in O.gru_param_shapes(units).values())
    rng = np.random.default_rng(1)
    wf.params.copy_(torch.tensor(wf.params.cpu().numpy() + rng.uniform(-0.3, 0.3, wf.num_params).astype(npd), device=wf.device))
    p = O.unflatten(wf.params.cpu().numpy(), O.gru_param_shapes(units), npd)
    opt = VMC(wf, H, ns)
    s = opt.draw()
    s_h = s.cpu().numpy().astype(np.int64)
    lp = wf.log_probability(s_h).cpu().numpy()
    np.testing.assert_allclose(lp, O.log_probability(p, s_h), rtol=tol)
    probs = O.gru_conditionals(p, s_h)
    assert 0.2 < s_h.mean() < 0.8 and np.isfinite(probs).all()
    e = opt.local_energies(s)
    if dtype == "f32":
        e_ref = O.ising_local_energies(np.ones(N), 1.0, s_h, lambda c: O.log_probability(p, c))
    else:
        e_ref = O.ising2d_local_energies(np.ones((3, 4)), 1.0, 3, 4, s_h, lambda c: O.log_probability(p, c), flat=True)
    np.testing.assert_allclose(e.cpu().numpy(), e_ref, rtol=max(tol, 1e-9))
    mean, var, n = opt.moments(e)
    g = opt.gradient(s, e, mean, n).cpu().numpy()
    w = (e_ref - e_ref.mean()) / ns
    g_ref = TG.gru_vmc_grad({k: v.astype(np.float64) for k, v in p.items()}, s_h, w)
    assert g.shape == g_ref.shape
    assert np.linalg.norm(g - g_ref) / np.linalg.norm(g_ref) < (1e-4 if dtype == "f32" else 1e-8)
    before = wf.params.clone()
    opt.apply(torch.as_tensor(g, device=wf.device), 1e-2)
    assert float((wf.params - before).abs().max()) > 0 and wf.params.numel() == before.numel()


def test_wave_function_on_a_device_that_is_not_current():
    """Every C-ABI wrapper runs with the device of its parameter tensor current (ops._on_device_of): a wave function built on cuda:1
    works while cuda:0 is the current device and gives the numbers of the same model on cuda:0."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from rnnwavefunctions_b200.vmc import TFIM, VMC
    from rnnwavefunctions_b200.wavefunction import RNNwavefunction1D
    torch.cuda.set_device(0)
    N, ns = 24, 300
    res = []
    for d in ("cuda:0", "cuda:1"):
        wf = RNNwavefunction1D(N, units=[50, 50], seed=3, device=d)
        opt = VMC(wf, TFIM(np.ones(N), 1.0), ns)
        s = opt.draw()
        e = opt.local_energies(s)
        mean, var, n = opt.moments(e)
        g = opt.gradient(s, e, mean, n)
        opt.apply(g, 1e-2)
        assert torch.cuda.current_device() == 0 and s.device == torch.device(d) and g.device == torch.device(d)
        res.append((s.cpu(), e.cpu(), g.cpu(), wf.params.cpu()))
    for a, b in zip(*res):
        assert torch.equal(a, b)


def test_wide_float32_stack_warns_once_and_runs_on_the_cuda_core_engine():
    """float32 stacks wider than 50 units miss the tcgen05 kernels: the class says so (RuntimeWarning) and the FFMA engine gives the
    oracle's numbers."""
    from rnnwavefunctions_b200.wavefunction import RNNwavefunction1D
    N, ns = 10, 64
    with pytest.warns(RuntimeWarning, match="CUDA-core FFMA engine"):
        wf = RNNwavefunction1D(N, units=[64], seed=2)
    assert ops.tfim_chain_mode(wf.model) == 0
    p = O.unflatten(wf.params.cpu().numpy(), O.gru_param_shapes([64]), np.float32)
    s = wf.sample(ns).cpu().numpy()
    np.testing.assert_allclose(wf.log_probability(s).cpu().numpy(), O.log_probability(p, s), rtol=1e-5)
