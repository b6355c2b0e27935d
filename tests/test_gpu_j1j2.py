"""GPU parity tests for the complex cRNN / J1-J2 path (J1J2/ComplexRNNwavefunction.py, J1J2/TrainingRNN_J1J2.py):
enumeration bit-exact against the reference's own J1J2MatrixElements (golden vectors), log-amplitudes, local
energies and the complex VMC gradient against the oracle."""
import numpy as np
import pytest
import torch

from oracle import rnnwf_oracle as O

pytestmark = pytest.mark.gpu

from rnnwavefunctions_b200 import ops  # noqa: E402

HEADS = ("wf_dense_ampl", "wf_dense_phase")


def dev():
    return torch.device("cuda:0")


def u8(samples):
    return torch.as_tensor(np.asarray(samples).reshape(len(samples), -1).astype(np.uint8), device=dev())


def crnn_setup(units, N, seed=1, scale=2.0):
    p = O.randomize_biases(O.init_gru_params(units, seed=seed, dtype=np.float32, heads=HEADS, scale=scale), seed=seed + 1)
    model = ops.make_model(cell=ops.CELL_GRU, head=ops.HEAD_COMPLEX, dtype=ops.F32, num_layers=len(units), units=units[0], n_sites=N)
    flat = torch.tensor(O.flatten(p), device=dev())
    assert flat.numel() == ops.param_count(model)
    return p, model, flat


def test_enumeration_bit_exact_vs_reference(golden):
    g = golden("j1j2")
    sig = g["sigmas"]
    N = sig.shape[1]
    for periodic in (False, True):
        for marshall in (False, True):
            tag = f"p{int(periodic)}m{int(marshall)}"
            rows, el, cnt = ops.j1j2_enumerate(u8(sig), g["J1"], g["J2"], g["Bz"], periodic=periodic, marshall_sign=marshall)
            rows, el, cnt = rows.cpu().numpy(), el.cpu().numpy(), cnt.cpu().numpy()
            assert np.array_equal(cnt, g[f"{tag}_num"])
            got_rows = np.concatenate([rows[b, :cnt[b]] for b in range(len(sig))])
            got_el = np.concatenate([el[b, :cnt[b]] for b in range(len(sig))])
            assert np.array_equal(got_rows, g[f"{tag}_sigmaH"])          # integer work: bit-exact
            assert np.array_equal(got_el, g[f"{tag}_elements"])          # float32 matrix elements: bit-exact
    assert N == 8


@pytest.mark.parametrize("units,N,ns", [([7], 10, 64), ([10, 10], 12, 200), ([50], 20, 130)])
def test_log_amplitude_and_sampler(units, N, ns):
    p, model, flat = crnn_setup(units, N)
    s = ops.sample(model, flat, ns, seed=9).cpu().numpy().astype(np.int64)
    assert (s.sum(axis=1) == N // 2).all()                               # U(1): zero magnetisation (:85-93)
    so = O.crnn_sample(p, ns, N, seed=9)
    assert (so == s).all(axis=1).mean() > 0.99
    la = ops.logpsi(model, flat, u8(s)).cpu().numpy()
    ref = O.crnn_log_amplitude(p, s)
    np.testing.assert_allclose(la.real, ref.real, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(la.imag, ref.imag, rtol=1e-5, atol=1e-5)


def test_sector_normalisation():
    N = 8
    p, model, flat = crnn_setup([6], N, scale=3.0)
    cfg = O.all_configs(N)
    cfg = cfg[cfg.sum(axis=1) == N // 2]
    la = ops.logpsi(model, flat, u8(cfg)).cpu().numpy()
    assert abs(np.exp(2 * la.real).sum() - 1) < 1e-5


def test_j1j2_eloc_golden(golden):
    g = golden("j1j2")
    units = [int(u) for u in g["full_units"]]
    samples = g["full_samples"]
    N = samples.shape[1]
    model = ops.make_model(cell=ops.CELL_GRU, head=ops.HEAD_COMPLEX, num_layers=len(units), units=units[0], n_sites=N)
    flat = torch.tensor(g["full_params"].astype(np.float32), device=dev())
    e, la = ops.j1j2_eloc(model, flat, u8(samples), np.ones(N), float(g["full_J2"]) * np.ones(N), np.zeros(N))
    e, la = e.cpu().numpy(), la.cpu().numpy()
    ref = g["full_eloc"]
    assert np.abs(e - ref).max() < 2e-5 * max(1.0, np.abs(ref).max())   # reference combine is complex64
    starts = g["full_starts"]
    np.testing.assert_allclose(la.real, g["full_logamps"][starts[:-1]].real, rtol=2e-5)


@pytest.mark.parametrize("units,N,ns,marshall,j2", [([8], 10, 150, False, 0.2), ([6, 6], 12, 70, True, 0.5), ([5], 8, 40, True, 0.0)])
def test_j1j2_eloc_prefix_reuse_vs_full_recompute(units, N, ns, marshall, j2):
    p, model, flat = crnn_setup(units, N, scale=2.5)
    s = O.crnn_sample(p, ns, N, seed=3)
    rng = np.random.default_rng(4)
    J1, J2, Bz = rng.uniform(0.5, 1.5, N), j2 * np.ones(N), rng.uniform(-0.2, 0.2, N)
    if j2:
        J2[2] = 0.0                                                       # exercises the J2[site] != 0 guard
    p64 = {k: v.astype(np.float64) for k, v in p.items()}
    ref = O.j1j2_local_energies(J1, J2, Bz, s, lambda c: O.crnn_log_amplitude(p, c), marshall_sign=marshall)
    e, _ = ops.j1j2_eloc(model, flat, u8(s), J1, J2, Bz, marshall_sign=marshall)
    e = e.cpu().numpy()
    assert np.abs(e - ref).max() < 3e-5 * max(1.0, np.abs(ref).max())
    del p64


@pytest.mark.parametrize("units,N,ns", [([8], 10, 90), ([6, 6], 8, 60)])
def test_complex_vmc_gradient(units, N, ns):
    from oracle import torch_grad as TG
    p, model, flat = crnn_setup(units, N)
    s = O.crnn_sample(p, ns, N, seed=5)
    rng = np.random.default_rng(3)
    e = rng.normal(size=ns) + 1j * rng.normal(size=ns)
    w = 2.0 * (e - e.mean()) / ns
    ref = TG.crnn_vmc_grad({k: v.astype(np.float64) for k, v in p.items()}, s, w)
    got = ops.vmc_grad(model, flat, u8(s), torch.tensor(w, device=dev())).cpu().numpy()
    err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    assert err < 1e-4, err
