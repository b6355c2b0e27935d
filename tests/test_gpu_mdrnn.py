"""GPU parity tests for the 2-D RNN wave function (2DTFIM_2DRNN/MDRNNcell.py, RNNwavefunction.py,
Training2DRNN_2DTFIM.py): float64 as in the reference, tolerances 1e-10 relative; enumeration and diagonal
energies bit-exact against the reference-generated golden vectors."""
import math

import numpy as np
import pytest
import torch

from oracle import rnnwf_oracle as O

pytestmark = pytest.mark.gpu

from rnnwavefunctions_b200 import ops  # noqa: E402


def dev():
    return torch.device("cuda:0")


def u8(samples):
    return torch.as_tensor(np.asarray(samples).reshape(len(samples), -1).astype(np.uint8), device=dev())


def md_setup(H, Nx, Ny, dtype=np.float64, seed=1, scale=1.5):
    p = O.randomize_biases(O.init_mdrnn_params(H, seed=seed, dtype=dtype, scale=scale), seed=seed + 1)
    model = ops.make_model(cell=ops.CELL_MDRNN, dtype=ops.F64 if dtype == np.float64 else ops.F32, num_layers=1, units=H,
                           n_sites=Nx * Ny, nx=Nx, ny=Ny)
    flat = torch.tensor(O.flatten(p), device=dev())
    assert flat.numel() == ops.param_count(model)
    return p, model, flat


@pytest.mark.parametrize("H,Nx,Ny,ns", [(7, 4, 4, 50), (10, 3, 5, 33), (100, 12, 12, 40), (9, 5, 2, 17)])
def test_logprob_matches_oracle(H, Nx, Ny, ns):
    p, model, flat = md_setup(H, Nx, Ny, scale=1.5 if H < 50 else 0.5)   # elu recurrences with 100 units blow up beyond ~0.6
    s = np.random.default_rng(0).integers(0, 2, size=(ns, Nx, Ny))
    got = ops.logpsi(model, flat, u8(s)).cpu().numpy()
    np.testing.assert_allclose(got, O.mdrnn_log_probability(p, s), rtol=1e-11)


def test_logprob_f32():
    p, model, flat = md_setup(20, 4, 5, dtype=np.float32)
    s = np.random.default_rng(1).integers(0, 2, size=(90, 4, 5))
    got = ops.logpsi(model, flat, u8(s)).cpu().numpy()
    np.testing.assert_allclose(got, O.mdrnn_log_probability(p, s), rtol=1e-5)


def test_normalisation_and_sampler():
    Nx, Ny = 3, 3
    p, model, flat = md_setup(6, Nx, Ny, scale=1.0)
    cfg = O.all_configs(Nx * Ny).reshape(-1, Nx, Ny)
    lp = ops.logpsi(model, flat, u8(cfg)).cpu().numpy()
    pe = np.exp(lp)
    assert abs(pe.sum() - 1) < 1e-12
    ns = 30000
    s = ops.sample(model, flat, ns, seed=3).cpu().numpy().astype(np.int64)
    so = O.mdrnn_sample(p, 500, Nx, Ny, seed=3)
    assert np.array_equal(s[:500].reshape(500, Nx, Ny), so)          # same Philox convention (counter = path position)
    cnt = np.bincount((s * (1 << np.arange(Nx * Ny - 1, -1, -1))).sum(1), minlength=2 ** (Nx * Ny))
    big = ns * pe >= 5                                                  # pool the rare configurations into one bin
    exp_ = np.append(ns * pe[big], ns * pe[~big].sum())
    obs = np.append(cnt[big], cnt[~big].sum())
    keep = exp_ > 0
    chi2 = ((obs[keep] - exp_[keep]) ** 2 / exp_[keep]).sum()
    dof = keep.sum()
    assert chi2 < dof + 6 * math.sqrt(2 * dof)


def test_tfim2d_mdrnn_golden(golden):
    g = golden("tfim2d")
    H = int(g["md_units"][0])
    model = ops.make_model(cell=ops.CELL_MDRNN, dtype=ops.F64, num_layers=1, units=H, n_sites=16, nx=4, ny=4)
    flat = torch.tensor(g["md_params"], device=dev())
    su8 = u8(g["md_samples"])
    eloc, logp = ops.tfim_eloc(model, flat, su8, g["Jz"], float(g["Bx"]))
    np.testing.assert_allclose(eloc.cpu().numpy(), g["md_eloc"], rtol=1e-11)
    np.testing.assert_allclose(logp.cpu().numpy(), g["md_logprobs"][:5], rtol=1e-11)
    # queue slot i*Ny+j+1 <-> flip of (i,j): the flat enumeration reshaped is the reference's queue (bit-exact)
    q = ops.tfim_enumerate(su8).cpu().numpy().reshape(17, 5, 4, 4)
    assert np.array_equal(q, g["md_queue"])
    assert np.array_equal(ops.tfim_diag(model, su8, g["Jz"]).cpu().numpy(), O.tfim2d_diag(g["Jz"], g["md_samples"]))


@pytest.mark.parametrize("H,Nx,Ny,ns,Bx", [(8, 4, 4, 70, 3.0), (12, 5, 3, 45, 2.0), (100, 6, 6, 10, 2.0)])
def test_eloc_prefix_reuse_equals_full_recompute(H, Nx, Ny, ns, Bx):
    p, model, flat = md_setup(H, Nx, Ny, scale=2.0 if H < 50 else 0.5)
    s = O.mdrnn_sample(p, ns, Nx, Ny, seed=4)
    Jz = np.random.default_rng(5).uniform(0.5, 1.5, size=(Nx, Ny))
    ref = O.ising2d_local_energies(Jz, Bx, Nx, Ny, s, lambda c: O.mdrnn_log_probability(p, c), flat=False)
    eloc, logp = ops.tfim_eloc(model, flat, u8(s), Jz, Bx)
    np.testing.assert_allclose(eloc.cpu().numpy(), ref, rtol=1e-10)
    np.testing.assert_allclose(logp.cpu().numpy(), O.mdrnn_log_probability(p, s), rtol=1e-11)


@pytest.mark.parametrize("H,Nx,Ny,ns", [(6, 3, 3, 40), (10, 4, 5, 77), (100, 4, 4, 30)])
def test_vmc_gradient(H, Nx, Ny, ns):
    from oracle import torch_grad as TG
    p, model, flat = md_setup(H, Nx, Ny, scale=1.5 if H < 50 else 0.5)
    s = O.mdrnn_sample(p, ns, Nx, Ny, seed=6)
    w = np.random.default_rng(0).normal(size=ns)
    ref = TG.mdrnn_vmc_grad(p, s, w)
    got = ops.vmc_grad(model, flat, u8(s), torch.tensor(w, device=dev())).cpu().numpy()
    np.testing.assert_allclose(got, ref, rtol=1e-8, atol=1e-11)


@pytest.mark.parametrize("H,Nx,Ny,ns", [(10, 3, 5, 70), (36, 4, 4, 130), (100, 2, 3, 5), (7, 1, 6, 33), (9, 6, 1, 20), (104, 3, 3, 12), (120, 3, 2, 9)])
def test_eloc_dmma_and_thread_tile_kernels_agree_with_oracle(H, Nx, Ny, ns):
    """2-D RNN local energies on the DMMA chain kernel (mdrnn_f64mma.cuh: up to 13 blocks of 8 units, odd widths and K padding
    included) and on the thread-tile kernel (RNNWF_CHAIN=ffma; widths beyond 104) against the oracle, with ragged sample counts,
    single-row / single-column lattices (no up / no left neighbour anywhere) and row turn-arounds."""
    import os
    p, model, flat = md_setup(H, Nx, Ny, scale=1.5 if H < 50 else 0.5)
    s = O.mdrnn_sample(p, ns, Nx, Ny, seed=4)
    Jz = np.random.default_rng(5).uniform(0.5, 1.5, size=(Nx, Ny))
    ref = O.ising2d_local_energies(Jz, 2.0, Nx, Ny, s, lambda c: O.mdrnn_log_probability(p, c), flat=False)
    eloc, logp = ops.tfim_eloc(model, flat, u8(s), Jz, 2.0)
    np.testing.assert_allclose(eloc.cpu().numpy(), ref, rtol=1e-10)
    np.testing.assert_allclose(logp.cpu().numpy(), O.mdrnn_log_probability(p, s), rtol=1e-11)
    os.environ["RNNWF_CHAIN"] = "ffma"
    try:
        e2, _ = ops.tfim_eloc(model, flat, u8(s), Jz, 2.0)
    finally:
        os.environ.pop("RNNWF_CHAIN", None)
    np.testing.assert_allclose(e2.cpu().numpy(), ref, rtol=1e-10)
