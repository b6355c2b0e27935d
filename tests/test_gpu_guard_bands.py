"""Out-of-bounds writes: compute-sanitizer is closed on the GPU pool this project is developed on (profiles/r2/compute_sanitizer_closed.txt),
so every kernel family runs here with sentinel-filled guard bands on both sides of the caller-provided workspace -- the one buffer
all the hand-rolled indexing (restart-state stash, per-CTA private grids, split-K partials, operand images) lives in -- and with the
workspace sized EXACTLY as rnnwf_workspace_bytes reports.  A kernel that writes outside its carve-outs trips the sentinels or faults."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from rnnwavefunctions_b200 import ops, params as P  # noqa: E402

GUARD = 1 << 16
SENTINEL = 0xA5


class GuardedWorkspace:
    def __init__(self):
        self.allocs = []

    def get(self, nbytes, device):
        raw = torch.full((int(nbytes) + 2 * GUARD,), SENTINEL, dtype=torch.uint8, device=device)
        self.allocs.append((raw, int(nbytes)))
        return raw[GUARD:GUARD + int(nbytes)]

    def check(self):
        torch.cuda.synchronize()
        assert self.allocs, "no workspace was requested"
        for raw, n in self.allocs:
            lo, hi = raw[:GUARD], raw[GUARD + n:]
            assert bool((lo == SENTINEL).all()), "write below the workspace"
            assert bool((hi == SENTINEL).all()), "write past the end of the workspace"
        self.allocs.clear()


@pytest.fixture
def guarded(monkeypatch):
    g = GuardedWorkspace()
    monkeypatch.setattr(ops, "_WS", g)
    yield g


def dev():
    return torch.device("cuda:0")


def gru(units, N, dtype=np.float32, heads=("wf_dense",), nx=0, ny=0):
    model = ops.make_model(head=ops.HEAD_COMPLEX if len(heads) == 2 else ops.HEAD_PROB, dtype=ops.F32 if dtype == np.float32 else ops.F64,
                           num_layers=len(units), units=units[0], n_sites=N, nx=nx, ny=ny)
    flat = torch.tensor(P.init_flat(P.gru_shapes(units, heads=heads), 7, dtype), device=dev())
    return model, flat


@pytest.mark.parametrize("units,N,ns", [([50, 50, 50], 14, 150), ([50], 9, 3), ([40, 40], 12, 129), ([10, 10], 11, 70), ([64], 10, 33)])
def test_f32_gru_every_op_stays_inside_its_workspace(guarded, units, N, ns):
    model, flat = gru(units, N)
    s = ops.sample(model, flat, ns, seed=1)
    guarded.check()
    ops.logpsi(model, flat, s)
    ops.logpsi(model, flat, s, flags=ops.PARITY_SYM)
    guarded.check()
    ops.tfim_eloc(model, flat, s, np.ones(N), 1.0)
    ops.tfim_eloc(model, flat, s, np.ones(N), 1.0, flags=ops.PARITY_SYM)
    ops.tfim_flip_ratios(model, flat, s, np.ones(N), 1.0)
    guarded.check()
    w = torch.randn(ns, dtype=torch.float64, device=dev())
    ops.vmc_grad(model, flat, s, w)
    ops.vmc_grad(model, flat, s, w, flags=ops.PARITY_SYM)
    guarded.check()


@pytest.mark.parametrize("units,N,ns", [([50], 12, 140), ([32, 32], 10, 65), ([12], 8, 20)])
def test_complex_rnn_every_op_stays_inside_its_workspace(guarded, units, N, ns):
    model, flat = gru(units, N, heads=("wf_dense_ampl", "wf_dense_phase"))
    s = ops.sample(model, flat, ns, seed=2)
    ops.logpsi(model, flat, s)
    ops.j1j2_eloc(model, flat, s, np.ones(N), 0.2 * np.ones(N), np.zeros(N), marshall_sign=True)
    ops.vmc_grad(model, flat, s, torch.randn(ns, dtype=torch.complex128, device=dev()))
    guarded.check()


@pytest.mark.parametrize("H,nx,ny,ns", [(100, 12, 12, 70), (12, 3, 4, 129), (50, 4, 5, 64)])
def test_f64_gru_every_op_stays_inside_its_workspace(guarded, H, nx, ny, ns):
    N = nx * ny
    model, flat = gru([H], N, dtype=np.float64, nx=nx, ny=ny)
    s = ops.sample(model, flat, ns, seed=3)
    ops.logpsi(model, flat, s)
    ops.tfim_eloc(model, flat, s, np.ones((nx, ny)), 2.0)           # DMMA chain kernel (one layer, float64)
    ops.vmc_grad(model, flat, s, torch.randn(ns, dtype=torch.float64, device=dev()))
    guarded.check()


@pytest.mark.parametrize("H,nx,ny,ns", [(100, 12, 12, 70), (10, 3, 5, 129), (36, 4, 4, 64)])
def test_mdrnn_every_op_stays_inside_its_workspace(guarded, H, nx, ny, ns):
    model = ops.make_model(cell=ops.CELL_MDRNN, dtype=ops.F64, num_layers=1, units=H, n_sites=nx * ny, nx=nx, ny=ny)
    flat = torch.tensor(P.init_flat(P.mdrnn_shapes(H), 7, np.float64, mdrnn=True) * 0.5, device=dev())
    s = ops.sample(model, flat, ns, seed=4)
    ops.logpsi(model, flat, s)
    ops.tfim_eloc(model, flat, s, np.ones((nx, ny)), 2.0)           # DMMA 2-D RNN chain kernel
    ops.vmc_grad(model, flat, s, torch.randn(ns, dtype=torch.float64, device=dev()))
    guarded.check()
