"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck / initcheck).

    compute-sanitizer --tool memcheck python scripts/sanitize_cases.py tc16p
Cases: tc16p (BASE + flip chains, 3 layers), tc16p1 (1 layer), cplx (J1-J2 exchange chains of the cRNN), grad (stash pass,
backward recurrence, tcgen05 weight-gradient reduction; plain + parity + complex), ffma (CUDA-core engine, FP32 + FP64, sampler),
mdrnn (2-D RNN: sample, E_loc, gradient), misc (enumerations, diag, Adam, moments)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from rnnwavefunctions_b200 import ops, params as P

dev = torch.device("cuda:0")
case = sys.argv[1]


def gru(units, N, dtype=np.float32, heads=("wf_dense",), nx=0, ny=0):
    model = ops.make_model(head=ops.HEAD_COMPLEX if len(heads) == 2 else ops.HEAD_PROB, dtype=ops.F32 if dtype == np.float32 else ops.F64,
                           num_layers=len(units), units=units[0], n_sites=N, nx=nx, ny=ny)
    flat = torch.tensor(P.init_flat(P.gru_shapes(units, heads=heads), 7, dtype), device=dev)
    return model, flat


if case in ("tc16p", "tc16p1"):
    N = 12
    model, flat = gru([50] * (3 if case == "tc16p" else 1), N)
    assert ops.tfim_chain_mode(model) == 3
    s = ops.sample(model, flat, 150, seed=1)
    e, lp = ops.tfim_eloc(model, flat, s, np.ones(N), 1.0)
    e2, lp2 = ops.tfim_eloc(model, flat, s, np.ones(N), 1.0, flags=ops.PARITY_SYM)
    _, _, r = ops.tfim_flip_ratios(model, flat, s, np.ones(N), 1.0)
    print(case, e.mean().item(), e2.mean().item(), r.mean().item())
elif case == "cplx":
    N = 10
    model, flat = gru([50], N, heads=("wf_dense_ampl", "wf_dense_phase"))
    s = ops.sample(model, flat, 140, seed=2)
    e, la = ops.j1j2_eloc(model, flat, s, np.ones(N), 0.2 * np.ones(N), np.zeros(N), marshall_sign=True)
    print(case, e.mean().item())
elif case == "grad":
    N = 10
    model, flat = gru([50, 50], N)
    s = ops.sample(model, flat, 200, seed=3)
    w = torch.randn(200, dtype=torch.float64, device=dev)
    g = ops.vmc_grad(model, flat, s, w)
    gp = ops.vmc_grad(model, flat, s, w, flags=ops.PARITY_SYM)
    modelc, flatc = gru([50], N, heads=("wf_dense_ampl", "wf_dense_phase"))
    sc = ops.sample(modelc, flatc, 130, seed=4)
    gc = ops.vmc_grad(modelc, flatc, sc, torch.randn(130, dtype=torch.complex128, device=dev))
    print(case, g.norm().item(), gp.norm().item(), gc.norm().item())
elif case == "ffma":
    for dtype, nx, ny in ((np.float32, 0, 0), (np.float64, 3, 4)):
        N = 12
        model, flat = gru([10, 10], N, dtype=dtype, nx=nx, ny=ny)
        s = ops.sample(model, flat, 70, seed=5)
        e, lp = ops.tfim_eloc(model, flat, s, np.ones(N), 1.0)
        lp2 = ops.logpsi(model, flat, s)
        g = ops.vmc_grad(model, flat, s, torch.randn(70, dtype=torch.float64, device=dev))
        print(case, dtype.__name__, e.mean().item(), lp2.mean().item(), g.norm().item())
elif case == "mdrnn":
    Nx, Ny, H = 4, 5, 12
    model = ops.make_model(cell=ops.CELL_MDRNN, dtype=ops.F64, num_layers=1, units=H, n_sites=Nx * Ny, nx=Nx, ny=Ny)
    flat = torch.tensor(P.init_flat(P.mdrnn_shapes(H), 7, np.float64, mdrnn=True), device=dev)
    s = ops.sample(model, flat, 90, seed=6)
    e, lp = ops.tfim_eloc(model, flat, s, np.ones((Nx, Ny)), 2.0)
    g = ops.vmc_grad(model, flat, s, torch.randn(90, dtype=torch.float64, device=dev))
    print(case, e.mean().item(), g.norm().item())
elif case == "misc":
    s = torch.randint(0, 2, (33, 10), dtype=torch.uint8, device=dev)
    q = ops.tfim_enumerate(s)
    sig, el, cnt = ops.j1j2_enumerate(s, np.ones(10), 0.2 * np.ones(10), np.zeros(10), periodic=True, marshall_sign=True)
    model = ops.make_model(dtype=ops.F64, units=4, n_sites=10, nx=2, ny=5)
    d = ops.tfim_diag(model, s, np.ones(10))
    th, m, v, g = (torch.randn(1000, dtype=torch.float64, device=dev) for _ in range(4))
    ops.adam_step(model, th, m, v.abs(), g, 1, 1e-3)
    st = ops.energy_moments(d)
    print(case, q.sum().item(), cnt.sum().item(), st.tolist())
torch.cuda.synchronize()
