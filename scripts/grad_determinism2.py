import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from rnnwavefunctions_b200 import ops, params as P
N, L, H = int(os.environ.get("NSITES", "300")), 3, 50
dev = torch.device("cuda:0")
model = ops.make_model(num_layers=L, units=H, n_sites=N)
flat = torch.tensor(P.init_flat(P.gru_shapes([H] * L), 111, np.float32), device=dev)
for ns in [int(x) for x in sys.argv[1:]]:
    s = ops.sample(model, flat, ns, seed=1)
    w = torch.randn(ns, dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(3)) / ns
    os.environ.pop("RNNWF_BWD_FFMA", None)
    ops.release_workspace()
    gs = [ops.vmc_grad(model, flat, s, w).clone() for _ in range(3)]
    os.environ["RNNWF_BWD_FFMA"] = "1"
    ops.release_workspace()
    gf = ops.vmc_grad(model, flat, s, w).clone()
    rel = lambda a, b: ((a - b).norm() / b.norm()).item()
    print(f"ns={ns:6d}: call1 vs call0 {rel(gs[1], gs[0]):.2e}, call2 vs call1 {rel(gs[2], gs[1]):.2e}, tc vs ffma {rel(gs[2], gf):.2e}")
