"""One VMC-gradient call at the cfg2 geometry (for ncu launch lists / captures).   python scripts/grad_once.py [ns]"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from rnnwavefunctions_b200 import ops, params as P
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 1280
N, L, H = 1000, 3, 50
dev = torch.device("cuda:0")
model = ops.make_model(num_layers=L, units=H, n_sites=N)
flat = torch.tensor(P.init_flat(P.gru_shapes([H] * L), 111, np.float32), device=dev)
s = ops.sample(model, flat, ns, seed=1)
w = torch.randn(ns, dtype=torch.float64, device=dev) / ns
for _ in range(2):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g = ops.vmc_grad(model, flat, s, w); b.record(); torch.cuda.synchronize()
    print("grad", ns, a.elapsed_time(b), "ms", g.norm().item())
