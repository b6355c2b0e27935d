"""Development aid: is the gradient bitwise reproducible call to call, and how far are the two backward paths apart?"""
import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from rnnwavefunctions_b200 import ops, params as P
N, L, H = int(sys.argv[2]) if len(sys.argv) > 2 else 1000, 3, 50
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
dev = torch.device("cuda:0")
model = ops.make_model(num_layers=L, units=H, n_sites=N)
flat = torch.tensor(P.init_flat(P.gru_shapes([H] * L), 111, np.float32), device=dev)
s = ops.sample(model, flat, ns, seed=1)
w = torch.randn(ns, dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(3)) / ns
gs = [ops.vmc_grad(model, flat, s, w).clone() for _ in range(4)]
for i in range(1, 4):
    print("tc call", i, "vs call 0: rel diff", ((gs[i] - gs[0]).norm() / gs[0].norm()).item(), "bitwise equal", bool(torch.equal(gs[i], gs[0])))
os.environ["RNNWF_BWD_FFMA"] = "1"
ops.release_workspace()
gf = [ops.vmc_grad(model, flat, s, w).clone() for _ in range(2)]
print("ffma call 1 vs 0 bitwise equal", bool(torch.equal(gf[1], gf[0])))
for i in range(4):
    print("tc call", i, "vs ffma: rel diff", ((gs[i] - gf[0]).norm() / gf[0].norm()).item())
