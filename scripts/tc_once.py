"""One E_loc call on the tensor-core path at the cfg2 geometry (for ncu captures)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from rnnwavefunctions_b200 import ops, params as P

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 128
N, L, H = 1000, 3, 50
dev = torch.device("cuda:0")
model = ops.make_model(num_layers=L, units=H, n_sites=N)
flat = torch.tensor(P.init_flat(P.gru_shapes([H] * L), 111, np.float32), device=dev)
s = ops.sample(model, flat, ns, seed=1)
e, lp = ops.tfim_eloc(model, flat, s, np.ones(N), 1.0)
torch.cuda.synchronize()
print("mean E", e.mean().item(), "mean lp", lp.mean().item())
