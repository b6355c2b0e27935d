"""Development aid: is the tensor-core base pass (log-probabilities, Bx = 0 => no flip chains) bitwise reproducible?"""
import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from rnnwavefunctions_b200 import ops, params as P
N, L, H = int(os.environ.get("NSITES", "800")), 3, 50
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
dev = torch.device("cuda:0")
model = ops.make_model(num_layers=L, units=H, n_sites=N)
flat = torch.tensor(P.init_flat(P.gru_shapes([H] * L), 111, np.float32), device=dev)
s = ops.sample(model, flat, ns, seed=1)
ref = None
bad = 0
for i in range(reps):
    e, lp = ops.tfim_eloc(model, flat, s, np.ones(N), 0.0)
    if ref is None: ref = lp.clone()
    elif not torch.equal(lp, ref):
        d = (lp != ref).nonzero().flatten()
        bad += 1
        print(f"rep {i}: {d.numel()} rows differ, first rows {d[:8].tolist()}, max abs diff {(lp - ref).abs().max().item():.3e}")
print("base pass:", reps, "repetitions,", bad, "differ")
