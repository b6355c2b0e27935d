"""Development aid: is the full E_loc (base pass + flip chains) bitwise reproducible at the cfg2 geometry?   python scripts/eloc_determinism.py [ns] [reps]"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from rnnwavefunctions_b200 import ops, params as P
N, L, H = 1000, 3, 50
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda:0")
model = ops.make_model(num_layers=L, units=H, n_sites=N)
flat = torch.tensor(P.init_flat(P.gru_shapes([H] * L), 111, np.float32), device=dev)
s = ops.sample(model, flat, ns, seed=1)
ref, bad = None, 0
for i in range(reps):
    e, lp = ops.tfim_eloc(model, flat, s, np.ones(N), 1.0)
    if ref is None: ref = (e.clone(), lp.clone())
    else: bad += int(not (torch.equal(e, ref[0]) and torch.equal(lp, ref[1])))
print(f"E_loc at cfg2, {ns} samples: {reps - 1} repeats, {bad} differ; mean E {ref[0].mean().item():.6f}")
