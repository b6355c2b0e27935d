"""Development aid: E_loc timing + agreement with the FFMA path for the library named by RNNWF_LIB (cfg2 geometry)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from rnnwavefunctions_b200 import ops, params as P

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
N, L, H = 1000, 3, 50
dev = torch.device("cuda:0")
model = ops.make_model(num_layers=L, units=H, n_sites=N)
flat = torch.tensor(P.init_flat(P.gru_shapes([H] * L), 111, np.float32), device=dev)
s = ops.sample(model, flat, ns, seed=1)
Jz = np.ones(N)
ts = []
for _ in range(3):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    e, lp = ops.tfim_eloc(model, flat, s, Jz, 1.0)
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
msg = f"{os.environ.get('RNNWF_LIB', 'default'):40s} ns={ns}: eloc {min(ts):8.1f} ms (runs {', '.join(f'{t:.0f}' for t in ts)})  mean E {e.mean().item():.5f}"
if len(sys.argv) > 2:
    nref = int(sys.argv[2])
    os.environ["RNNWF_CHAIN"] = "ffma"
    er, _ = ops.tfim_eloc(model, flat, s[:nref].contiguous(), Jz, 1.0)
    d = ((e[:nref] - er).abs() / er.abs()).max().item()
    msg += f"  max rel diff vs ffma ({nref} samples) {d:.2e}"
print(msg)
