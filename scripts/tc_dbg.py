"""One E_loc call at a chosen geometry (development aid; with -DRNNWF_TC16P_DEBUG the kernel prints its phase timers)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from rnnwavefunctions_b200 import ops, params as P

N, L, ns = (int(v) for v in sys.argv[1:4])
H = 50
dev = torch.device("cuda:0")
model = ops.make_model(num_layers=L, units=H, n_sites=N)
flat = torch.tensor(P.init_flat(P.gru_shapes([H] * L), 111, np.float32), device=dev)
s = ops.sample(model, flat, ns, seed=1)
for _ in range(2):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    e, lp = ops.tfim_eloc(model, flat, s, np.ones(N), 1.0)
    b.record()
    torch.cuda.synchronize()
    print(f"N={N} L={L} ns={ns} chain={os.environ.get('RNNWF_CHAIN', 'default')}: {a.elapsed_time(b):.1f} ms, mean E {e.mean().item():.4f}")
