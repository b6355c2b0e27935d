"""Quick device timing of the cfg2 stages (not the bench contract; used while developing)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from rnnwavefunctions_b200 import ops, params as P

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 1200
N, L, H = 1000, 3, 50
dev = torch.device("cuda:0")
model = ops.make_model(num_layers=L, units=H, n_sites=N)
flat = torch.tensor(P.init_flat(P.gru_shapes([H] * L), 111, np.float32), device=dev)
Jz = np.ones(N)


def timed(fn, reps=1):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        r = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, r


t, s = timed(lambda: ops.sample(model, flat, ns, seed=1))
t, s = timed(lambda: ops.sample(model, flat, ns, seed=1))
print(f"sample  ns={ns}: {t:.1f} ms  mean spin {s.float().mean().item():.3f}")
F = 75800.0
steps = ns * N * (N + 1) / 2
res = {}
for chain in os.environ.get("CHAINS", "ffma,tc16,tc16p").split(","):
    os.environ["RNNWF_CHAIN"] = chain
    t, (e, lp2) = timed(lambda: ops.tfim_eloc(model, flat, s, Jz, 1.0))
    t, (e, lp2) = timed(lambda: ops.tfim_eloc(model, flat, s, Jz, 1.0))
    res[chain] = e
    print(f"eloc[{chain:5s}] ns={ns}: {t:.1f} ms  mean E {e.mean().item():.4f}  -> {steps * F / t / 1e9:.2f} TFLOP/s algorithmic, "
          f"{ns / t * 1e3:.1f} samples/s")
for c in [k for k in res if k != "ffma" and "ffma" in res]:
    d = (res[c] - res["ffma"]).abs() / res["ffma"].abs()
    print(f"{c} vs ffma: max rel diff {d.max().item():.2e}")
w = (e - e.mean()) / ns
t, g = timed(lambda: ops.vmc_grad(model, flat, s, w))
t, g = timed(lambda: ops.vmc_grad(model, flat, s, w))
print(f"grad    ns={ns}: {t:.1f} ms  |g| {g.norm().item():.4e}")
