"""Development aid: time of rnnwf_logpsi at cfg2 (10^4 x 1000, 3 x GRU(50)), tensor-core base pass against the CUDA-core forward kernel."""
import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from rnnwavefunctions_b200 import ops, params as P
N, L, H = 1000, 3, 50
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
dev = torch.device("cuda:0")
model = ops.make_model(num_layers=L, units=H, n_sites=N)
flat = torch.tensor(P.init_flat(P.gru_shapes([H] * L), 111, np.float32), device=dev)
s = ops.sample(model, flat, ns, seed=1)
res = {}
for mode in ("tc", "ffma"):
    os.environ["RNNWF_LOGPSI"] = mode
    for _ in range(2): lp = ops.logpsi(model, flat, s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): lp = ops.logpsi(model, flat, s)
    e1.record(); torch.cuda.synchronize()
    res[mode] = lp.clone()
    print(f"{mode}: {e0.elapsed_time(e1) / 5:.2f} ms per call, mean log P {lp.mean().item():.6f}")
print("max rel diff:", ((res["tc"] - res["ffma"]).abs() / res["ffma"].abs()).max().item())
