"""Development aid: time of the sampler at cfg2 (10^4 x 1000, 3 x GRU(50)), tensor-core against CUDA-core kernel."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from rnnwavefunctions_b200 import ops, params as P
N, L, H = int(os.environ.get("NSITES", "1000")), int(os.environ.get("LAYERS", "3")), 50
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
dev = torch.device("cuda:0")
model = ops.make_model(num_layers=L, units=H, n_sites=N)
flat = torch.tensor(P.init_flat(P.gru_shapes([H] * L), 111, np.float32), device=dev)
res = {}
for mode in ("tc", "ffma"):
    os.environ["RNNWF_SAMPLER"] = mode
    for _ in range(2): s = ops.sample(model, flat, ns, seed=1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): s = ops.sample(model, flat, ns, seed=1)
    e1.record(); torch.cuda.synchronize()
    res[mode] = s.clone()
    print(f"{mode}: {e0.elapsed_time(e1) / 5:.2f} ms per call, mean sigma {s.float().mean().item():.4f}")
print("rows equal:", (res["tc"] == res["ffma"]).all(dim=1).float().mean().item())
