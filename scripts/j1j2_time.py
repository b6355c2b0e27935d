"""cfg5 timing: J1-J2 N=100, cRNN 1 x GRU(50): local energies, FFMA vs tensor-core exchange chains."""
import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from rnnwavefunctions_b200 import ops, params as P
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
N, H = 100, 50
dev = torch.device("cuda:0")
model = ops.make_model(head=ops.HEAD_COMPLEX, num_layers=1, units=H, n_sites=N)
flat = torch.tensor(P.init_flat(P.gru_shapes([H], heads=("wf_dense_ampl", "wf_dense_phase")), 111, np.float32), device=dev)
s = ops.sample(model, flat, ns, seed=1)
J1, J2, Bz = np.ones(N), 0.2 * np.ones(N), np.zeros(N)
res = {}
for chain in ("ffma", "tc16"):
    os.environ["RNNWF_CHAIN"] = chain
    for rep in range(2):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        e, la = ops.j1j2_eloc(model, flat, s, J1, J2, Bz, marshall_sign=True)
        b.record()
        torch.cuda.synchronize()
    res[chain] = e
    print(f"j1j2 eloc[{chain}] ns={ns}: {a.elapsed_time(b):.1f} ms  mean E {e.mean().item():.5f}")
print("max |diff|", (res["tc16"] - res["ffma"]).abs().max().item())
