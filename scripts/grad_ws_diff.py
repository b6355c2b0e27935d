"""Development aid: which workspace regions differ between two identical gradient calls (race hunting)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from rnnwavefunctions_b200 import ops, params as P
N, L, H = int(os.environ.get("NSITES", "800")), 3, 50
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
dev = torch.device("cuda:0")
model = ops.make_model(num_layers=L, units=H, n_sites=N)
flat = torch.tensor(P.init_flat(P.gru_shapes([H] * L), 111, np.float32), device=dev)
s = ops.sample(model, flat, ns, seed=1)
w = torch.randn(ns, dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(3)) / ns
M = 68 if ns >= 9800 else None
rows = -(-ns // 68) * 68
def a256(x): return (x + 255) & ~255
# carve order (grad.cuh carve_grad / gru.cu carve_gru), sizes in bytes
names = []
off = 0
def take(name, nbytes):
    global off
    off = a256(off); names.append((name, off, nbytes)); off += nbytes
PK = 38502 + 4000
take("pk", 0); take("sigT", rows * N); take("lp_re", rows * 8); take("lp_im", 0); take("counter", 16)
take("hstore", rows * N * L * H * 4); take("la_sel", rows * N * 8); take("la_oth", rows * N * 8); take("la_self", rows * N * 4)
CH = 1 << 26
def sums():
    b = next(iter(ops._WS.buf.values()))
    n = b.numel() // 8 * 8
    v = b[:n].view(torch.int64)
    per = CH // 8
    k = (v.numel() + per - 1) // per
    out = torch.zeros(k, dtype=torch.int64, device=dev)
    for i in range(k):
        out[i] = v[i * per:(i + 1) * per].sum()
    return out
prev = None
gs = []
for c in range(int(os.environ.get("CALLS", "5"))):
    g = ops.vmc_grad(model, flat, s, w); torch.cuda.synchronize()
    gs.append(g.clone())
    cur = sums()
    if prev is not None:
        d = (cur != prev).nonzero().flatten().tolist()
        rel = ((gs[-1] - gs[-2]).norm() / gs[-2].norm()).item()
        print(f"call {c}: grad rel diff vs previous {rel:.2e}; {len(d)} of {len(cur)} 64-MB chunks differ: {d[:40]}")
    prev = cur
b = next(iter(ops._WS.buf.values()))
print("workspace bytes", b.numel(), "chunks", (b.numel() + CH - 1) // CH)
sz = {"hstore": rows * N * L * H * 4, "la": rows * N * 8, "gstore": rows * N * L * 5 * H * 4, "Gbuf": rows * N * 4 * H * 4, "dxbuf": rows * N * H * 4, "dzbuf": rows * N * 2 * 4}
print({k: (v, v // CH) for k, v in sz.items()})
