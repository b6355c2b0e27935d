"""Development aid: when two identical gradient calls differ, which (tile, site, row) entries of la_sel changed?"""
import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from rnnwavefunctions_b200 import ops, params as P
N, L, H, M = int(os.environ.get("NSITES", "800")), 3, 50, 68
ns = 10000
dev = torch.device("cuda:0")
model = ops.make_model(num_layers=L, units=H, n_sites=N)
flat = torch.tensor(P.init_flat(P.gru_shapes([H] * L), 111, np.float32), device=dev)
s = ops.sample(model, flat, ns, seed=1)
w = torch.randn(ns, dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(3)) / ns
tiles = -(-ns // M); rows = tiles * M
off = 0
def take(nbytes):
    global off
    off = (off + 255) & ~255; o = off; off += nbytes; return o
o_pk = take(38504 * 4); o_sig = take(rows * N); o_lp = take(rows * 8); take(0); take(16)
o_h = take(rows * N * L * H * 4); o_ls = take(rows * N * 8); o_lo = take(rows * N * 8)
prev = None
for c in range(int(os.environ.get("CALLS", "12"))):
    g = ops.vmc_grad(model, flat, s, w); torch.cuda.synchronize()
    b = next(iter(ops._WS.buf.values()))
    ls = b[o_ls:o_ls + rows * N * 8].view(torch.float64).view(tiles, N, M).clone()
    lp = b[o_lp:o_lp + rows * 8].view(torch.float64).clone()
    if prev is not None and not torch.equal(ls, prev[0]):
        d = (ls != prev[0]).nonzero()
        print(f"call {c}: {d.shape[0]} la_sel entries differ; tiles {sorted(set(d[:,0].tolist()))[:10]}, sites min {d[:,1].min().item()} max {d[:,1].max().item()} (#distinct {len(set(d[:,1].tolist()))}), rows-in-tile {sorted(set(d[:,2].tolist()))[:40]}")
        rr = sorted(set((d[:,0] * M + d[:,2]).tolist()))
        print("   global rows:", rr[:20], "... count", len(rr), " 128-row work items:", sorted(set(r // 128 for r in rr)))
        k = d[0]
        print("   example:", k.tolist(), prev[0][k[0], k[1], k[2]].item(), "->", ls[k[0], k[1], k[2]].item(), " lp rows differ:", (lp != prev[1]).nonzero().flatten()[:10].tolist())
    prev = (ls, lp)
print("done")
