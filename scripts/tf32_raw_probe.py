"""Development aid: what does tcgen05.mma kind::tf32 do with the low 13 mantissa bits of an FP32 operand (A from TMEM, B from shared memory)?"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from rnnwavefunctions_b200 import ops
g = torch.Generator(device="cpu").manual_seed(1)
K, N = 8, 16
a = (torch.rand(128, K, generator=g) + 0.5)
b = (torch.rand(N, K, generator=g) + 0.5)
d = ops.umma_selftest(a.cuda(), b.cuda(), passes=4).cpu().double().numpy()
def trunc(x): return (x.numpy().view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32).astype(np.float64)
def rn(x):
    u = x.numpy().view(np.uint32).astype(np.uint64)
    u = ((u + 0x1000) & 0xFFFFE000).astype(np.uint32)          # round half up in magnitude (positive inputs here)
    return u.view(np.float32).astype(np.float64)
for name, fa, fb in (("truncate both", trunc(a), trunc(b)), ("round both", rn(a), rn(b)), ("trunc A / round B", trunc(a), rn(b)), ("round A / trunc B", rn(a), trunc(b)), ("exact", a.double().numpy(), b.double().numpy())):
    ref = fa @ fb.T
    print(f"{name:20s}: max rel diff {np.abs(d - ref).max() / np.abs(ref).max():.3e}")
