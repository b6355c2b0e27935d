import sys, torch
sys.path.insert(0, ".")
from rnnwavefunctions_b200 import ops
a = torch.zeros(128, 16, device="cuda"); b = torch.zeros(16, 16, device="cuda")
for w in (4, 8, 16):
    ops.umma_selftest(a, b, passes=2, f16=True, dcol=w)
    torch.cuda.synchronize()
for n in (16, 32, 48, 64, 80, 96, 112, 128, 160, 208, 256):
    ops.umma_selftest(a, b, passes=2, f16=True, dcol=100 + n)
    torch.cuda.synchronize()
