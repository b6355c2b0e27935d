import os, sys, torch
sys.path.insert(0, ".")
from rnnwavefunctions_b200 import ops
a = torch.zeros(128, 16, device="cuda"); b = torch.zeros(16, 16, device="cuda")
for spec in sys.argv[1:]:
    os.environ["RNNWF_PROBE"] = spec
    ops.umma_selftest(a, b, passes=2, f16=True, dcol=8)
    torch.cuda.synchronize()
