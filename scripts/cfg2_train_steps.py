"""Evidence run: K full VMC steps at the cfg2 geometry (1-D TFIM N = 1000, 3 x GRU(50), 10^4 samples, lr 5e-3): energy per site and
variance per step (exact ground-state energy per site of the open chain at Bx = 1: -1.2726 for N -> infinity).   python scripts/cfg2_train_steps.py [K]"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from rnnwavefunctions_b200.vmc import TFIM, VMC
from rnnwavefunctions_b200.wavefunction import RNNwavefunction1D
K = int(sys.argv[1]) if len(sys.argv) > 1 else 30
N, ns = 1000, 10000
wf = RNNwavefunction1D(N, units=[50, 50, 50], seed=111)
opt = VMC(wf, TFIM(np.ones(N), 1.0), ns)
torch.cuda.synchronize(); t0 = time.time()
for k in range(K):
    mean, var = opt.step(5e-3)
    if k % 5 == 0 or k == K - 1:
        print(f"step {k:3d}: E/N = {mean.item() / N:.5f}, var(E)/N = {var.item() / N:.4f}", flush=True)
torch.cuda.synchronize()
print(f"{K} steps in {time.time() - t0:.1f} s; parameters finite: {bool(torch.isfinite(wf.params).all())}")
