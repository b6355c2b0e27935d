"""Development aid: where the end-to-end (host-API) step of cfg2 spends its host time."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from rnnwavefunctions_b200 import ops
from rnnwavefunctions_b200.training import Ising_local_energies
from rnnwavefunctions_b200.vmc import TFIM, VMC
from rnnwavefunctions_b200.wavefunction import RNNwavefunction1D

N, ns = 1000, 10000
dev = torch.device("cuda:0")
wf = RNNwavefunction1D(N, units=[50] * 3, seed=111, device=dev)
opt = VMC(wf, TFIM(np.ones(N), 1.0), ns)
Jz = np.ones(N)

def T(label, fn, reps=3):
    torch.cuda.synchronize(); fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): r = fn()
    torch.cuda.synchronize()
    print(f"{label:60s} {(time.perf_counter() - t0) / reps * 1e3:8.2f} ms")
    return r

print("torch threads", torch.get_num_threads())
s_dev = T("wf.sample (device, incl. int64 widening)", lambda: wf.sample(ns, 2))
s_np = T("  .cpu().numpy() of int64 [1e4,1e3] (80 MB, pageable)", lambda: s_dev.cpu().numpy())
u8 = T("  device u8 -> pinned host -> int64 numpy", lambda: wf._last_u8.cpu().numpy().astype(np.int64))
T("as_u8_samples(np int64) (host narrow + H2D)", lambda: ops.as_u8_samples(s_np, dev, N))
T("  host narrow only: torch .to(uint8)", lambda: torch.as_tensor(s_np).to(torch.uint8))
T("  host narrow only: numpy astype", lambda: s_np.astype(np.uint8))
e = T("Ising_local_energies(host samples) -> host", lambda: Ising_local_energies(Jz, 1.0, s_np, None, wf, None, None, None), reps=2)
su8 = ops.as_u8_samples(s_np, dev, N)
T("device E_loc only", lambda: opt.local_energies(su8), reps=2)
el = torch.as_tensor(e).to(dev)
T("step_from (moments + gradient + adam) + mean.item()", lambda: float(opt.step_from(su8, el, 5e-3)[0].item()))
