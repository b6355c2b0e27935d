"""One sample + E_loc + gradient call of a BASELINE config other than cfg2 (for ncu captures).   python scripts/cfg_once.py cfg3 [ns]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from rnnwavefunctions_b200.vmc import J1J2, TFIM, VMC
from rnnwavefunctions_b200.wavefunction import ComplexRNNwavefunction, RNNwavefunction1D, RNNwavefunction2D, RNNwavefunction2DFlat

cfg = sys.argv[1]
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 1280
if cfg == "cfg1":
    wf, H = RNNwavefunction1D(20, units=[50]), TFIM(np.ones(20), 1.0)
elif cfg == "cfg3":
    wf, H = RNNwavefunction2DFlat(12, 12, units=[100]), TFIM(np.ones((12, 12)), 3.0)
elif cfg == "cfg4":
    wf, H = RNNwavefunction2D(12, 12, units=[100]), TFIM(np.ones((12, 12)), 3.0)
elif cfg == "cfg5":
    wf, H = ComplexRNNwavefunction(100, units=[50]), J1J2(np.ones(100), 0.2 * np.ones(100), np.zeros(100), True)
opt = VMC(wf, H, ns)
s = opt.draw()
e = opt.local_energies(s)
mean, var, n = opt.moments(e)
g = opt.gradient(s, e, mean, n)
torch.cuda.synchronize()
print(cfg, ns, complex(mean.item()).real, g.norm().item())
