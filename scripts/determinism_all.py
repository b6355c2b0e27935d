"""Development aid: are sample / E_loc / gradient bitwise reproducible call to call for every BASELINE configuration family?
   python scripts/determinism_all.py [ns] [reps]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from rnnwavefunctions_b200.vmc import J1J2, TFIM, VMC
from rnnwavefunctions_b200.wavefunction import (ComplexRNNwavefunction, RNNwavefunction1D, RNNwavefunction2D, RNNwavefunction2DFlat,
                                                RNNwavefunctionParity)

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
cases = {
    "cfg1 (1 x 50, N = 20)": lambda: (RNNwavefunction1D(20, units=[50]), TFIM(np.ones(20), 1.0)),
    "cfg2-like (3 x 50, N = 120)": lambda: (RNNwavefunction1D(120, units=[50] * 3), TFIM(np.ones(120), 1.0)),
    "cfg2p-like (parity, 3 x 50, N = 96)": lambda: (RNNwavefunctionParity(96, units=[50] * 3), TFIM(np.ones(96), 1.0)),
    "2 x 50, N = 150": lambda: (RNNwavefunction1D(150, units=[50] * 2), TFIM(np.ones(150), 1.0)),
    "cfg3 (12 x 12 flat GRU(100) f64)": lambda: (RNNwavefunction2DFlat(12, 12, units=[100]), TFIM(np.ones((12, 12)), 3.0)),
    "cfg4 (12 x 12 MDRNN(100) f64)": lambda: (RNNwavefunction2D(12, 12, units=[100]), TFIM(np.ones((12, 12)), 3.0)),
    "cfg5 (cRNN 1 x 50, N = 100)": lambda: (ComplexRNNwavefunction(100, units=[50]), J1J2(np.ones(100), 0.2 * np.ones(100), np.zeros(100), True)),
    "FFMA engine (3 x 20, N = 60)": lambda: (RNNwavefunction1D(60, units=[20] * 3), TFIM(np.ones(60), 1.0)),
}
for name, make in cases.items():
    wf, H = make()
    opt = VMC(wf, H, ns)
    ref = None
    bad = [0, 0, 0]
    for r in range(reps):
        opt.draws = 0 if hasattr(opt, "draws") else None
        wf._seed_counter = 0 if hasattr(wf, "_seed_counter") else None
        s = opt.draw() if ref is None else ref[0]
        e = opt.local_energies(s)
        mean, var, n = opt.moments(e)
        g = opt.gradient(s, e, mean, n)
        cur = (s.clone(), e.clone(), g.clone())
        if ref is None:
            ref = cur
            s2 = None
        else:
            bad[1] += int(not torch.equal(cur[1], ref[1]))
            bad[2] += int(not torch.equal(cur[2], ref[2]))
    print(f"{name:40s} ns={ns}: {reps - 1} repeats, E_loc differs {bad[1]}, gradient differs {bad[2]}; mean E {complex(mean.item()).real:.6f}")
