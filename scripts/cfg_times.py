"""Stage timings of the other BASELINE configs (parity-test cases, not bench lines): cfg1, cfg3, cfg4, cfg5."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from rnnwavefunctions_b200.vmc import TFIM, J1J2, VMC
from rnnwavefunctions_b200.wavefunction import (ComplexRNNwavefunction, RNNwavefunction1D, RNNwavefunction2D, RNNwavefunction2DFlat)

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): r = fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, r

def run(name, wf, H, ns):
    opt = VMC(wf, H, ns)
    ts, s = timed(lambda: opt.draw())
    te, e = timed(lambda: opt.local_energies(s))
    mean, var, n = opt.moments(e)
    tg, g = timed(lambda: opt.gradient(s, e, mean, n))
    print(f"{name:58s} ns={ns:6d}: sample {ts:8.2f} ms | E_loc {te:9.2f} ms | gradient {tg:8.2f} ms | step {ts+te+tg:9.2f} ms | mean E {complex(mean.item()).real:.4f}")

for ns in (500, 10000):
    run("cfg1 1D TFIM N=20 1xGRU(50) f32", RNNwavefunction1D(20, units=[50]), TFIM(np.ones(20), 1.0), ns)
    run("cfg3 2D TFIM 12x12 1D-RNN GRU(100) f64", RNNwavefunction2DFlat(12, 12, units=[100]), TFIM(np.ones((12, 12)), 3.0), ns)
    run("cfg4 2D TFIM 12x12 2D-RNN MDRNN(100) f64", RNNwavefunction2D(12, 12, units=[100]), TFIM(np.ones((12, 12)), 3.0), ns)
    run("cfg5 J1-J2 N=100 J2=0.2 cRNN 1xGRU(50) f32, Marshall", ComplexRNNwavefunction(100, units=[50]), J1J2(np.ones(100), 0.2 * np.ones(100), np.zeros(100), True), ns)
