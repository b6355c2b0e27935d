"""Load a TF 1.13 dump written by scripts/tf113_dump.py into the CUDA wave function and compare log-probabilities with the
ones TensorFlow computed (1e-5 relative, BASELINE.json north_star).  Needs a B200.

    python scripts/check_tf_dump.py dump.npz [--parity] [--complex]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("path")
    ap.add_argument("--parity", action="store_true")
    ap.add_argument("--complex", action="store_true")
    a = ap.parse_args()
    from rnnwavefunctions_b200 import wavefunction as W
    z = np.load(a.path)
    samples = z["samples"]
    named = {k: z[k] for k in z.files if k not in ("samples", "log_probs")}
    units = W.units_from_named(named)
    N = samples.shape[1]
    if a.complex:
        wf = W.ComplexRNNwavefunction(N, units=units)
    else:
        wf = (W.RNNwavefunctionParity if a.parity else W.RNNwavefunction1D)(N, units=units)
    wf.set_named_parameters(named)
    got = (wf.log_amplitude(samples) if a.complex else wf.log_probability(samples)).cpu().numpy()
    ref = z["log_probs"]
    err = np.abs(got - ref).max() / np.abs(ref).max()
    print(f"{len(samples)} samples, N = {N}, units {units}: max relative deviation {err:.3e}")
    sys.exit(0 if err < 1e-5 else 1)


if __name__ == "__main__":
    main()
