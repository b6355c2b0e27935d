"""Generate golden vectors from the REFERENCE's own NumPy functions (run in the build container only).

The reference's Hamiltonian / enumeration code is pure NumPy but lives in modules that `import
tensorflow`.  We install a stub `tensorflow` module, import the reference modules from
/root/reference (read-only, never copied), and call:
    1DTFIM/TrainingRNN_1DTFIM.py:13        Ising_local_energies
    2DTFIM_1DRNN/Training1DRNN_2DTFIM.py:13 Ising2D_local_energies (flat samples)
    2DTFIM_2DRNN/Training2DRNN_2DTFIM.py:13 Ising2D_local_energies ([ns,Nx,Ny] samples)
    J1J2/TrainingRNN_J1J2.py:12,95          J1J2MatrixElements, J1J2Slices (+ the combine of :264-279)
with a fake `sess` whose .run() returns log-probabilities / log-amplitudes from the CPU oracle.
Outputs (inputs + reference outputs) are written as small .npz fixtures next to this script; the
fixtures travel to the GPU box, /root/reference does not.

Usage:  python tests/golden/make_golden.py
"""
import importlib
import os
import sys
import types
from math import ceil
from unittest import mock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import rnnwf_oracle as O  # noqa: E402


def load_ref(subdir, module):
    sys.modules["tensorflow"] = mock.MagicMock()
    for m in ("RNNwavefunction", "ComplexRNNwavefunction", "MDRNNcell", module):
        sys.modules.pop(m, None)
    sys.path.insert(0, os.path.join(REF, subdir))
    try:
        return importlib.import_module(module)
    finally:
        sys.path.pop(0)


class FakeSess:
    """sess.run(tensor, feed_dict={placeholder: configs}) -> oracle values."""

    def __init__(self, fn):
        self.fn = fn
        self.calls = 0

    def run(self, tensor, feed_dict):
        (cfg,) = feed_dict.values()
        self.calls += 1
        return self.fn(np.asarray(cfg))


def golden_tfim1d():
    ref = load_ref("1DTFIM", "TrainingRNN_1DTFIM")
    rng = np.random.default_rng(20260101)
    out = {}
    for tag, (ns, N, units, Bx) in {"a": (7, 9, [6], 1.0), "b": (5, 12, [5, 5], 0.7), "c": (4, 6, [4], 0.0)}.items():
        p = O.randomize_biases(O.init_gru_params(units, seed=3 + N, dtype=np.float32, scale=2.0), seed=N)
        samples = rng.integers(0, 2, size=(ns, N)).astype(np.int32)
        Jz = rng.uniform(0.5, 1.5, size=N)
        queue = np.zeros((N + 1, ns, N), np.int32)
        lps = np.zeros((N + 1) * ns, np.float64)
        sess = FakeSess(lambda c: O.log_probability(p, c))
        e = ref.Ising_local_energies(Jz, Bx, samples, queue, "lp_tensor", "ph", lps, sess)
        out.update({f"{tag}_params": O.flatten(p), f"{tag}_units": np.asarray(units), f"{tag}_samples": samples,
                    f"{tag}_Jz": Jz, f"{tag}_Bx": np.float64(Bx), f"{tag}_queue": queue.copy(),
                    f"{tag}_logprobs": lps.copy(), f"{tag}_eloc": e})
        e0 = ref.Ising_local_energies(Jz, 0.0, samples, np.zeros_like(queue), "lp", "ph", np.zeros_like(lps),
                                      FakeSess(lambda c: np.zeros(len(c))))
        out[f"{tag}_diag"] = e0
    # integer-valued diagonal at the benchmark geometry (Jz = ones, N = 1000), few samples
    samples = rng.integers(0, 2, size=(3, 1000)).astype(np.int32)
    Jz = np.ones(1000)
    q = np.zeros((1001, 3, 1000), np.int32)
    e = ref.Ising_local_energies(Jz, 0.0, samples, q, "lp", "ph", np.zeros(1001 * 3), FakeSess(lambda c: np.zeros(len(c))))
    out.update(big_samples=np.packbits(samples.astype(np.uint8), axis=1), big_diag=e)
    np.savez_compressed(os.path.join(HERE, "tfim1d.npz"), **out)


def golden_tfim2d():
    rng = np.random.default_rng(20260102)
    out = {}
    Nx = Ny = 4
    N = Nx * Ny
    ns = 5
    Jz = rng.uniform(0.5, 1.5, size=(Nx, Ny))
    Bx = 3.0
    # 1-D RNN over the flattened lattice (float64 as in 2DTFIM_1DRNN/RNNwavefunction.py:38)
    ref = load_ref("2DTFIM_1DRNN", "Training1DRNN_2DTFIM")
    p = O.randomize_biases(O.init_gru_params([6], seed=21, dtype=np.float64, scale=2.0), seed=2)
    samples = rng.integers(0, 2, size=(ns, N)).astype(np.int32)
    queue = np.zeros((N + 1, ns, N), np.int32)
    lps = np.zeros((N + 1) * ns)
    e = ref.Ising2D_local_energies(Jz, Bx, Nx, Ny, samples, queue, "t", "ph", lps, FakeSess(lambda c: O.log_probability(p, c)))
    out.update(flat_params=O.flatten(p), flat_units=np.asarray([6]), flat_samples=samples, Jz=Jz, Bx=np.float64(Bx),
               flat_queue=queue.copy(), flat_logprobs=lps.copy(), flat_eloc=e)
    # 2-D RNN
    ref = load_ref("2DTFIM_2DRNN", "Training2DRNN_2DTFIM")
    p2 = O.init_mdrnn_params(7, seed=22, dtype=np.float64, scale=1.5)
    samples2 = rng.integers(0, 2, size=(ns, Nx, Ny)).astype(np.int32)
    queue2 = np.zeros((N + 1, ns, Nx, Ny), np.int32)
    lps2 = np.zeros((N + 1) * ns)
    e2 = ref.Ising2D_local_energies(Jz, Bx, Nx, Ny, samples2, queue2, "t", "ph", lps2,
                                    FakeSess(lambda c: O.mdrnn_log_probability(p2, c)))
    out.update(md_params=O.flatten(p2), md_units=np.asarray([7]), md_samples=samples2, md_queue=queue2.copy(),
               md_logprobs=lps2.copy(), md_eloc=e2)
    np.savez_compressed(os.path.join(HERE, "tfim2d.npz"), **out)


def golden_j1j2():
    ref = load_ref("J1J2", "TrainingRNN_J1J2")
    rng = np.random.default_rng(20260103)
    out = {}
    N = 8
    J1 = rng.uniform(0.5, 1.5, size=N)
    J2 = rng.uniform(0.1, 0.6, size=N)
    J2[3] = 0.0                                     # exercises the `J2[site] != 0.0` guard (:52, :84)
    Bz = rng.uniform(-0.3, 0.3, size=N)
    sig = rng.integers(0, 2, size=(6, N)).astype(np.int32)
    out.update(J1=J1, J2=J2, Bz=Bz, sigmas=sig)
    for periodic in (False, True):
        for marshall in (False, True):
            tag = f"p{int(periodic)}m{int(marshall)}"
            rows, els, nums = [], [], []
            for s in sig:
                sigmaH = np.zeros((2 * N + 1, N), np.int32)
                me = np.zeros(2 * N + 1, np.float32)
                num = ref.J1J2MatrixElements(J1, J2, Bz, s, sigmaH, me, periodic=periodic, Marshall_sign=marshall)
                rows.append(sigmaH[:num].copy())
                els.append(me[:num].copy())
                nums.append(num)
            out[f"{tag}_num"] = np.asarray(nums)
            out[f"{tag}_sigmaH"] = np.concatenate(rows)
            out[f"{tag}_elements"] = np.concatenate(els)
    # Full local-energy path at zero magnetisation (J1J2Slices :95-127 + combine :264-279)
    N = 10
    ns = 6
    p = O.randomize_biases(O.init_gru_params([7], seed=31, dtype=np.float32, heads=("wf_dense_ampl", "wf_dense_phase"),
                                             scale=2.0), seed=3)
    samples = O.crnn_sample(p, ns, N, seed=17).astype(np.int32)
    J1 = np.ones(N)
    J2 = 0.2 * np.ones(N)
    Bz = np.zeros(N)
    sigmas = np.zeros((2 * N * ns, N), np.int32)
    H = np.zeros(2 * N * ns, np.float32)
    sigmaH = np.zeros((2 * N, N), np.int32)
    me = np.zeros(2 * N, np.float32)
    slices, length = ref.J1J2Slices(J1, J2, Bz, samples, sigmas, H, sigmaH, me, False)
    la = O.crnn_log_amplitude(p, sigmas[:length]).astype(np.complex64)
    e = np.zeros(ns, np.complex64)
    for n in range(ns):                               # TrainingRNN_J1J2.py:277-279
        s = slices[n]
        e[n] = H[s].dot(np.exp(la[s] - la[s][0]))
    out.update(full_params=O.flatten(p), full_units=np.asarray([7]), full_samples=samples, full_sigmas=sigmas[:length].copy(),
               full_H=H[:length].copy(), full_starts=np.asarray([s.start for s in slices] + [length]),
               full_logamps=la, full_eloc=e, full_J2=np.float64(0.2))
    np.savez_compressed(os.path.join(HERE, "j1j2.npz"), **out)


def golden_known_answers():
    """Exact-diagonalisation helpers exec'd from the tutorial notebooks (pure NumPy cells):
    Tutorials/1DTFIM/Tutorial_1DTFIM.ipynb#cell6 (ED_1DTFIM), Tutorials/J1J2/Tutorial_1DJ1J2.ipynb#cell6
    (ED_j1j2).  Small N only (the notebook code is O(4^N)); the N=10 values are the recorded cell
    outputs (#cell8 of both notebooks)."""
    import json
    out = {}
    nb = json.load(open(os.path.join(REF, "Tutorials/1DTFIM/Tutorial_1DTFIM.ipynb")))
    ns_ = {"np": np}
    exec("".join(nb["cells"][6]["source"]), ns_)
    for N in (4, 6, 8):
        out[f"tfim_N{N}"] = np.float64(np.min(ns_["ED_1DTFIM"](N=N, h=1)[0]))
    out["tfim_N10_recorded"] = np.float64(-12.38148999965476)
    nb = json.load(open(os.path.join(REF, "Tutorials/J1J2/Tutorial_1DJ1J2.ipynb")))
    ns2 = {"np": np}
    exec("".join(nb["cells"][6]["source"]), ns2)
    for N in (4, 6, 8):
        out[f"j1j2_N{N}_J2_0.2"] = np.float64(np.min(ns2["ED_j1j2"](N, j1=1.0, j2=0.2, periodic=False)[0]))
    out["j1j2_N10_J2_0.2_recorded"] = np.float64(-3.9855798336170905)
    np.savez_compressed(os.path.join(HERE, "known_answers.npz"), **out)


if __name__ == "__main__":
    golden_tfim1d()
    golden_tfim2d()
    golden_j1j2()
    golden_known_answers()
    print("wrote", [f for f in os.listdir(HERE) if f.endswith(".npz")])
