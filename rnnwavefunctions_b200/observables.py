"""Observables beyond the energy: the reference's README.md:3 names correlation functions and
entanglement entropies as what RNN wave functions estimate; its code only ships the energy loops.  Everything here reuses
the hot-path kernels through the C ABI (samples from K1, single-flip amplitude ratios from the K2 chain kernel,
log psi from the teacher-forced pass) — there is no CPU path.

    sz_moments(samples)                    <sigma^z_i>, <sigma^z_i sigma^z_j>                (diagonal: sample averages)
    sigma_x(wf, samples)                   <sigma^x_i> = E_sigma[ psi(sigma^(i)) / psi(sigma) ] (rnnwf_tfim_flip_ratios)
    renyi2_entropy(wf, samples, n_A)       S_2(A) = -log E[ psi(a',b) psi(a,b') / (psi(a,b) psi(a',b')) ]   (swap trick)

Spin convention as the reference's local energies: sample value 0/1 -> sigma^z = 2 v - 1 (1DTFIM/TrainingRNN_1DTFIM.py:31-38).
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from ._lib import HEAD_COMPLEX, PARITY_SYM


def _u8(wf, samples):
    """Samples as uint8 [ns, N] on the device.  The parity-symmetric model draws from the plain RNN, P(sigma)
    (1DTFIM/RNNwavefunction_paritysym.py:35-78), while its amplitude is the symmetrised P_sym = (P(sigma) + P(reversed sigma))/2
    (:125-145): estimators under |psi_sym|^2 need sigma ~ P_sym, which reversing every other (i.i.d.) sample provides."""
    su8 = ops.as_u8_samples(samples, wf.device, wf.model.n_sites)
    if wf._flags & PARITY_SYM:
        su8 = su8.clone()
        su8[1::2] = su8[1::2].flip(1)
    return su8


def sz_moments(samples, device=None):
    """samples [ns, N] (0/1, host or device) -> (<sz_i> [N], <sz_i sz_j> [N, N]) as float64 device tensors.
    The second moment is one plain library GEMM of +-1 values in float64 (integer-valued sums: exact)."""
    s = torch.as_tensor(samples)
    if device is not None:
        s = s.to(device)
    if not s.is_cuda:
        raise RuntimeError("observables run on the GPU: pass device samples or a CUDA `device`")
    z = s.reshape(s.shape[0], -1).to(torch.float64) * 2.0 - 1.0
    ns = z.shape[0]
    return z.mean(0), (z.T @ z) / ns


def sz_connected(samples, device=None):
    """Connected correlation C_ij = <sz_i sz_j> - <sz_i><sz_j>."""
    m, c = sz_moments(samples, device)
    return c - torch.outer(m, m)


def sigma_x(wf, samples, return_error=False):
    """<sigma^x_k> for every site k of a TFIM-type (probability-head) wave function: the sample mean of
    psi(sigma with k flipped)/psi(sigma), computed for all k by the prefix-reuse chain kernel.
    -> [N] float64 (and the standard error [N] if `return_error`)."""
    if wf.model.head == HEAD_COMPLEX:
        raise NotImplementedError("sigma_x is defined here for the positive (probability-head) wave functions")
    su8 = _u8(wf, samples)
    N = wf.model.n_sites
    jz = torch.zeros(N, dtype=torch.float64, device=wf.device)
    _, _, ratios = ops.tfim_flip_ratios(wf.model, wf.kernel_params, su8, jz, 1.0, wf._flags)
    mean = ratios.mean(0)
    if return_error:
        return mean, ratios.std(0, unbiased=True) / np.sqrt(ratios.shape[0])
    return mean


def _log_psi(wf, su8):
    """log psi as a complex tensor [ns]: 0.5 log P for the positive models, log amplitude + i phase for the cRNN."""
    out = ops.logpsi(wf.model, wf.kernel_params, su8, wf._flags)
    if wf.model.head == HEAD_COMPLEX:
        return out                                   # complex128: log amplitude + i phase
    return torch.complex(0.5 * out, torch.zeros_like(out))


def renyi2_entropy(wf, samples, n_A, return_error=False):
    """Second Renyi entropy of the first `n_A` sites (in sampling order) by the swap trick on two independent replicas:
    the first and second half of `samples` are paired, the A parts exchanged, and
        exp(-S_2) = E[ psi(a',b) psi(a,b') / (psi(a,b) psi(a',b')) ]
    needs four log psi evaluations per pair (two teacher-forced launches of the log-probability kernel).
    -> S_2 (float), optionally with its delta-method standard error."""
    su8 = _u8(wf, samples)
    ns = su8.shape[0] // 2
    if ns < 1:
        raise ValueError("need at least two samples")
    if not 0 < n_A < wf.model.n_sites:
        raise ValueError("n_A must be inside the system")
    s1, s2 = su8[:ns].reshape(ns, -1), su8[ns:2 * ns].reshape(ns, -1)
    w1 = torch.cat([s2[:, :n_A], s1[:, n_A:]], dim=1).contiguous()      # (a', b)
    w2 = torch.cat([s1[:, :n_A], s2[:, n_A:]], dim=1).contiguous()      # (a, b')
    lp = _log_psi(wf, torch.cat([s1, s2, w1, w2], dim=0).contiguous())
    l1, l2, l3, l4 = lp[:ns], lp[ns:2 * ns], lp[2 * ns:3 * ns], lp[3 * ns:]
    swap = torch.exp(l3 + l4 - l1 - l2).real
    m = swap.mean()
    s2_val = float(-torch.log(m))
    if return_error:
        err = float(swap.std(unbiased=True) / np.sqrt(ns) / m) if ns > 1 else float("nan")
        return s2_val, err
    return s2_val
