"""Autograd oracle for the VMC gradient (torch float64 on CPU).  TEST INFRASTRUCTURE ONLY.

Restates the differentiable forward passes of the reference in torch so that autograd gives
    grad = sum_s [ w_re[s] d Re log psi_s + w_im[s] d Im log psi_s ]
which is what `optimizer.compute_gradients(cost)` produces for cost = mean(lp*E) - mean(E) mean(lp)
(1DTFIM/TrainingRNN_1DTFIM.py:156-160) with w_s = (E_s - mean E)/ns, and for the complex cost
(J1J2/TrainingRNN_J1J2.py:197) with w = 2 (E_s - mean E)/ns.  Parity unpinned against TF1.13 (no TF here).
"""
import math

import numpy as np
import torch

from . import rnnwf_oracle as O

S = O.SCOPE


def _t(p):
    return {k: torch.tensor(np.asarray(v, np.float64), requires_grad=True) for k, v in p.items()}


def _gru_stack(tp, units, x, hs):
    out = []
    for l, H in enumerate(units):
        b = f"{S}/multi_rnn_cell/cell_{l}/cudnn_compatible_gru_cell/"
        h = hs[l]
        g = torch.sigmoid(torch.cat([x, h], 1) @ tp[b + "gates/kernel"] + tp[b + "gates/bias"])
        r, u = g[:, :H], g[:, H:]
        c = torch.tanh(x @ tp[b + "candidate/input_projection/kernel"] + tp[b + "candidate/input_projection/bias"]
                       + r * (h @ tp[b + "candidate/hidden_projection/kernel"] + tp[b + "candidate/hidden_projection/bias"]))
        h2 = (1 - u) * c + u * h
        out.append(h2)
        x = h2
    return x, out


def _onehot(col):
    return torch.nn.functional.one_hot(torch.as_tensor(col, dtype=torch.long), 2).double()


def gru_logprob_t(tp, units, samples):
    samples = np.asarray(samples)
    B, N = samples.shape
    hs = [torch.zeros(B, h, dtype=torch.float64) for h in units]
    x = torch.zeros(B, 2, dtype=torch.float64)
    lp = torch.zeros(B, dtype=torch.float64)
    idx = torch.as_tensor(samples, dtype=torch.long)
    for n in range(N):
        o, hs = _gru_stack(tp, units, x, hs)
        logp = torch.log_softmax(o @ tp[f"{S}/wf_dense/kernel"] + tp[f"{S}/wf_dense/bias"], 1)
        lp = lp + logp.gather(1, idx[:, n:n + 1])[:, 0]
        x = _onehot(samples[:, n])
    return lp


def crnn_logamp_t(tp, units, samples):
    """returns (Re log psi, Im log psi) following J1J2/ComplexRNNwavefunction.py:138-167."""
    samples = np.asarray(samples)
    B, N = samples.shape
    hs = [torch.zeros(B, h, dtype=torch.float64) for h in units]
    x = torch.zeros(B, 2, dtype=torch.float64)
    re = torch.zeros(B, dtype=torch.float64)
    im = torch.zeros(B, dtype=torch.float64)
    idx = torch.as_tensor(samples, dtype=torch.long)
    for n in range(N):
        o, hs = _gru_stack(tp, units, x, hs)
        amp = torch.sqrt(torch.softmax(o @ tp[f"{S}/wf_dense_ampl/kernel"] + tp[f"{S}/wf_dense_ampl/bias"], 1))
        z = o @ tp[f"{S}/wf_dense_phase/kernel"] + tp[f"{S}/wf_dense_phase/bias"]
        ph = math.pi * z / (1 + z.abs())
        if n >= N / 2:
            n_up = torch.as_tensor(samples[:, :n].sum(1), dtype=torch.float64)
            n_dn = n - n_up
            base = N // 2 - 1
            mask = torch.stack([(base - n_dn >= 0).double(), (base - n_up >= 0).double()], 1)
            amp = amp * mask
            amp = amp / torch.sqrt(torch.clamp((amp * amp).sum(1, keepdim=True), min=1e-30))
        re = re + torch.log(amp.gather(1, idx[:, n:n + 1])[:, 0])
        im = im + ph.gather(1, idx[:, n:n + 1])[:, 0]
        x = _onehot(samples[:, n])
    return re, im


def mdrnn_logprob_t(tp, samples):
    samples = np.asarray(samples)
    B, Nx, Ny = samples.shape
    H = tp[f"{S}/b_rnn_0"].shape[0]
    zero_h = torch.zeros(B, H, dtype=torch.float64)
    zero_x = torch.zeros(B, 2, dtype=torch.float64)
    hg, xg = {}, {}
    lp = torch.zeros(B, dtype=torch.float64)
    idx = torch.as_tensor(samples, dtype=torch.long)
    for (x, y, xn) in O.mdrnn_path(Nx, Ny):
        pre = (xg.get((xn, y), zero_x) @ tp[f"{S}/Uh_rnn_0"] + hg.get((xn, y), zero_h) @ tp[f"{S}/Wh_rnn_0"]
               + xg.get((x, y - 1), zero_x) @ tp[f"{S}/Uv_rnn_0"] + hg.get((x, y - 1), zero_h) @ tp[f"{S}/Wv_rnn_0"]
               + tp[f"{S}/b_rnn_0"])
        h = torch.nn.functional.elu(pre)
        hg[(x, y)] = h
        logp = torch.log_softmax(h @ tp[f"{S}/wf_dense/kernel"] + tp[f"{S}/wf_dense/bias"], 1)
        lp = lp + logp.gather(1, idx[:, x, y].reshape(-1, 1))[:, 0]
        xg[(x, y)] = _onehot(samples[:, x, y])
    return lp


def _flat_grad(tp, scalar):
    scalar.backward()
    return np.concatenate([v.grad.numpy().reshape(-1) for v in tp.values()])


def gru_vmc_grad(p, samples, weights, parity=False):
    tp = _t(p)
    units = O._units_of(p)
    w = torch.as_tensor(np.asarray(weights, np.float64))
    lp = gru_logprob_t(tp, units, samples)
    if parity:
        lp2 = gru_logprob_t(tp, units, np.asarray(samples)[:, ::-1].copy())
        lp = torch.logaddexp(lp, lp2) - math.log(2.0)
    return _flat_grad(tp, (w * lp).sum())


def crnn_vmc_grad(p, samples, weights_complex):
    tp = _t(p)
    units = O._units_of(p)
    w = np.asarray(weights_complex)
    re, im = crnn_logamp_t(tp, units, samples)
    return _flat_grad(tp, (torch.as_tensor(w.real.copy()) * re).sum() + (torch.as_tensor(w.imag.copy()) * im).sum())


def mdrnn_vmc_grad(p, samples, weights):
    tp = _t(p)
    lp = mdrnn_logprob_t(tp, samples)
    return _flat_grad(tp, (torch.as_tensor(np.asarray(weights, np.float64)) * lp).sum())
