"""Device-resident VMC iteration shared by the run_* drivers and bench.py.

One iteration = what the reference's training loops do between two `optstep`s
(1DTFIM/TrainingRNN_1DTFIM.py:199-227, J1J2/TrainingRNN_J1J2.py:241-306):

    samples  = wf.sample(numsamples)                       K1   (rnnwf_sample)
    E_loc    = H.local_energies(wf, samples)               K2   (rnnwf_tfim_eloc / rnnwf_j1j2_eloc)
    mean/var = moments(E_loc)                  [all-reduce of 3-4 doubles over the data-parallel group]
    grad     = sum_i w_i d log psi_i,  w_i = (E_i - mean)/n_total      K3   (rnnwf_vmc_grad)
                                               [all-reduce of the P-vector]
    theta    = TF1-Adam(theta, grad)                             (rnnwf_adam_step)

Data parallelism: samples are sharded, every rank draws `numsamples` rows whose Philox
counters are the *global* sample ids, parameters and Adam moments are replicated and updated identically
on every rank.  The only collectives are the two small sums above (torch.distributed: NCCL on GPUs,
gloo in the CPU tests of this host logic).
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from ._lib import HEAD_COMPLEX


class TFIM:
    """H = -sum Jz s_i s_j - Bx sum sx on an open chain (Jz [N]) or an open Nx x Ny lattice (Jz [Nx,Ny]);
    local energies as Ising_local_energies / Ising2D_local_energies (1DTFIM/TrainingRNN_1DTFIM.py:13-75,
    2DTFIM_*/Training*.py:13-83)."""

    def __init__(self, Jz, Bx):
        self.Jz = np.asarray(Jz, dtype=np.float64)
        self.Bx = float(Bx)
        self._dev = {}

    def _jz(self, device):
        t = self._dev.get(str(device))
        if t is None:
            t = torch.as_tensor(self.Jz.reshape(-1)).to(device)
            self._dev[str(device)] = t
        return t

    def local_energies(self, wf, samples_u8, want_logp=False):
        e, lp = ops.tfim_eloc(wf.model, wf.kernel_params, samples_u8, self._jz(wf.params.device), self.Bx, wf._flags, want_logp=want_logp)
        return (e, lp) if want_logp else e


class J1J2:
    """H = J1 sum S_i.S_{i+1} + J2 sum S_i.S_{i+2}, open chain, optional Marshall sign on the J1 exchange
    (J1J2/TrainingRNN_J1J2.py:12-93); local energies as :255-279."""

    def __init__(self, J1, J2, Bz, marshall_sign=False):
        self.J1, self.J2, self.Bz = (np.asarray(a, dtype=np.float64) for a in (J1, J2, Bz))
        self.marshall_sign = bool(marshall_sign)
        self._dev = {}

    def _arrs(self, device):
        t = self._dev.get(str(device))
        if t is None:
            t = tuple(torch.as_tensor(a).to(device) for a in (self.J1, self.J2, self.Bz))
            self._dev[str(device)] = t
        return t

    def local_energies(self, wf, samples_u8, want_logp=False):
        j1, j2, bz = self._arrs(wf.params.device)
        e, la = ops.j1j2_eloc(wf.model, wf.kernel_params, samples_u8, j1, j2, bz, self.marshall_sign, want_logpsi=want_logp)
        return (e, la) if want_logp else e


def _world(group):
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return None, 0, 1
    return dist, dist.get_rank(group), dist.get_world_size(group)


class VMC:
    """Replicated-parameter, sample-sharded VMC optimiser with the reference's TF1 Adam."""

    def __init__(self, wf, hamiltonian, numsamples, beta1=0.9, beta2=0.999, eps=1e-8, group=None):
        self.wf, self.H, self.ns = wf, hamiltonian, int(numsamples)
        self.beta1, self.beta2, self.eps = beta1, beta2, eps
        self.group = group
        self.dist, self.rank, self.world = _world(group)
        self.mom = torch.zeros_like(wf.params)
        self.vel = torch.zeros_like(wf.params)
        self.t = 0
        wf.sample_offset = self.rank * self.ns
        self.complex = wf.model.head == HEAD_COMPLEX
        self.launches = 0

    # -- stages (each one is a C-ABI call; bench.py times them separately) -----------------------
    def draw(self):
        return ops.sample(self.wf.model, self.wf.kernel_params, self.ns, self.wf._next_seed(), self.wf.sample_offset)

    def local_energies(self, samples_u8):
        return self.H.local_energies(self.wf, samples_u8)

    def moments(self, eloc):
        """-> (mean, var, n_total); mean complex for the cRNN, var of the real part (J1J2/...:281-282)."""
        if self.complex:
            flat = torch.view_as_real(eloc).reshape(-1)
            ns = eloc.numel()
            st = torch.cat([ops.energy_moments(flat, 2, ns), ops.energy_moments(flat[1:], 2, ns)[:1]])
        else:
            st = ops.energy_moments(eloc)
        if self.world > 1:
            self.dist.all_reduce(st, op=self.dist.ReduceOp.SUM, group=self.group)
        n = st[2]
        mean_re = st[0] / n
        var = st[1] / n - mean_re * mean_re
        mean = torch.complex(mean_re, st[3] / n) if self.complex else mean_re
        return mean, var, n

    def gradient(self, samples_u8, eloc, mean, n):
        scale = 2.0 if self.complex else 1.0       # complex cost carries the factor 2 (J1J2/TrainingRNN_J1J2.py:197)
        w = (eloc - mean) * (scale / n)
        g = self.wf.unpad_gradient(ops.vmc_grad(self.wf.model, self.wf.kernel_params, samples_u8, w, self.wf._flags))
        if self.world > 1:
            self.dist.all_reduce(g, op=self.dist.ReduceOp.SUM, group=self.group)
        return g

    def apply(self, grad, lr):
        self.t += 1
        ops.adam_step(self.wf.model, self.wf.params, self.mom, self.vel, grad, self.t, lr, 1.0, self.beta1, self.beta2, self.eps)

    # -- one full iteration ----------------------------------------------------------------------
    def step(self, lr):
        """-> (mean E, var E) as device scalars (no host sync here)."""
        s = self.draw()
        e = self.local_energies(s)
        mean, var, n = self.moments(e)
        g = self.gradient(s, e, mean, n)
        self.apply(g, lr)
        return mean, var

    def step_from(self, samples_u8, eloc, lr):
        """Optimiser half of an iteration from caller-supplied samples and local energies: the analogue of
        sess.run(optstep, feed_dict={Eloc:..., samp:..., learningrate_placeholder: lr}) (:221)."""
        mean, var, n = self.moments(eloc)
        g = self.gradient(samples_u8, eloc, mean, n)
        self.apply(g, lr)
        return mean, var

    # -- checkpoint (replaces tf.train.Saver) --------------------------------------
    def state_dict(self):
        return {"params": self.wf.params.detach().cpu().numpy(), "adam_m": self.mom.cpu().numpy(),
                "adam_v": self.vel.cpu().numpy(), "adam_t": np.int64(self.t), "draws": np.int64(self.wf._draws)}

    def load_state_dict(self, sd):
        dev = self.wf.params.device
        self.wf.params.copy_(torch.as_tensor(sd["params"]).to(dev))
        self.mom.copy_(torch.as_tensor(sd["adam_m"]).to(dev))
        self.vel.copy_(torch.as_tensor(sd["adam_v"]).to(dev))
        self.t = int(sd["adam_t"])
        self.wf._draws = int(sd["draws"])
