"""Host-side mirror of the reference's training modules: local-energy functions and run_* drivers.

    1DTFIM/TrainingRNN_1DTFIM.py            Ising_local_energies (:13),  run_1DTFIM (:79)
    2DTFIM_1DRNN/Training1DRNN_2DTFIM.py    Ising2D_local_energies (:13), run_2DTFIM (:85)   -> run_2DTFIM_1DRNN
    2DTFIM_2DRNN/Training2DRNN_2DTFIM.py    Ising2D_local_energies (:13), run_2DTFIM (:88)   -> run_2DTFIM_2DRNN
    J1J2/TrainingRNN_J1J2.py                J1J2MatrixElements (:12), J1J2Slices (:95), run_J1J2 (:131)

Same names, argument order and return values.  The TF-specific arguments (`log_probs_tensor`,
`samples_placeholder`, `sess`) are accepted; `log_probs_tensor` carries the wave-function object (what the
TF tensor stood for) and the scratch arrays `queue_samples` / `log_probs` may be None: the fused CUDA path
(rnnwf_tfim_eloc) never materialises the (N+1) x numsamples queue.  Everything below runs on the GPU
through the C ABI; there is no CPU path.
"""
from __future__ import annotations

import os
import random

import numpy as np
import torch

from . import ops
from .vmc import J1J2, TFIM, VMC
from .wavefunction import (ComplexRNNwavefunction, RNNwavefunction1D, RNNwavefunction2D, RNNwavefunction2DFlat,
                           RNNwavefunctionParity, _WavefunctionBase)


def _resolve_wf(log_probs_tensor, sess):
    """The reference passes a TF tensor + session; here the wave-function object is what evaluates log psi."""
    for cand in (log_probs_tensor, sess, getattr(log_probs_tensor, "__self__", None)):
        if isinstance(cand, _WavefunctionBase):
            return cand
    raise TypeError("pass the wave-function object (or its bound .log_probability) as `log_probs_tensor`")


def _to_host(t):
    return t.detach().cpu().numpy()


# ------------------------------------------------------------------------------------------------
# local energies
# ------------------------------------------------------------------------------------------------
def Ising_local_energies(Jz, Bx, samples, queue_samples=None, log_probs_tensor=None, samples_placeholder=None,
                         log_probs=None, sess=None):
    """Local energies of the open 1-D TFIM for `samples` [numsamples, N] (host or device integers)
    -> float64 NumPy array [numsamples]  (1DTFIM/TrainingRNN_1DTFIM.py:13-75).

    One fused launch sequence evaluates the diagonal (bit-exact, :31-38), all N single-flip
    log-probability ratios by prefix reuse and the combine of :74."""
    wf = _resolve_wf(log_probs_tensor, sess)
    su8 = ops.as_u8_samples(samples, wf.device, wf.model.n_sites)
    e = TFIM(Jz, Bx).local_energies(wf, su8)
    if queue_samples is not None:   # the reference leaves the base configurations in slot 0 (:40)
        queue_samples[0] = np.asarray(samples.cpu() if isinstance(samples, torch.Tensor) else samples).reshape(queue_samples[0].shape)
    return _to_host(e)


def Ising2D_local_energies(Jz, Bx, Nx, Ny, samples, queue_samples=None, log_probs_tensor=None, samples_placeholder=None,
                           log_probs=None, sess=None):
    """Local energies of the open 2-D TFIM; `samples` is [numsamples, Nx*Ny] (1-D RNN,
    2DTFIM_1DRNN/Training1DRNN_2DTFIM.py:13-81) or [numsamples, Nx, Ny] (2-D RNN,
    2DTFIM_2DRNN/Training2DRNN_2DTFIM.py:13-83) -> float64 [numsamples]."""
    wf = _resolve_wf(log_probs_tensor, sess)
    if (wf.Nx, wf.Ny) != (Nx, Ny):
        raise ValueError(f"wave function is {wf.Nx}x{wf.Ny}, local energies requested for {Nx}x{Ny}")
    su8 = ops.as_u8_samples(samples, wf.device, wf.model.n_sites)
    e = TFIM(np.asarray(Jz, dtype=np.float64).reshape(Nx, Ny), Bx).local_energies(wf, su8)
    return _to_host(e)


def J1J2MatrixElements(J1, J2, Bz, sigmap, sigmaH, matrixelements, periodic=False, Marshall_sign=False):
    """Connected configurations and matrix elements of one configuration `sigmap` [N]
    (J1J2/TrainingRNN_J1J2.py:12-93): fills sigmaH [>=num, N] and matrixelements [>=num], returns num.
    Row order: diagonal, NN exchanges (ascending site), NNN exchanges (ascending site)."""
    dev = torch.device("cuda", torch.cuda.current_device())
    s = ops.as_u8_samples(np.asarray(sigmap).reshape(1, -1), dev, len(sigmap))
    sig, el, cnt = ops.j1j2_enumerate(s, J1, J2, Bz, periodic=periodic, marshall_sign=Marshall_sign)
    num = int(cnt[0].item())
    sigmaH[:num] = _to_host(sig[0, :num])
    matrixelements[:num] = _to_host(el[0, :num])
    return num


def J1J2Slices(J1, J2, Bz, sigmasp, sigmas, H, sigmaH, matrixelements, Marshall_sign, reference_compat=False):
    """Ragged pack of the connected configurations of every sample (J1J2/TrainingRNN_J1J2.py:95-127)
    -> (slices, total).  The reference passes `Marshall_sign` positionally into `periodic` (:118,
    SURVEY.md B1); `reference_compat=True` reproduces that, the default applies the intended Marshall sign."""
    sigmasp = np.asarray(sigmasp)
    ns, N = sigmasp.shape
    dev = torch.device("cuda", torch.cuda.current_device())
    s = ops.as_u8_samples(sigmasp, dev, N)
    periodic, marshall = (bool(Marshall_sign), False) if reference_compat else (False, bool(Marshall_sign))
    sig, el, cnt = ops.j1j2_enumerate(s, J1, J2, Bz, periodic=periodic, marshall_sign=marshall)
    sig, el, cnt = _to_host(sig), _to_host(el), _to_host(cnt)
    slices, total = [], 0
    for n in range(ns):
        c = int(cnt[n])
        slices.append(slice(total, total + c))
        sigmas[total:total + c] = sig[n, :c]
        H[total:total + c] = el[n, :c]
        total += c
    return slices, total


# ------------------------------------------------------------------------------------------------
# drivers
# ------------------------------------------------------------------------------------------------
def _seed_everything(seed):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def _print_params(wf, verbose):
    """Names, flattened shapes and total count, as 1DTFIM/TrainingRNN_1DTFIM.py:127-136 prints them."""
    total = 0
    for k, v in wf.named_parameters().items():
        if verbose:
            print(k + ":0", (v.size,))
        total += v.size
    if verbose:
        print("The number of params is {0}".format(total))
    return total


def _schedule(lr, lr_schedule):
    """Learning-rate schedules the reference ships: constant (1DTFIM/TrainingRNN_1DTFIM.py:221,
    J1J2/TrainingRNN_J1J2.py:304), 'inverse' = 1/((1/lr) + it/10) (2DTFIM_1DRNN/Training1DRNN_2DTFIM.py:229 and the
    commented alternative at J1J2/TrainingRNN_J1J2.py:301-302), 'inverse5000' = lr (1 + it/5000)^-1
    (2DTFIM_2DRNN/Training2DRNN_2DTFIM.py:228); or any callable it -> lr."""
    if callable(lr_schedule):
        return lr_schedule
    if lr_schedule in (None, "constant"):
        return lambda it: lr
    if lr_schedule == "inverse":
        return lambda it: 1.0 / ((1.0 / lr) + it / 10.0)
    if lr_schedule == "inverse5000":
        return lambda it: lr * (1.0 + it / 5000.0) ** -1
    raise ValueError(f"unknown lr_schedule {lr_schedule!r}")


def _units_ending(units):
    return "_units" + "".join("_{0}".format(u) for u in units)


def _run(wf, hamiltonian, numsteps, numsamples, lr_of_it, numsamples_tag, save_prefix, mean_name, var_name, ckpt_name,
         save, verbose, resume, complex_energy=False):
    """The loop of :199-227: sample -> E_loc -> record -> Adam step -> (checkpoint) -> (save energies)."""
    opt = VMC(wf, hamiltonian, numsamples)
    rank0 = opt.rank == 0
    meanEnergy, varEnergy = [], []
    if save and rank0:
        os.makedirs(save_prefix, exist_ok=True)
    ck = os.path.join(save_prefix, ckpt_name)
    if resume and os.path.exists(ck):
        with np.load(ck) as z:
            opt.load_state_dict({k: z[k] for k in z.files})
            meanEnergy, varEnergy = list(z["meanEnergy"]), list(z["varEnergy"])
    for it in range(len(meanEnergy), numsteps + 1):
        samples = opt.draw()
        eloc = opt.local_energies(samples)
        mean, var, n = opt.moments(eloc)
        meanE = complex(mean.item()) if complex_energy else float(mean.item())
        varE = float(var.item())
        meanEnergy.append(meanE)
        varEnergy.append(varE)
        if it % 10 == 0 and verbose and rank0:
            print("mean(E): {0}, var(E): {1}, #samples {2}, #Step {3} \n\n".format(meanE, varE, numsamples_tag, it))
        g = opt.gradient(samples, eloc, mean, n)
        opt.apply(g, lr_of_it(it))
        if save and rank0 and it % 500 == 0:
            # written AFTER update `it` (the reference's Saver runs before the optstep, :217-221, but it never resumes):
            # the file holds the parameters that iteration it+1 starts from, so a resumed run replays nothing and skips nothing
            sd = opt.state_dict()
            np.savez(ck, meanEnergy=np.asarray(meanEnergy), varEnergy=np.asarray(varEnergy), **sd, **wf.named_parameters())
        if save and rank0 and it % 10 == 0:
            np.save(os.path.join(save_prefix, mean_name), meanEnergy)
            np.save(os.path.join(save_prefix, var_name), varEnergy)
    return meanEnergy, varEnergy


def run_1DTFIM(numsteps=10 ** 4, systemsize=20, num_units=50, Bx=1, num_layers=1, numsamples=500, learningrate=5e-3, seed=111, *,
               parity_symmetric=False, save=True, checkpoint_dir="../Check_Points/1DTFIM", verbose=True, resume=False,
               device=None, lr_schedule=None):
    """VMC of the open 1-D TFIM with a stacked-GRU pRNN (1DTFIM/TrainingRNN_1DTFIM.py:79-229).
    Returns (meanEnergy, varEnergy): lists of length numsteps+1; entry `it` belongs to the parameters before
    update `it`.  `parity_symmetric` selects RNNwavefunction_paritysym (the reference swaps an import, :9-10).
    Under torch.distributed every rank draws `numsamples` samples (weak scaling)."""
    _seed_everything(seed)
    N = systemsize
    Jz = +np.ones(N)
    lr = np.float64(learningrate)
    units = [num_units] * num_layers
    cls = RNNwavefunctionParity if parity_symmetric else RNNwavefunction1D
    wf = cls(N, units=units, seed=seed, device=device)
    _print_params(wf, verbose)
    ending = _units_ending(units)
    tag = "_N" + str(N) + "_samp" + str(numsamples) + "_Jz" + str(Jz[0]) + "_Bx" + str(Bx) + "_GRURNN_OBC" + "_TFIM" + ending
    return _run(wf, TFIM(Jz, Bx), numsteps, numsamples, _schedule(lr, lr_schedule), numsamples, checkpoint_dir,
                "meanEnergy" + tag + ".npy", "varEnergy" + tag + ".npy",
                "RNNwavefunction_N" + str(N) + "_samp" + str(numsamples) + "_Jz1Bx" + str(Bx) + "_GRURNN_OBC" + ending + ".npz",
                save, verbose, resume)


def run_2DTFIM_1DRNN(numsteps=2 * 10 ** 4, systemsize_x=5, systemsize_y=5, Bx=+2, num_units=50, num_layers=1, numsamples=500,
                     learningrate=1e-3, seed=333, *, save=True, checkpoint_dir="../Check_Points/2DTFIM", verbose=True,
                     resume=False, device=None, lr_schedule="inverse"):
    """2-D TFIM with a 1-D GRU pRNN over the flattened lattice, float64
    (2DTFIM_1DRNN/Training1DRNN_2DTFIM.py:85-233); lr schedule 1/((1/lr)+it/10) (:229)."""
    _seed_everything(seed)
    Nx, Ny = systemsize_x, systemsize_y
    Jz = +np.ones((Nx, Ny))
    lr = np.float64(learningrate)
    units = [num_units] * num_layers
    wf = RNNwavefunction2DFlat(Nx, Ny, units=units, seed=seed, device=device)   # the reference forgets to forward seed (B6)
    _print_params(wf, verbose)
    ending = _units_ending(units)[1:]    # the 2-D apps spell it 'units_50' (:137-139)
    tag = "_" + str(Nx) + "x" + str(Ny) + "_Bx" + str(Bx) + "_lradap" + str(lr) + "_samp" + str(numsamples) + ending
    return _run(wf, TFIM(Jz, Bx), numsteps, numsamples, _schedule(lr, lr_schedule), numsamples, checkpoint_dir,
                "meanEnergy_GRURNN" + tag + "_2DTFIM.npy", "varEnergy_GRURNN" + tag + "_2DTFIM.npy",
                "RNNwavefunction_GRURNN" + tag + ".npz", save, verbose, resume)


def run_2DTFIM_2DRNN(numsteps=2 * 10 ** 4, systemsize_x=5, systemsize_y=5, Bx=+2, num_units=50, numsamples=500,
                     learningrate=5e-3, seed=111, *, save=True, checkpoint_dir="../Check_Points/2DTFIM", verbose=True,
                     resume=False, device=None, lr_schedule="inverse5000"):
    """2-D TFIM with the 2-D RNN (MDRNNcell on the zig-zag path), float64
    (2DTFIM_2DRNN/Training2DRNN_2DTFIM.py:88-231); lr schedule lr (1+it/5000)^-1 (:228).
    (As shipped the reference raises UnboundLocalError at :99; this is the intended function.)"""
    _seed_everything(seed)
    Nx, Ny = systemsize_x, systemsize_y
    Jz = +np.ones((Nx, Ny))
    lr = np.float64(learningrate)
    units = [num_units]
    wf = RNNwavefunction2D(Nx, Ny, units=units, seed=seed, device=device)
    _print_params(wf, verbose)
    ending = _units_ending(units)[1:]
    tag = "_" + str(Nx) + "x" + str(Ny) + "_Bx" + str(Bx) + "_lradap" + str(lr) + "_samp" + str(numsamples) + ending
    return _run(wf, TFIM(Jz, Bx), numsteps, numsamples, _schedule(lr, lr_schedule), numsamples, checkpoint_dir,
                "meanEnergy_2DVanillaRNN" + tag + "_2DTFIM.npy", "varEnergy_2DVanillaRNN" + tag + "_2DTFIM.npy",
                "RNNwavefunction_2DVanillaRNN" + tag + ".npz", save, verbose, resume)


def run_J1J2(numsteps=10 ** 5, systemsize=20, J1_=1.0, J2_=0.0, Marshall_sign=False, num_units=50, num_layers=1, numsamples=500,
             learningrate=2.5e-4, seed=111, *, reference_compat=False, save=True, checkpoint_dir="../Check_Points/J1J2",
             verbose=True, resume=False, device=None, lr_schedule=None):
    """VMC of the open J1-J2 chain with the complex cRNN in the zero-magnetisation sector
    (J1J2/TrainingRNN_J1J2.py:131-308).  Returns (meanEnergy [complex], varEnergy [variance of the real part]).
    `Marshall_sign` applies the intended Marshall rotation; `reference_compat=True` reproduces the reference's
    mis-binding of the flag to `periodic`, which the fused open-chain kernel does not support."""
    if reference_compat and Marshall_sign:
        raise NotImplementedError("reference_compat reproduces periodic wrap-around bonds (SURVEY.md B1); use J1J2Slices(..., "
                                  "reference_compat=True) with log_amplitude for that path")
    _seed_everything(seed)
    N = systemsize
    J1, J2, Bz = +J1_ * np.ones(N), +J2_ * np.ones(N), +0.0 * np.ones(N)
    lr = np.float64(learningrate)
    units = [num_units] * num_layers
    wf = ComplexRNNwavefunction(N, units=units, seed=seed, device=device)
    _print_params(wf, verbose)
    ending = _units_ending(units)
    tag = "_N" + str(N) + "_samp" + str(numsamples) + "_lradap" + str(lr) + "_complexGRURNN" + "_J1J2" + str(J2[0]) + ending   # :182-188
    return _run(wf, J1J2(J1, J2, Bz, Marshall_sign), numsteps, numsamples, _schedule(lr, lr_schedule), numsamples, checkpoint_dir,
                "meanEnergy" + tag + "_zeromag.npy", "varEnergy" + tag + "_zeromag.npy", "RNNwavefunction" + tag + "_zeromag.npz",
                save, verbose, resume, complex_energy=True)
