"""Torch-tensor level wrappers over the C ABI (device pointers in, device tensors out).

PyTorch is plumbing here: it owns device memory and streams.  All arithmetic happens in librnnwf_b200.so.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from ._lib import (CELL_GRU, CELL_MDRNN, F32, F64, HEAD_COMPLEX, HEAD_PROB, OP_J1J2_ELOC, OP_LOGPSI, OP_SAMPLE,
                   OP_TFIM_ELOC, OP_VMC_GRAD, PARITY_SYM, Model, check)


def make_model(cell=CELL_GRU, head=HEAD_PROB, dtype=F32, num_layers=1, units=10, n_sites=1, nx=0, ny=0) -> Model:
    return Model(cell, head, dtype, num_layers, units, n_sites, nx, ny)


def torch_dtype(model: Model):
    return torch.float32 if model.dtype == F32 else torch.float64


def param_count(model: Model) -> int:
    n = _lib.load().rnnwf_param_count(C.byref(model))
    if n < 0:
        check(-1)
    return int(n)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _on_device_of(index):
    """Decorator: run the wrapped C-ABI call with the CUDA device of positional argument `index` current (the
    library launches on the current device and on the current stream of that device), so wave functions built
    with device='cuda:1' work while another device is current."""
    import functools

    def deco(fn):
        @functools.wraps(fn)
        def wrapped(*args, **kw):
            t = args[index]
            if not isinstance(t, torch.Tensor) or not t.is_cuda:
                raise ValueError(f"{fn.__name__}: argument {index} must be a CUDA tensor")
            with torch.cuda.device(t.device):
                return fn(*args, **kw)
        return wrapped
    return deco


class _Workspace:
    """Grow-only device scratch buffer per device (the library itself keeps no device state)."""

    def __init__(self):
        self.buf = {}

    def get(self, nbytes: int, device):
        key = str(device)
        b = self.buf.get(key)
        if b is None or b.numel() < nbytes:
            self.buf.pop(key, None)
            b = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
            self.buf[key] = b
        return b


_WS = _Workspace()


def release_workspace():
    _WS.buf.clear()


def _ws_for(model, op, ns, flags, device):
    n = _lib.load().rnnwf_workspace_bytes(C.byref(model), op, ns, flags)
    if n == 0:
        raise _lib.RnnwfError(f"workspace query failed for op {op}: {_lib.load().rnnwf_last_error().decode()}")
    return _WS.get(n, device), n


def workspace_bytes(model, op, ns, flags=0) -> int:
    return int(_lib.load().rnnwf_workspace_bytes(C.byref(model), op, int(ns), flags))


def workspace_budget(device) -> int:
    """Bytes one call may use for its workspace: RNNWF_WS_BUDGET_GB, else 80 % of the device memory that is free or already held by
    this module's scratch buffer."""
    env = os.environ.get("RNNWF_WS_BUDGET_GB")
    if env:
        return int(float(env) * 2 ** 30)
    free, _ = torch.cuda.mem_get_info(device)
    held = _WS.buf.get(str(device))
    return int(0.8 * (free + (held.numel() if held is not None else 0)))


def sample_chunk(model, op, ns, flags, device) -> int:
    """Largest number of samples per call whose workspace fits the budget.  The E_loc stash grows as ns * N * L * H and the gradient
    scratch is ~7x that (47 GB for 10^4 samples at N = 1000, 3 x 50), so large batches are evaluated in slices, as the reference
    evaluates its queue of (N + 1) * ns configurations in chunks of `numsamples` (1DTFIM/TrainingRNN_1DTFIM.py:56-72).  Samples are
    independent (E_loc) or enter linearly (gradient): slicing changes nothing but the order of the gradient's final sum."""
    ns = int(ns)
    if ns <= 256:
        return ns
    need = workspace_bytes(model, op, ns, flags)
    held = _WS.buf.get(str(device))
    if held is not None and need <= held.numel():      # fits the scratch buffer this module already holds: no device query
        return ns
    budget = workspace_budget(device)
    if need <= budget:
        return ns
    lo, hi = 1, ns // 256                       # multiples of 256 rows (whole 128-row tiles in both directions of the parity model)
    while lo < hi:
        mid = (lo + hi + 1) // 2
        if workspace_bytes(model, op, mid * 256, flags) <= budget:
            lo = mid
        else:
            hi = mid - 1
    return lo * 256


def _sliced(fn):
    """Decorator for the per-sample ops: evaluate `samples_u8` (positional argument 2) in slices that fit the workspace budget and
    concatenate the per-sample results."""
    import functools
    import inspect
    sig = inspect.signature(fn)

    @functools.wraps(fn)
    def wrapped(model, params, samples_u8, *args, **kw):
        ns = samples_u8.shape[0]
        flags = sig.bind(model, params, samples_u8, *args, **kw).arguments.get("flags", 0)
        step = sample_chunk(model, wrapped.op, ns, flags, params.device)
        if step >= ns:
            return fn(model, params, samples_u8, *args, **kw)
        parts = [fn(model, params, samples_u8[i:i + step].contiguous(), *args, **kw) for i in range(0, ns, step)]
        if isinstance(parts[0], tuple):
            return tuple(None if col[0] is None else torch.cat(col) for col in zip(*parts))
        return torch.cat(parts)
    return wrapped


def _check_params(model, params):
    if not params.is_cuda or params.dtype != torch_dtype(model) or not params.is_contiguous():
        raise ValueError("params must be a contiguous CUDA tensor of the model dtype")
    if params.numel() != param_count(model):
        raise ValueError(f"params has {params.numel()} entries, model needs {param_count(model)}")


def as_u8_samples(samples, device, n_sites):
    """Accept numpy / torch integer samples of any int dtype and [ns, N] or [ns, Nx, Ny] shape."""
    t = torch.as_tensor(samples)
    t = t.reshape(t.shape[0], -1)
    if t.shape[1] != n_sites:
        raise ValueError(f"samples have {t.shape[1]} sites, model has {n_sites}")
    if t.device.type == "cpu" and t.dtype != torch.uint8:
        t = t.to(torch.uint8)            # narrow on the host: one byte per site crosses PCIe, not the reference's int64
    return t.to(device=device, dtype=torch.uint8, non_blocking=True).contiguous()


@_on_device_of(1)
def sample(model, params, ns, seed=0, sample_offset=0):
    """-> uint8 [ns, N] on params.device."""
    _check_params(model, params)
    ws, nb = _ws_for(model, OP_SAMPLE, ns, 0, params.device)
    out = torch.empty((ns, model.n_sites), dtype=torch.uint8, device=params.device)
    check(_lib.load().rnnwf_sample(C.byref(model), _ptr(params), ns, seed & (2**64 - 1), sample_offset, _ptr(out), _ptr(ws), nb,
                                   _stream()))
    return out


def _op(op):
    def deco(fn):
        w = _sliced(fn)
        w.op = op
        return w
    return deco


@_op(OP_LOGPSI)
@_on_device_of(1)
def logpsi(model, params, samples_u8, flags=0):
    """-> float64 [ns] (probability head) or complex128 [ns] (complex head)."""
    _check_params(model, params)
    ns = samples_u8.shape[0]
    ws, nb = _ws_for(model, OP_LOGPSI, ns, flags, params.device)
    cplx = model.head == HEAD_COMPLEX
    out = torch.empty((ns, 2) if cplx else (ns,), dtype=torch.float64, device=params.device)
    check(_lib.load().rnnwf_logpsi(C.byref(model), _ptr(params), _ptr(samples_u8), ns, flags, _ptr(out), _ptr(ws), nb, _stream()))
    return torch.view_as_complex(out) if cplx else out


@_op(OP_TFIM_ELOC)
@_on_device_of(1)
def tfim_eloc(model, params, samples_u8, jz, bx, flags=0, want_logp=True):
    _check_params(model, params)
    ns = samples_u8.shape[0]
    ws, nb = _ws_for(model, OP_TFIM_ELOC, ns, flags, params.device)
    jz = torch.as_tensor(jz, dtype=torch.float64).reshape(-1).to(params.device).contiguous()
    eloc = torch.empty(ns, dtype=torch.float64, device=params.device)
    logp = torch.empty(ns, dtype=torch.float64, device=params.device) if want_logp else None
    check(_lib.load().rnnwf_tfim_eloc(C.byref(model), _ptr(params), _ptr(samples_u8), ns, _ptr(jz), float(bx), flags, _ptr(eloc),
                                      _ptr(logp), _ptr(ws), nb, _stream()))
    return eloc, logp


@_op(OP_TFIM_ELOC)
@_on_device_of(1)
def tfim_flip_ratios(model, params, samples_u8, jz, bx, flags=0):
    """-> (eloc [ns], logp [ns], ratios [ns, N]) with ratios[s, k] = psi(sigma_s, site k flipped) / psi(sigma_s) (rnnwf_tfim_flip_ratios)."""
    _check_params(model, params)
    ns = samples_u8.shape[0]
    ws, nb = _ws_for(model, OP_TFIM_ELOC, ns, flags, params.device)
    jz = torch.as_tensor(jz, dtype=torch.float64).reshape(-1).to(params.device).contiguous()
    eloc = torch.empty(ns, dtype=torch.float64, device=params.device)
    logp = torch.empty(ns, dtype=torch.float64, device=params.device)
    ratios = torch.empty((ns, model.n_sites), dtype=torch.float64, device=params.device)
    check(_lib.load().rnnwf_tfim_flip_ratios(C.byref(model), _ptr(params), _ptr(samples_u8), ns, _ptr(jz), float(bx), flags, _ptr(eloc),
                                             _ptr(logp), _ptr(ratios), _ptr(ws), nb, _stream()))
    return eloc, logp, ratios


def tfim_chain_mode(model) -> int:
    """3: pipelined tcgen05 3xFP16 chain kernel, 2: unpipelined 3xFP16, 1: tcgen05 3xTF32, 0: CUDA-core FFMA (see include/rnnwf.h)."""
    return int(_lib.load().rnnwf_tfim_chain_mode(C.byref(model)))


@_on_device_of(1)
def tfim_diag(model, samples_u8, jz):
    ns = samples_u8.shape[0]
    jz = torch.as_tensor(jz, dtype=torch.float64).reshape(-1).to(samples_u8.device).contiguous()
    out = torch.empty(ns, dtype=torch.float64, device=samples_u8.device)
    check(_lib.load().rnnwf_tfim_diag(C.byref(model), _ptr(samples_u8), ns, _ptr(jz), _ptr(out), _stream()))
    return out


@_on_device_of(0)
def tfim_enumerate(samples_u8):
    ns, N = samples_u8.shape
    out = torch.empty((N + 1, ns, N), dtype=torch.int32, device=samples_u8.device)
    check(_lib.load().rnnwf_tfim_enumerate(_ptr(samples_u8), ns, N, _ptr(out), _stream()))
    return out


@_on_device_of(0)
def j1j2_enumerate(samples_u8, j1, j2, bz, periodic=False, marshall_sign=False, want_sigmas=True):
    ns, N = samples_u8.shape
    dev = samples_u8.device
    j1, j2, bz = (torch.as_tensor(a, dtype=torch.float64).to(dev).contiguous() for a in (j1, j2, bz))
    rows = 2 * N + 1
    sig = torch.zeros((ns, rows, N), dtype=torch.int32, device=dev) if want_sigmas else None
    el = torch.zeros((ns, rows), dtype=torch.float32, device=dev)
    cnt = torch.zeros(ns, dtype=torch.int32, device=dev)
    check(_lib.load().rnnwf_j1j2_enumerate(_ptr(samples_u8), ns, N, _ptr(j1), _ptr(j2), _ptr(bz), int(periodic), int(marshall_sign),
                                           _ptr(sig), _ptr(el), _ptr(cnt), _stream()))
    return sig, el, cnt


@_op(OP_J1J2_ELOC)
@_on_device_of(1)
def j1j2_eloc(model, params, samples_u8, j1, j2, bz, marshall_sign=False, want_logpsi=True):
    _check_params(model, params)
    ns = samples_u8.shape[0]
    dev = params.device
    ws, nb = _ws_for(model, OP_J1J2_ELOC, ns, 0, dev)
    j1, j2, bz = (torch.as_tensor(a, dtype=torch.float64).to(dev).contiguous() for a in (j1, j2, bz))
    eloc = torch.empty((ns, 2), dtype=torch.float64, device=dev)
    lpsi = torch.empty((ns, 2), dtype=torch.float64, device=dev) if want_logpsi else None
    check(_lib.load().rnnwf_j1j2_eloc(C.byref(model), _ptr(params), _ptr(samples_u8), ns, _ptr(j1), _ptr(j2), _ptr(bz),
                                      int(marshall_sign), _ptr(eloc), _ptr(lpsi), _ptr(ws), nb, _stream()))
    return torch.view_as_complex(eloc), (torch.view_as_complex(lpsi) if want_logpsi else None)


def vmc_grad(model, params, samples_u8, weights, flags=0):
    """weights: float64 [ns] (probability head) or complex128 / float64 [ns,2] (complex head). -> float64 [P].
    The gradient is linear in the per-sample weights: batches whose scratch does not fit the workspace budget are summed slice by slice."""
    ns = samples_u8.shape[0]
    step = sample_chunk(model, OP_VMC_GRAD, ns, flags, params.device)
    if step >= ns:
        return _vmc_grad(model, params, samples_u8, weights, flags)
    total = None
    for i in range(0, ns, step):
        g = _vmc_grad(model, params, samples_u8[i:i + step].contiguous(), weights[i:i + step], flags)
        total = g.clone() if total is None else total.add_(g)
    return total


@_on_device_of(1)
def _vmc_grad(model, params, samples_u8, weights, flags=0):
    _check_params(model, params)
    ns = samples_u8.shape[0]
    ws, nb = _ws_for(model, OP_VMC_GRAD, ns, flags, params.device)
    w = weights
    if w.is_complex():
        w = torch.view_as_real(w.to(torch.complex128))
    w = w.to(device=params.device, dtype=torch.float64).contiguous()
    grad = torch.empty(param_count(model), dtype=torch.float64, device=params.device)
    check(_lib.load().rnnwf_vmc_grad(C.byref(model), _ptr(params), _ptr(samples_u8), ns, _ptr(w), flags, _ptr(grad), _ptr(ws), nb,
                                     _stream()))
    return grad


@_on_device_of(1)
def adam_step(model, theta, mom, vel, grad, t, lr, grad_scale=1.0, beta1=0.9, beta2=0.999, eps=1e-8):
    check(_lib.load().rnnwf_adam_step(model.dtype, theta.numel(), _ptr(theta), _ptr(mom), _ptr(vel), _ptr(grad), float(grad_scale),
                                      float(lr), beta1, beta2, eps, int(t), _stream()))


@_on_device_of(0)
def energy_moments(eloc_f64, stride=1, count=None):
    """-> float64 [3] = (sum, sum of squares, n) over eloc_f64[0], eloc_f64[stride], ... (count entries)."""
    ns = eloc_f64.numel() // stride if count is None else int(count)
    out = torch.empty(3, dtype=torch.float64, device=eloc_f64.device)
    check(_lib.load().rnnwf_energy_moments(_ptr(eloc_f64), ns, stride, _ptr(out), _stream()))
    return out


def profile_begin():
    """Start counting this library's kernel launches and timing the dominant (chain) kernel."""
    check(_lib.load().rnnwf_profile_begin())


def profile_end():
    """-> (kernel launches, dominant-kernel launches, dominant-kernel device ms) since profile_begin()."""
    a, b, c = C.c_int64(0), C.c_int64(0), C.c_double(0.0)
    check(_lib.load().rnnwf_profile_end(C.byref(a), C.byref(b), C.byref(c)))
    return a.value, b.value, c.value


def ffma_peak(iters=20000):
    """Measured FP32 FFMA throughput of the current device in TFLOP/s."""
    out = C.c_double(0.0)
    check(_lib.load().rnnwf_ffma_peak(int(iters), C.byref(out), _stream()))
    return out.value


def fp64_peak(mode=0, iters=5000):
    """Measured FP64 throughput of the current device in TFLOP/s (mode 0: DFMA, 1: DMMA m8n8k4)."""
    out = C.c_double(0.0)
    check(_lib.load().rnnwf_fp64_peak(int(mode), int(iters), C.byref(out), _stream()))
    return out.value


@_on_device_of(0)
def umma_selftest(a, b, passes=3, f16=False, dcol=0):
    """d = a @ b.T on the tcgen05 path (a [128,K], b [N,K] float32 CUDA tensors); f16: FP16 hi/lo operands."""
    n, k = b.shape
    if f16:
        passes = -(passes + 4 * dcol)
    d = torch.empty((128, n), dtype=torch.float32, device=a.device)
    check(_lib.load().rnnwf_umma_selftest(n, k, _ptr(a.contiguous()), _ptr(b.contiguous()), _ptr(d), passes, _stream()))
    return d
