"""Host-side mirror of the reference's wave-function classes.

Same constructor arguments, method names and result shapes as
    1DTFIM/RNNwavefunction.py:7            RNNwavefunction              -> RNNwavefunction1D
    1DTFIM/RNNwavefunction_paritysym.py:7  RNNwavefunction (parity)     -> RNNwavefunctionParity
    2DTFIM_1DRNN/RNNwavefunction.py:8      RNNwavefunction (flat 2-D)   -> RNNwavefunction2DFlat
    2DTFIM_2DRNN/RNNwavefunction.py:5      RNNwavefunction (MDRNN)      -> RNNwavefunction2D
    J1J2/ComplexRNNwavefunction.py:15      RNNwavefunction (complex)    -> ComplexRNNwavefunction
The TF graph is gone: `.sample` / `.log_probability` / `.log_amplitude` launch the sm_100a kernels
through the C ABI and return CUDA tensors (the analogue of the reference's tf.Tensor; `Session.run`
below plays the role of `sess.run` and returns NumPy arrays).  There is no CPU path.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops, params as P
from ._lib import CELL_GRU, CELL_MDRNN, F32, F64, HEAD_COMPLEX, HEAD_PROB, PARITY_SYM


class Session:
    """Stand-in for tf.Session: `run` materialises device tensors as NumPy arrays (the reference's `sess.run(samples_)`,
    1DTFIM/TrainingRNN_1DTFIM.py:203).  Device-to-host copies go through a cached pinned staging buffer; samples cross PCIe as the
    one byte per site the sampler produced and are widened to the reference's int64 on the host (a pageable 80 MB int64 copy of
    10^4 x 1000 samples takes 37 ms, this path 4 ms)."""

    def __init__(self):
        self._pinned = {}

    def _stage(self, t):
        key = (t.dtype, str(t.device))
        buf = self._pinned.get(key)
        if buf is None or buf.numel() < t.numel():
            buf = torch.empty(t.numel(), dtype=t.dtype, pin_memory=True)
            self._pinned[key] = buf
        view = buf[:t.numel()].view(t.shape)
        view.copy_(t, non_blocking=True)
        torch.cuda.current_stream(t.device).synchronize()
        return view

    def run(self, fetches, feed_dict=None):
        if callable(fetches):
            fetches = fetches(**(feed_dict or {}))
        if isinstance(fetches, (list, tuple)):
            return type(fetches)(self.run(f) for f in fetches)
        if isinstance(fetches, torch.Tensor):
            t = fetches.detach()
            if not t.is_cuda:
                return t.numpy()
            narrow = getattr(fetches, "_rnnwf_u8", None)       # samples: the sampler's uint8 tensor behind the int64 view of the API
            if narrow is not None and narrow.numel() == t.numel():
                out = torch.empty(t.shape, dtype=t.dtype)
                out.copy_(self._stage(narrow.contiguous()).view(t.shape))      # widening on the host, multi-threaded
                return out.numpy()
            return self._stage(t.contiguous()).clone().numpy()
        return fetches


def _device(device):
    if device is None:
        if not torch.cuda.is_available():
            raise RuntimeError("rnnwavefunctions_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        device = torch.device("cuda", torch.cuda.current_device())
    return torch.device(device)


class _WavefunctionBase:
    """Shared plumbing: flat parameter tensor (TF-variable order), seeds, (de)serialisation."""

    _flags = 0

    def _setup(self, model, shapes, seed, device, mdrnn=False, pad=None):
        self.model = model
        self.shapes = shapes
        self.device = _device(device)
        self.seed = int(seed)
        self.dtype = torch.float32 if model.dtype == F32 else torch.float64
        npdtype = np.float32 if model.dtype == F32 else np.float64
        self.params = torch.tensor(P.init_flat(shapes, seed, npdtype, mdrnn=mdrnn), device=self.device)
        # unequal layer widths: the kernels see the stack zero-padded to its widest layer (params.gru_pad_index); `params` stays the
        # real TF-order vector (what the optimiser updates and what .npz files hold)
        self._pad_index, self._kbuf = None, None
        if pad is not None:
            index, padded_count = pad
            assert padded_count == ops.param_count(model), (padded_count, ops.param_count(model))
            self._pad_index = torch.as_tensor(index, device=self.device)
            self._kbuf = torch.zeros(padded_count, dtype=self.dtype, device=self.device)
        else:
            assert self.params.numel() == ops.param_count(model), (self.params.numel(), ops.param_count(model))
        if model.cell == CELL_GRU and model.dtype == F32 and model.units > 50 and ops.tfim_chain_mode(model) == 0:
            import warnings
            warnings.warn(f"float32 GRU stacks wider than 50 units (here {model.units}) evaluate local energies on the CUDA-core FFMA engine, "
                          "about 8x slower per flop than the tcgen05 chain kernel that covers 26..50 units and up to 3 layers "
                          "(rnnwavefunctions_b200/csrc/gru_tc16p.cuh)", RuntimeWarning, stacklevel=3)
        self._draws = 0          # number of sample() calls so far: each call uses a fresh Philox stream offset
        self.sample_offset = 0   # global id of this rank's first sample (set by the data-parallel driver)
        self.graph = None        # the reference exposes .graph (TrainingRNN_1DTFIM.py:107); nothing to expose here
        self.samples = None
        self.log_probs = None

    # -- parameters ---------------------------------------------------------------------------------
    @property
    def num_params(self):
        return self.params.numel()

    @property
    def kernel_params(self):
        """The flat parameter vector the CUDA kernels consume: `params` itself, or its zero-padded image (unequal layer widths)."""
        if self._pad_index is None:
            return self.params
        self._kbuf.index_copy_(0, self._pad_index, self.params)      # padding entries are never written: they stay zero
        return self._kbuf

    def unpad_gradient(self, grad):
        """Gradient with respect to `params` from the kernels' gradient with respect to `kernel_params`."""
        return grad if self._pad_index is None else grad.index_select(0, self._pad_index)

    def named_parameters(self):
        return P.split_flat(self.params.detach().cpu().numpy(), self.shapes)

    def set_named_parameters(self, named):
        npdtype = np.float32 if self.model.dtype == F32 else np.float64
        self.params.copy_(torch.tensor(P.join_named(named, self.shapes, npdtype), device=self.device))

    def save_npz(self, path, **extra):
        np.savez(path, **self.named_parameters(), **extra)

    def load_npz(self, path):
        with np.load(path) as z:
            self.set_named_parameters({k: z[k] for k in z.files})

    # -- helpers -------------------------------------------------------------------------------------
    def _u8(self, samples):
        return ops.as_u8_samples(samples, self.device, self.model.n_sites)

    def _next_seed(self):
        # one Philox key per (seed, call index): calls never reuse a stream, ranks share the key and
        # differ by sample id, so the union of samples does not depend on the number of GPUs.
        s = (self.seed * 0x9E3779B97F4A7C15 + self._draws * 0xD1B54A32D192ED03) & (2 ** 64 - 1)
        self._draws += 1
        return s


class RNNwavefunction1D(_WavefunctionBase):
    """1-D positive RNN wave function: stacked GRU + Dense(2)+softmax (1DTFIM/RNNwavefunction.py:8-33)."""

    def __init__(self, systemsize, cell=None, units=[10], scope="RNNwavefunction", seed=111, device=None):
        self.N = systemsize
        self.scope = scope
        model = ops.make_model(CELL_GRU, HEAD_PROB, F32, len(units), max(units), systemsize)
        pad = P.gru_pad_index(units, scope=scope) if len(set(units)) != 1 else None
        self._setup(model, P.gru_shapes(list(units), scope=scope), seed, device, pad=pad)

    def sample(self, numsamples, inputdim=2):
        """-> int64 CUDA tensor [numsamples, N] of 0/1 (1DTFIM/RNNwavefunction.py:35-74)."""
        assert inputdim == 2
        self.numsamples, self.inputdim, self.outputdim = numsamples, inputdim, inputdim
        self._last_u8 = ops.sample(self.model, self.kernel_params, numsamples, self._next_seed(), self.sample_offset)
        self.samples = self._last_u8.to(torch.int64)
        self.samples._rnnwf_u8 = self._last_u8
        return self.samples

    def log_probability(self, samples, inputdim=2):
        """-> float64 CUDA tensor [numsamples] (1DTFIM/RNNwavefunction.py:76-118)."""
        assert inputdim == 2
        self.log_probs = ops.logpsi(self.model, self.kernel_params, self._u8(samples), self._flags)
        return self.log_probs


class RNNwavefunctionParity(RNNwavefunction1D):
    """Parity-symmetrised variant: log(0.5 (P(s) + P(reversed s))) (RNNwavefunction_paritysym.py:125,145),
    evaluated as a log-add-exp (the literal exp underflows for N ~ 1000).  `sample` is the
    plain autoregressive sampler, as in the reference (:35-78)."""
    _flags = PARITY_SYM


class RNNwavefunction2DFlat(RNNwavefunction1D):
    """1-D RNN over the flattened Nx x Ny lattice in float64 (2DTFIM_1DRNN/RNNwavefunction.py:9-38)."""

    def __init__(self, systemsize_x, systemsize_y, cell=None, activation=None, units=[10], scope="RNNwavefunction", seed=111,
                 device=None):
        self.Nx, self.Ny = systemsize_x, systemsize_y
        self.N = systemsize_x * systemsize_y
        self.scope = scope
        model = ops.make_model(CELL_GRU, HEAD_PROB, F64, len(units), max(units), self.N, systemsize_x, systemsize_y)
        pad = P.gru_pad_index(units, scope=scope) if len(set(units)) != 1 else None
        self._setup(model, P.gru_shapes(list(units), scope=scope), seed, device, pad=pad)


class MDRNNcell:
    """Parameter container with the reference's constructor (2DTFIM_2DRNN/MDRNNcell.py:13); `call` evaluates
    one cell step on the device for API parity (the fused kernels do not go through it)."""

    def __init__(self, num_units=None, num_in=None, name="rnn_0", dtype=None, reuse=None):
        self._num_units, self._num_in, self.name = num_units, num_in, name

    @property
    def input_size(self):
        return self._num_in

    @property
    def state_size(self):
        return self._num_units

    @property
    def output_size(self):
        return self._num_units

    def bind(self, wf):
        self._wf = wf
        return self

    def call(self, inputs, states):
        """elu(x_l Uh + h_l Wh + x_u Uv + h_u Wv + b) -> (output, new_state) (MDRNNcell.py:51-66)."""
        named = {k: torch.tensor(v, device=self._wf.device) for k, v in self._wf.named_parameters().items()}
        s = self._wf.scope
        pre = (inputs[0] @ named[f"{s}/Uh_{self.name}"] + states[0] @ named[f"{s}/Wh_{self.name}"]
               + inputs[1] @ named[f"{s}/Uv_{self.name}"] + states[1] @ named[f"{s}/Wv_{self.name}"] + named[f"{s}/b_{self.name}"])
        out = torch.nn.functional.elu(pre)
        return out, out

    __call__ = call


class RNNwavefunction2D(_WavefunctionBase):
    """2-D RNN wave function on the zig-zag path (2DTFIM_2DRNN/RNNwavefunction.py:6-33); samples are
    [numsamples, Nx, Ny] indexed [b, x, y] (:116).  Only units[0] is used, as in the reference (:32)."""

    def __init__(self, systemsize_x, systemsize_y, cell=None, units=[10], scope="RNNwavefunction", seed=111, device=None):
        self.Nx, self.Ny = systemsize_x, systemsize_y
        self.N = systemsize_x * systemsize_y
        self.scope = scope
        model = ops.make_model(CELL_MDRNN, HEAD_PROB, F64, 1, units[0], self.N, systemsize_x, systemsize_y)
        self._setup(model, P.mdrnn_shapes(units[0], scope=scope), seed, device, mdrnn=True)
        self.rnn = MDRNNcell(num_units=units[0], num_in=2, name="rnn_0").bind(self)

    def sample(self, numsamples, inputdim=2):
        assert inputdim == 2
        self.numsamples, self.inputdim, self.outputdim = numsamples, inputdim, inputdim
        self._last_u8 = ops.sample(self.model, self.kernel_params, numsamples, self._next_seed(), self.sample_offset)
        self.samples = self._last_u8.to(torch.int64).reshape(numsamples, self.Nx, self.Ny)
        self.samples._rnnwf_u8 = self._last_u8
        return self.samples

    def log_probability(self, samples, inputdim=2):
        assert inputdim == 2
        self.log_probs = ops.logpsi(self.model, self.kernel_params, self._u8(samples), 0)
        return self.log_probs


class ComplexRNNwavefunction(_WavefunctionBase):
    """Complex RNN wave function with U(1) zero-magnetisation masking (J1J2/ComplexRNNwavefunction.py:16-43)."""

    def __init__(self, systemsize, cell=None, units=[10, 10], scope="RNNwavefunction", seed=111, device=None):
        if systemsize % 2:
            raise ValueError("zero magnetisation needs an even number of sites (SURVEY.md B10)")
        self.N = systemsize
        self.scope = scope
        model = ops.make_model(CELL_GRU, HEAD_COMPLEX, F32, len(units), max(units), systemsize)
        heads = ("wf_dense_ampl", "wf_dense_phase")
        pad = P.gru_pad_index(units, heads=heads, scope=scope) if len(set(units)) != 1 else None
        self._setup(model, P.gru_shapes(list(units), heads=heads, scope=scope), seed, device, pad=pad)

    def sample(self, numsamples, inputdim=2):
        """-> int64 [numsamples, N], every row has N/2 up spins (ComplexRNNwavefunction.py:45-103)."""
        assert inputdim == 2
        self.numsamples, self.inputdim, self.outputdim = numsamples, inputdim, inputdim
        self._last_u8 = ops.sample(self.model, self.kernel_params, numsamples, self._next_seed(), self.sample_offset)
        self.samples = self._last_u8.to(torch.int64)
        self.samples._rnnwf_u8 = self._last_u8
        return self.samples

    def log_amplitude(self, samples, inputdim=2):
        """-> complex128 CUDA tensor [numsamples] (reference: complex64, ComplexRNNwavefunction.py:105-169)."""
        assert inputdim == 2
        self.log_amplitudes = ops.logpsi(self.model, self.kernel_params, self._u8(samples), 0)
        return self.log_amplitudes


units_from_named = P.units_from_named


# The reference gives every class the same name, selected by which directory is on sys.path.
RNNwavefunction = RNNwavefunction1D
