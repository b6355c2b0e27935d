"""CPU oracle for the RNN-wavefunction VMC hot path.  TEST INFRASTRUCTURE ONLY.

This module is a NumPy restatement of the reference algorithm (MatteoMartinelli97/RNNWavefunctions).
It exists to *check* the CUDA path.  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline`
/ `--impl reference` legs of `bench.py` may import it.  The product package `rnnwavefunctions_b200`
never imports anything from `oracle/`.

Parity status: the reference ships no tests and its RNN arithmetic lives in TensorFlow 1.13.1, which
is not installable here.  What *is* pinned:
  * enumeration / diagonal energies / E_loc combine: bit-exact against the reference's own NumPy
    functions executed under a stub `tensorflow` (tests/golden/make_golden.py -> tests/golden/*.npz);
  * GRU structure: parameter counts 422 / 444 (Tutorial_1DTFIM.ipynb#cell15, Tutorial_1DJ1J2.ipynb#cell15);
  * physics: exact diagonalisation / free-fermion energies (Tutorial_1DTFIM.ipynb#cell8, #cell24).
The floating-point RNN arithmetic itself is "parity unpinned" against TF1.13 (no TF run available);
it is cross-checked against `torch.nn.GRUCell` (same equations, tests/test_oracle.py).

All `file:line` citations are relative to /root/reference.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np

# ----------------------------------------------------------------------------------------------
# Parameter inventory (TF1.13 variable names / creation order)
# ----------------------------------------------------------------------------------------------

SCOPE = "RNNwavefunction"


def gru_param_shapes(units, inputdim=2, heads=("wf_dense",), scope=SCOPE):
    """Ordered {tf_variable_name: shape} of a stacked CudnnCompatibleGRUCell + Dense head(s).

    Follows 1DTFIM/RNNwavefunction.py:32-33 (MultiRNNCell + Dense(2)) and
    J1J2/ComplexRNNwavefunction.py:40-43 (two heads).  Per-cell creation order is the one of
    TF1.13 `CudnnCompatibleGRUCell.build` (gate kernel, gate bias, candidate input kernel,
    candidate hidden kernel, candidate input bias, candidate hidden bias).
    """
    shapes = OrderedDict()
    d = inputdim
    for l, h in enumerate(units):
        base = f"{scope}/multi_rnn_cell/cell_{l}/cudnn_compatible_gru_cell/"
        shapes[base + "gates/kernel"] = (d + h, 2 * h)
        shapes[base + "gates/bias"] = (2 * h,)
        shapes[base + "candidate/input_projection/kernel"] = (d, h)
        shapes[base + "candidate/hidden_projection/kernel"] = (h, h)
        shapes[base + "candidate/input_projection/bias"] = (h,)
        shapes[base + "candidate/hidden_projection/bias"] = (h,)
        d = h
    for head in heads:
        shapes[f"{scope}/{head}/kernel"] = (units[-1], 2)
        shapes[f"{scope}/{head}/bias"] = (2,)
    return shapes


def mdrnn_param_shapes(h, inputdim=2, scope=SCOPE):
    """2DTFIM_2DRNN/MDRNNcell.py:21-35 (creation order Wh, Uh, Wv, Uv, b) + Dense (RNNwavefunction.py:33)."""
    shapes = OrderedDict()
    shapes[f"{scope}/Wh_rnn_0"] = (h, h)
    shapes[f"{scope}/Uh_rnn_0"] = (inputdim, h)
    shapes[f"{scope}/Wv_rnn_0"] = (h, h)
    shapes[f"{scope}/Uv_rnn_0"] = (inputdim, h)
    shapes[f"{scope}/b_rnn_0"] = (h,)
    shapes[f"{scope}/wf_dense/kernel"] = (h, 2)
    shapes[f"{scope}/wf_dense/bias"] = (2,)
    return shapes


def _glorot(rng, shape, dtype):
    if len(shape) == 1:
        fan_in = fan_out = shape[0]
    else:
        fan_in, fan_out = shape[0], shape[1]
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=shape).astype(dtype)


def init_gru_params(units, seed=111, dtype=np.float32, inputdim=2, heads=("wf_dense",), scale=1.0):
    """Random-init weights with the reference's initialiser *distributions* (SURVEY.md A.3):
    glorot-uniform kernels, gate bias 1, candidate biases 0, dense glorot / bias 0.
    `scale` multiplies the kernels (tests use >1 to get non-trivial conditionals)."""
    rng = np.random.default_rng(seed)
    p = OrderedDict()
    for name, shape in gru_param_shapes(units, inputdim, heads).items():
        if name.endswith("gates/bias"):
            p[name] = np.ones(shape, dtype)
        elif name.endswith("bias"):
            p[name] = np.zeros(shape, dtype)
        else:
            p[name] = (_glorot(rng, shape, np.float64) * scale).astype(dtype)
    return p


def init_mdrnn_params(h, seed=111, dtype=np.float64, inputdim=2, scale=1.0):
    """MDRNNcell.py:21-35: all five cell tensors xavier-uniform (bias included); Dense glorot/0."""
    rng = np.random.default_rng(seed)
    p = OrderedDict()
    for name, shape in mdrnn_param_shapes(h, inputdim).items():
        if name.endswith("wf_dense/bias"):
            p[name] = np.zeros(shape, dtype)
        else:
            p[name] = (_glorot(rng, shape, np.float64) * scale).astype(dtype)
    return p


def randomize_biases(p, seed=5, amp=0.3):
    """Give every bias a non-trivial value (so tests exercise the bias paths)."""
    rng = np.random.default_rng(seed)
    for k in p:
        if k.endswith("bias") or "/b_rnn" in k:
            p[k] = (p[k] + rng.uniform(-amp, amp, size=p[k].shape)).astype(p[k].dtype)
    return p


def flatten(p):
    return np.concatenate([np.asarray(v).reshape(-1) for v in p.values()])


def unflatten(flat, shapes, dtype=None):
    p = OrderedDict()
    o = 0
    for name, shape in shapes.items():
        n = int(np.prod(shape))
        a = np.asarray(flat[o:o + n]).reshape(shape)
        p[name] = a.astype(dtype) if dtype is not None else a.copy()
        o += n
    assert o == len(flat)
    return p


def num_params(shapes):
    return int(sum(int(np.prod(s)) for s in shapes.values()))


def _units_of(p):
    units = []
    l = 0
    while f"{SCOPE}/multi_rnn_cell/cell_{l}/cudnn_compatible_gru_cell/gates/bias" in p:
        units.append(p[f"{SCOPE}/multi_rnn_cell/cell_{l}/cudnn_compatible_gru_cell/gates/bias"].shape[0] // 2)
        l += 1
    return units


# ----------------------------------------------------------------------------------------------
# Cells
# ----------------------------------------------------------------------------------------------

def _sigmoid(x):
    return (1.0 / (1.0 + np.exp(-x))).astype(x.dtype)


def gru_cell(p, l, x, h):
    """One CudnnCompatibleGRUCell step (TF1.13 contrib):
    [r|u] = sigmoid([x,h] Kg + bg); c = tanh(x Kci + bci + r*(h Kch + bch)); h' = (1-u) c + u h."""
    base = f"{SCOPE}/multi_rnn_cell/cell_{l}/cudnn_compatible_gru_cell/"
    H = h.shape[1]
    g = _sigmoid(np.concatenate([x, h], 1) @ p[base + "gates/kernel"] + p[base + "gates/bias"])
    r, u = g[:, :H], g[:, H:]
    c = np.tanh(x @ p[base + "candidate/input_projection/kernel"] + p[base + "candidate/input_projection/bias"]
                + r * (h @ p[base + "candidate/hidden_projection/kernel"] + p[base + "candidate/hidden_projection/bias"]))
    return ((1 - u) * c + u * h).astype(h.dtype)


def gru_stack(p, x, hs):
    """MultiRNNCell (1DTFIM/RNNwavefunction.py:32): layer l input = layer l-1 output."""
    out = []
    for l, h in enumerate(hs):
        h2 = gru_cell(p, l, x, h)
        out.append(h2)
        x = h2
    return x, out


def _softmax2(z):
    z = z - z.max(axis=1, keepdims=True)
    e = np.exp(z)
    return (e / e.sum(axis=1, keepdims=True)).astype(z.dtype)


def _dense(p, head, h):
    return h @ p[f"{SCOPE}/{head}/kernel"] + p[f"{SCOPE}/{head}/bias"]


def _onehot(col, dtype):
    o = np.zeros((len(col), 2), dtype)
    o[np.arange(len(col)), col] = 1
    return o


# ----------------------------------------------------------------------------------------------
# Philox4x32-10 counter RNG (shared convention with the CUDA sampler)
#   key = (seed_lo, seed_hi); counter = (sample_id_lo, sample_id_hi, site, stream)
#   u = (x0 >> 8) * 2^-24  in [0,1);   draw = 1 if u >= p[0] else 0
# ----------------------------------------------------------------------------------------------

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint64(0x9E3779B9), np.uint64(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = [np.asarray(c, np.uint64) & _MASK for c in (c0, c1, c2, c3)]
    k0 = np.uint64(k0) & _MASK
    k1 = np.uint64(k1) & _MASK
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & _MASK, lo1, (hi0 ^ c3 ^ k1) & _MASK, lo0
        k0 = (k0 + _W0) & _MASK
        k1 = (k1 + _W1) & _MASK
    return c0, c1, c2, c3


def philox_uniform(seed, sample_ids, site, stream=0):
    ids = np.asarray(sample_ids, np.uint64)
    x0, _, _, _ = philox4x32(ids & _MASK, ids >> np.uint64(32), np.uint64(site), np.uint64(stream),
                             np.uint64(seed) & _MASK, np.uint64(seed) >> np.uint64(32))
    return ((x0 >> np.uint64(8)).astype(np.float64) * (1.0 / 16777216.0)).astype(np.float32)


# ----------------------------------------------------------------------------------------------
# 1-D pRNN  (1DTFIM/RNNwavefunction.py; 2-D flat variant 2DTFIM_1DRNN/RNNwavefunction.py)
# ----------------------------------------------------------------------------------------------

def gru_conditionals(p, samples, return_states=False):
    """Teacher-forced pass (1DTFIM/RNNwavefunction.py:97-111): probs[b,n,:] = softmax(Dense(h_n)).
    First input is the zero vector (:97-100), then onehot(samples[:,n-1]) (:111)."""
    dtype = next(iter(p.values())).dtype
    samples = np.asarray(samples)
    B, N = samples.shape
    units = _units_of(p)
    hs = [np.zeros((B, h), dtype) for h in units]
    x = np.zeros((B, 2), dtype)
    probs = np.zeros((B, N, 2), dtype)
    states = []
    for n in range(N):
        out, hs = gru_stack(p, x, hs)
        probs[:, n] = _softmax2(_dense(p, "wf_dense", out))
        x = _onehot(samples[:, n], dtype)
        if return_states:
            states.append([h.copy() for h in hs])
    return (probs, states) if return_states else probs


def log_probability(p, samples):
    """1DTFIM/RNNwavefunction.py:113-116: cast probs to f64, pick p_n[sigma_n], log, sum over sites."""
    samples = np.asarray(samples)
    probs = gru_conditionals(p, samples).astype(np.float64)
    sel = np.take_along_axis(probs, samples[:, :, None].astype(np.int64), axis=2)[:, :, 0]
    return np.log(sel).sum(axis=1)


def log_probability_parity(p, samples, reference_exact=False):
    """1DTFIM/RNNwavefunction_paritysym.py:125-145: log(0.5 (exp lp(s) + exp lp(reversed s))).
    reference_exact=True reproduces the literal f64 exp/log (underflows for lp < -745, Appendix B7);
    the default is the mathematically identical log-add-exp."""
    samples = np.asarray(samples)
    lp1 = log_probability(p, samples)
    lp2 = log_probability(p, samples[:, ::-1])
    if reference_exact:
        with np.errstate(divide="ignore"):
            return np.log(0.5 * (np.exp(lp1) + np.exp(lp2)))
    return np.logaddexp(lp1, lp2) - math.log(2.0)


def sample(p, numsamples, N, seed=111, sample_offset=0):
    """Autoregressive sampling (1DTFIM/RNNwavefunction.py:52-72) with the Philox draw convention
    above (tf.multinomial is not reproducible; only the distribution is)."""
    dtype = next(iter(p.values())).dtype
    units = _units_of(p)
    hs = [np.zeros((numsamples, h), dtype) for h in units]
    x = np.zeros((numsamples, 2), dtype)
    ids = np.arange(numsamples, dtype=np.uint64) + np.uint64(sample_offset)
    out = np.zeros((numsamples, N), np.int64)
    for n in range(N):
        o, hs = gru_stack(p, x, hs)
        pr = _softmax2(_dense(p, "wf_dense", o))
        u = philox_uniform(seed, ids, n)
        s = (u >= pr[:, 0].astype(np.float32)).astype(np.int64)
        out[:, n] = s
        x = _onehot(s, dtype)
    return out


# ----------------------------------------------------------------------------------------------
# complex cRNN  (J1J2/ComplexRNNwavefunction.py)
# ----------------------------------------------------------------------------------------------

def _heavyside(x):
    """ComplexRNNwavefunction.py:11-13: 1 if x >= 0 else 0."""
    return (x >= 0).astype(x.dtype)


def _crnn_site(p, out, n, N, n_up, dtype):
    """amplitude / phase of site n given top-layer output (ComplexRNNwavefunction.py:83-93,143-155)."""
    amp = np.sqrt(_softmax2(_dense(p, "wf_dense_ampl", out)))                 # :5-6, :42
    z = _dense(p, "wf_dense_phase", out)
    phase = (np.pi * (z / (1 + np.abs(z)))).astype(dtype)                      # :8-9, :43
    if n >= N / 2:                                                            # :85 (float compare)
        base = dtype.type(N // 2 - 1)
        n_up = n_up.astype(dtype)
        n_dn = dtype.type(n) - n_up
        mask = np.stack([_heavyside(base - n_dn), _heavyside(base - n_up)], 1)  # :89-92 (index 0 = down)
        amp = amp * mask
        nrm = np.sqrt(np.maximum((amp * amp).sum(1, keepdims=True), dtype.type(1e-30)))  # l2_normalize :93
        amp = (amp / nrm).astype(dtype)
    return amp, phase


def crnn_log_amplitude(p, samples):
    """ComplexRNNwavefunction.py:105-169 -> complex64 [B] (complex128 if the params are f64)."""
    dtype = next(iter(p.values())).dtype
    samples = np.asarray(samples)
    B, N = samples.shape
    units = _units_of(p)
    hs = [np.zeros((B, h), dtype) for h in units]
    x = np.zeros((B, 2), dtype)
    ctype = np.complex64 if dtype == np.float32 else np.complex128
    acc = np.zeros(B, ctype)
    for n in range(N):
        out, hs = gru_stack(p, x, hs)
        amp, phase = _crnn_site(p, out, n, N, samples[:, :n].sum(1), dtype)
        a = np.take_along_axis(amp, samples[:, n:n + 1].astype(np.int64), 1)[:, 0]
        ph = np.take_along_axis(phase, samples[:, n:n + 1].astype(np.int64), 1)[:, 0]
        with np.errstate(divide="ignore"):
            acc = acc + (np.log(a.astype(ctype)) + 1j * ph.astype(ctype)).astype(ctype)   # :157-167
        x = _onehot(samples[:, n], dtype)
    return acc


def crnn_sample(p, numsamples, N, seed=111, sample_offset=0):
    """ComplexRNNwavefunction.py:45-103 (draw from amp^2, :95) with the Philox convention."""
    dtype = next(iter(p.values())).dtype
    units = _units_of(p)
    hs = [np.zeros((numsamples, h), dtype) for h in units]
    x = np.zeros((numsamples, 2), dtype)
    ids = np.arange(numsamples, dtype=np.uint64) + np.uint64(sample_offset)
    out = np.zeros((numsamples, N), np.int64)
    for n in range(N):
        o, hs = gru_stack(p, x, hs)
        amp, _ = _crnn_site(p, o, n, N, out[:, :n].sum(1), dtype)
        p0 = (amp[:, 0] * amp[:, 0]).astype(np.float32)
        p1 = (amp[:, 1] * amp[:, 1]).astype(np.float32)
        u = philox_uniform(seed, ids, n)
        s = (u * (p0 + p1) >= p0).astype(np.int64)
        s = np.where(p1 == 0, 0, np.where(p0 == 0, 1, s))
        out[:, n] = s
        x = _onehot(s, dtype)
    return out


# ----------------------------------------------------------------------------------------------
# 2-D RNN (2DTFIM_2DRNN/MDRNNcell.py, 2DTFIM_2DRNN/RNNwavefunction.py)
# ----------------------------------------------------------------------------------------------

def _elu(x):
    return np.where(x > 0, x, np.expm1(np.minimum(x, 0))).astype(x.dtype)


def mdrnn_cell(p, x_l, x_u, h_l, h_u):
    """MDRNNcell.py:51-66: elu(x_l Uh + h_l Wh + x_u Uv + h_u Wv + b); output == new state."""
    pre = (x_l @ p[f"{SCOPE}/Uh_rnn_0"] + h_l @ p[f"{SCOPE}/Wh_rnn_0"]
           + x_u @ p[f"{SCOPE}/Uv_rnn_0"] + h_u @ p[f"{SCOPE}/Wv_rnn_0"] + p[f"{SCOPE}/b_rnn_0"])
    return _elu(pre)


def mdrnn_path(Nx, Ny):
    """Zig-zag visiting order (RNNwavefunction.py:90-113): rows y ascending; x ascending on even rows,
    descending on odd rows.  Returns list of (x, y, x_horizontal_neighbour)."""
    path = []
    for y in range(Ny):
        xs = range(Nx) if y % 2 == 0 else range(Nx - 1, -1, -1)
        for x in xs:
            path.append((x, y, x - 1 if y % 2 == 0 else x + 1))
    return path


def mdrnn_conditionals(p, samples=None, numsamples=None, Nx=None, Ny=None, seed=111, sample_offset=0):
    """Teacher-forced (samples given; RNNwavefunction.py:120-200) or sampling (samples None; :35-118).
    samples are indexed [b, x, y] (:116).  Returns (probs[b,x,y,2], samples[b,x,y])."""
    dtype = next(iter(p.values())).dtype
    H = p[f"{SCOPE}/b_rnn_0"].shape[0]
    if samples is not None:
        samples = np.asarray(samples)
        B, Nx, Ny = samples.shape
        out = samples.astype(np.int64)
        draw = False
    else:
        B = numsamples
        out = np.zeros((B, Nx, Ny), np.int64)
        draw = True
    ids = np.arange(B, dtype=np.uint64) + np.uint64(sample_offset)
    zero_h = np.zeros((B, H), dtype)
    zero_x = np.zeros((B, 2), dtype)
    hgrid, xin = {}, {}
    probs = np.zeros((B, Nx, Ny, 2), dtype)
    for pos, (x, y, xn) in enumerate(mdrnn_path(Nx, Ny)):
        h_l = hgrid.get((xn, y), zero_h)
        x_l = xin.get((xn, y), zero_x)
        h_u = hgrid.get((x, y - 1), zero_h)
        x_u = xin.get((x, y - 1), zero_x)
        h = mdrnn_cell(p, x_l, x_u, h_l, h_u)
        hgrid[(x, y)] = h
        pr = _softmax2(_dense(p, "wf_dense", h))
        probs[:, x, y] = pr
        if draw:
            u = philox_uniform(seed, ids, pos)
            out[:, x, y] = (u >= pr[:, 0].astype(np.float32)).astype(np.int64)
        xin[(x, y)] = _onehot(out[:, x, y], dtype)
    return probs, out


def mdrnn_log_probability(p, samples):
    samples = np.asarray(samples)
    probs, _ = mdrnn_conditionals(p, samples)
    sel = np.take_along_axis(probs.astype(np.float64), samples[..., None].astype(np.int64), 3)[..., 0]
    return np.log(sel).sum(axis=(1, 2))          # RNNwavefunction.py:195-198


def mdrnn_sample(p, numsamples, Nx, Ny, seed=111, sample_offset=0):
    return mdrnn_conditionals(p, None, numsamples, Nx, Ny, seed, sample_offset)[1]


# ----------------------------------------------------------------------------------------------
# Hamiltonians / local energies
# ----------------------------------------------------------------------------------------------

def tfim1d_diag(Jz, samples):
    """1DTFIM/TrainingRNN_1DTFIM.py:31-38: -sum_i Jz[i] s_i s_{i+1}, accumulated bond by bond in f64."""
    samples = np.asarray(samples)
    e = np.zeros(samples.shape[0], np.float64)
    for i in range(samples.shape[1] - 1):
        same = samples[:, i] == samples[:, i + 1]
        e += np.where(same, 1, -1) * (-Jz[i])
    return e


def tfim1d_queue(samples):
    """1DTFIM/TrainingRNN_1DTFIM.py:40-48: slot 0 = samples, slot i+1 = samples with site i flipped."""
    samples = np.asarray(samples)
    ns, N = samples.shape
    q = np.repeat(samples[None].astype(np.int32), N + 1, axis=0)
    idx = np.arange(N)
    q[idx + 1, :, idx] = 1 - q[idx + 1, :, idx]
    return q


def _chunked(fn, configs, chunk=25000):
    """1DTFIM/TrainingRNN_1DTFIM.py:56-65: ceil(len/25000) chunks with integer-division bounds."""
    n = len(configs)
    steps = max(1, math.ceil(n / chunk))
    out = None
    for i in range(steps):
        lo = (i * n) // steps
        hi = ((i + 1) * n) // steps if i < steps - 1 else n
        r = np.asarray(fn(configs[lo:hi]))
        if out is None:
            out = np.zeros(n, r.dtype)
        out[lo:hi] = r
    return out


def ising_local_energies(Jz, Bx, samples, logprob_fn, chunk=25000):
    """Full-recompute restatement of Ising_local_energies (1DTFIM/TrainingRNN_1DTFIM.py:13-75):
    E = diag - Bx sum_i exp(0.5 (lp_i - lp_0)) over all N single-flip configurations (:74)."""
    samples = np.asarray(samples)
    ns, N = samples.shape
    e = tfim1d_diag(Jz, samples)
    if Bx != 0:
        q = tfim1d_queue(samples).reshape((N + 1) * ns, N)
        lp = _chunked(logprob_fn, q, chunk).reshape(N + 1, ns)
        e += -Bx * np.exp(0.5 * lp[1:] - 0.5 * lp[0]).sum(axis=0)
    return e


def tfim2d_diag(Jz, samples_xy):
    """2DTFIM_2DRNN/Training2DRNN_2DTFIM.py:33-49 (identical in the 1DRNN file): bonds along axis 1
    weighted Jz[i,:], then bonds along axis 2 weighted Jz[:,i]; f64 accumulation in that order."""
    s = np.asarray(samples_xy)
    ns, Nx, Ny = s.shape
    e = np.zeros(ns, np.float64)
    for i in range(Nx - 1):
        v = np.where(s[:, i] == s[:, i + 1], 1, -1)
        e += np.sum(v * (-Jz[i, :]), axis=1)
    for i in range(Ny - 1):
        v = np.where(s[:, :, i] == s[:, :, i + 1], 1, -1)
        e += np.sum(v * (-Jz[:, i]), axis=1)
    return e


def ising2d_local_energies(Jz, Bx, Nx, Ny, samples, logprob_fn, flat, chunk=25000):
    """Ising2D_local_energies: flat=True -> 2DTFIM_1DRNN/Training1DRNN_2DTFIM.py:13-81 (samples
    [ns,Nx*Ny], reshaped [ns,Nx,Ny] for the bonds :27, flips in flat order :55-60); flat=False ->
    2DTFIM_2DRNN/Training2DRNN_2DTFIM.py:13-83 (samples [ns,Nx,Ny], flip (i,j) -> slot i*Ny+j+1 :55-61)."""
    samples = np.asarray(samples)
    ns = samples.shape[0]
    N = Nx * Ny
    e = tfim2d_diag(Jz, samples.reshape(ns, Nx, Ny))
    if Bx != 0:
        q = tfim1d_queue(samples.reshape(ns, N))            # slot i*Ny+j+1 == flat index i*Ny+j of [Nx,Ny]
        q = q.reshape((N + 1) * ns, N) if flat else q.reshape((N + 1) * ns, Nx, Ny)
        lp = _chunked(logprob_fn, q, chunk).reshape(N + 1, ns)
        e += -Bx * np.exp(0.5 * lp[1:] - 0.5 * lp[0]).sum(axis=0)
    return e


def j1j2_matrix_elements(J1, J2, Bz, sigma, periodic=False, marshall_sign=False):
    """Restatement of J1J2MatrixElements (J1J2/TrainingRNN_J1J2.py:12-93) for one configuration.
    Returns (configs[num,N] int32, elements[num] float32).  Order: diagonal, NN exchanges by
    ascending site, NNN exchanges by ascending site.  Diagonal accumulated in f64 in the reference's
    order (:30-57) then stored as f32 (:59 with the f32 buffer of :239)."""
    sigma = np.asarray(sigma)
    N = len(Bz)
    lim1 = N if periodic else N - 1
    lim2 = N if periodic else N - 2
    diag = np.dot(sigma - 0.5, Bz)
    for s in range(lim1):
        diag += (-0.25 if sigma[s] != sigma[(s + 1) % N] else 0.25) * J1[s]
    for s in range(lim2):
        if J2[s] != 0.0:
            diag += (-0.25 if sigma[s] != sigma[(s + 2) % N] else 0.25) * J2[s]
    cfgs = [sigma.astype(np.int32).copy()]
    els = [np.float32(diag)]
    for dist, J, lim in ((1, J1, lim1), (2, J2, lim2)):
        for s in range(lim):
            t = (s + dist) % N
            if J[s] != 0.0 and sigma[s] != sigma[t]:
                c = sigma.astype(np.int32).copy()
                c[s], c[t] = sigma[t], sigma[s]
                cfgs.append(c)
                els.append(np.float32((-J[s] if (marshall_sign and dist == 1) else J[s]) / 2))
    return np.stack(cfgs), np.asarray(els, np.float32)


def j1j2_local_energies(J1, J2, Bz, samples, logamp_fn, periodic=False, marshall_sign=False, chunk=30000):
    """J1J2Slices + chunked log-amplitudes + combine (J1J2/TrainingRNN_J1J2.py:95-127, :255-279):
    E_n = H[s] . exp(la[s] - la[s][0]) in complex64."""
    samples = np.asarray(samples)
    allc, allh, slices, o = [], [], [], 0
    for s in samples:
        c, h = j1j2_matrix_elements(J1, J2, Bz, s, periodic, marshall_sign)
        allc.append(c)
        allh.append(h)
        slices.append(slice(o, o + len(h)))
        o += len(h)
    sig = np.concatenate(allc)
    Hm = np.concatenate(allh)
    la = _chunked(logamp_fn, sig, chunk).astype(np.complex64)
    e = np.zeros(len(samples), np.complex64)
    for n, sl in enumerate(slices):
        e[n] = Hm[sl].dot(np.exp(la[sl] - la[sl][0]))
    return e


# ----------------------------------------------------------------------------------------------
# Optimiser (TF1 Adam) and exact energies
# ----------------------------------------------------------------------------------------------

def adam_tf1(theta, g, m, v, t, lr, b1=0.9, b2=0.999, eps=1e-8):
    """tf.train.AdamOptimizer.apply_gradients: eps is added to sqrt(v) (not to sqrt(v_hat))."""
    t = t + 1
    lr_t = lr * math.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    theta = theta - lr_t * m / (np.sqrt(v) + eps)
    return theta, m, v, t


def tfim1d_exact_energy(N, Bx=1.0, Jz=1.0):
    """Free-fermion ground-state energy of the open 1-D TFIM (cross-checked against the DMRG table of
    Tutorials/1DTFIM/Tutorial_1DTFIM.ipynb#cell24 in SURVEY.md §4)."""
    M = Bx * np.eye(N) + Jz * np.eye(N, k=1)
    return -np.linalg.svd(M, compute_uv=False).sum()


def all_configs(N):
    return ((np.arange(2 ** N)[:, None] >> np.arange(N - 1, -1, -1)) & 1).astype(np.int32)
