"""Parameter inventory: TF1.13 variable names, shapes, flat order and the reference's initialisers.

Names follow what the reference prints at start-up (1DTFIM/TrainingRNN_1DTFIM.py:127-136): one flat buffer
in TF-variable creation order is what the CUDA library consumes (include/rnnwf.h) and what `.npz`
checkpoints are keyed by, so a TF1.13 user can dump `{v.name: sess.run(v)}` and load it here.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np

SCOPE = "RNNwavefunction"


def gru_shapes(units, inputdim=2, heads=("wf_dense",), scope=SCOPE):
    """MultiRNNCell([CudnnCompatibleGRUCell(u) ...]) + Dense head(s)  (1DTFIM/RNNwavefunction.py:32-33,
    J1J2/ComplexRNNwavefunction.py:40-43).  Creation order inside a cell: gate kernel, gate bias, candidate
    input kernel, candidate hidden kernel, candidate input bias, candidate hidden bias."""
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


def gru_pad_index(units, inputdim=2, heads=("wf_dense",), scope=SCOPE):
    """Unequal layer widths (`MultiRNNCell([cell(units[n]) for n ...])`, 1DTFIM/RNNwavefunction.py:32, accepts any list) run on the
    equal-width kernels zero-padded to the widest layer: a padded unit has zero weights and biases in and out, so its gates are
    exactly 1/2, its candidate exactly 0 and its state stays exactly 0; nothing of it reaches a real unit.
    -> (index, padded_count): index[i] = position of element i of the real flat vector (TF order) in the padded flat vector."""
    units = [int(u) for u in units]
    H = max(units)
    real, padded = gru_shapes(units, inputdim, heads, scope), gru_shapes([H] * len(units), inputdim, heads, scope)
    base, o = {}, 0
    for name, shape in padded.items():
        base[name] = o
        o += int(np.prod(shape))
    padded_count = o

    def pos(segments):          # [(real length, offset in the padded axis), ...] -> padded positions of the real entries
        return np.concatenate([off + np.arange(n) for n, off in segments]).astype(np.int64)

    idx = []
    d = inputdim
    for l, h in enumerate(units):
        D = inputdim if l == 0 else H
        pre = f"{scope}/multi_rnn_cell/cell_{l}/cudnn_compatible_gru_cell/"
        axes = {
            pre + "gates/kernel": ([(d, 0), (h, D)], [(h, 0), (h, H)]),
            pre + "gates/bias": (None, [(h, 0), (h, H)]),
            pre + "candidate/input_projection/kernel": ([(d, 0)], [(h, 0)]),
            pre + "candidate/hidden_projection/kernel": ([(h, 0)], [(h, 0)]),
            pre + "candidate/input_projection/bias": (None, [(h, 0)]),
            pre + "candidate/hidden_projection/bias": (None, [(h, 0)]),
        }
        for name, (rows, cols) in axes.items():
            c = pos(cols)
            if rows is None:
                idx.append(base[name] + c)
            else:
                idx.append((base[name] + pos(rows)[:, None] * padded[name][1] + c[None, :]).reshape(-1))
        d = h
    for head in heads:
        idx.append((base[f"{scope}/{head}/kernel"] + np.arange(units[-1])[:, None] * 2 + np.arange(2)[None, :]).reshape(-1))
        idx.append(base[f"{scope}/{head}/bias"] + np.arange(2))
    index = np.concatenate(idx)
    assert index.size == count(real) and len(np.unique(index)) == index.size and index.max() < padded_count
    return index, padded_count


def mdrnn_shapes(h, inputdim=2, scope=SCOPE):
    """MDRNNcell variables (2DTFIM_2DRNN/MDRNNcell.py:21-35, name 'rnn_0' from RNNwavefunction.py:32) + Dense."""
    shapes = OrderedDict()
    shapes[f"{scope}/Wh_rnn_0"] = (h, h)
    shapes[f"{scope}/Uh_rnn_0"] = (inputdim, h)
    shapes[f"{scope}/Wv_rnn_0"] = (h, h)
    shapes[f"{scope}/Uv_rnn_0"] = (inputdim, h)
    shapes[f"{scope}/b_rnn_0"] = (h,)
    shapes[f"{scope}/wf_dense/kernel"] = (h, 2)
    shapes[f"{scope}/wf_dense/bias"] = (2,)
    return shapes


def count(shapes):
    return int(sum(int(np.prod(s)) for s in shapes.values()))


def _glorot(rng, shape):
    fan_in, fan_out = (shape[0], shape[0]) if len(shape) == 1 else (shape[0], shape[1])
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=shape)


def init_flat(shapes, seed, dtype, mdrnn=False):
    """Reference initialiser distributions: glorot-uniform kernels; GRU gate bias 1, candidate
    biases 0; Dense bias 0; MDRNN: all five cell tensors xavier-uniform including the bias.
    (TF's seed -> stream mapping is not reproducible; only the distribution is.)"""
    rng = np.random.default_rng(seed)
    parts = []
    for name, shape in shapes.items():
        if name.endswith("gates/bias"):
            a = np.ones(shape)
        elif name.endswith("bias"):
            a = np.zeros(shape)
        else:
            a = _glorot(rng, shape)
        parts.append(a.reshape(-1))
    return np.concatenate(parts).astype(dtype)


def split_flat(flat, shapes):
    out = OrderedDict()
    o = 0
    for name, shape in shapes.items():
        n = int(np.prod(shape))
        out[name] = np.asarray(flat[o:o + n]).reshape(shape).copy()
        o += n
    if o != len(flat):
        raise ValueError(f"flat parameter vector has {len(flat)} entries, layout needs {o}")
    return out


def join_named(named, shapes, dtype):
    parts = []
    for name, shape in shapes.items():
        key = name if name in named else name + ":0"
        if key not in named:
            raise KeyError(f"missing variable {name}")
        a = np.asarray(named[key])
        if tuple(a.shape) != tuple(shape):
            raise ValueError(f"{name}: shape {a.shape} != {shape}")
        parts.append(a.reshape(-1))
    return np.concatenate(parts).astype(dtype)


def units_from_named(named, scope=SCOPE):
    """Layer widths of a GRU stack from a `{variable name: array}` dump (names with or without TF's ':0' suffix)."""
    units = []
    while True:
        key = f"{scope}/multi_rnn_cell/cell_{len(units)}/cudnn_compatible_gru_cell/gates/bias"
        a = named.get(key, named.get(key + ":0"))
        if a is None:
            break
        units.append(int(np.asarray(a).shape[0]) // 2)
    if not units:
        raise KeyError(f"no {scope}/multi_rnn_cell/cell_0/... variables in the dump")
    return units
