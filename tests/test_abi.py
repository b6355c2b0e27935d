"""CPU-side checks of the drop-in boundary: librnnwf_b200.so builds for sm_100a, loads without a GPU and
exports every symbol include/rnnwf.h declares; argument validation happens before any CUDA call."""
import ctypes as C
import os
import re

import pytest

from rnnwavefunctions_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build(verbose=False)
    return _lib.load()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "rnnwf.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rnnwf_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound(lib):
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/rnnwf.h but not exported"
        assert n in _lib.EXPORTS, f"{n} has no ctypes signature in _lib.py"
    assert sorted(_lib.EXPORTS) == names


def test_abi_version_and_param_counts(lib):
    assert lib.rnnwf_abi_version() == 1
    # parameter counts pinned by the reference notebooks: 422 (Tutorial_1DTFIM.ipynb#cell15), 444 (Tutorial_1DJ1J2.ipynb#cell15)
    m = _lib.Model(_lib.CELL_GRU, _lib.HEAD_PROB, _lib.F32, 1, 10, 10, 0, 0)
    assert lib.rnnwf_param_count(C.byref(m)) == 422
    m = _lib.Model(_lib.CELL_GRU, _lib.HEAD_COMPLEX, _lib.F32, 1, 10, 10, 0, 0)
    assert lib.rnnwf_param_count(C.byref(m)) == 444
    m = _lib.Model(_lib.CELL_GRU, _lib.HEAD_PROB, _lib.F32, 3, 50, 1000, 0, 0)
    assert lib.rnnwf_param_count(C.byref(m)) == 38502          # cfg2
    m = _lib.Model(_lib.CELL_MDRNN, _lib.HEAD_PROB, _lib.F64, 1, 100, 144, 12, 12)
    assert lib.rnnwf_param_count(C.byref(m)) == 20702          # cfg4


def test_errors_are_reported_not_raised(lib):
    bad = _lib.Model(7, 0, 0, 1, 10, 10, 0, 0)
    assert lib.rnnwf_param_count(C.byref(bad)) == -1
    assert b"unknown cell" in lib.rnnwf_last_error()
    odd = _lib.Model(_lib.CELL_GRU, _lib.HEAD_COMPLEX, _lib.F32, 1, 10, 11, 0, 0)
    assert lib.rnnwf_param_count(C.byref(odd)) == -1           # zero magnetisation needs even N
    m = _lib.Model(_lib.CELL_GRU, _lib.HEAD_PROB, _lib.F32, 1, 10, 10, 0, 0)
    rc = lib.rnnwf_sample(C.byref(m), None, 10, 0, 0, None, None, 0, None)
    assert rc == -1 and b"bad arguments" in lib.rnnwf_last_error()
    with pytest.raises(_lib.RnnwfError):
        _lib.check(rc)


def test_workspace_sizes_scale_with_samples(lib):
    m = _lib.Model(_lib.CELL_GRU, _lib.HEAD_PROB, _lib.F32, 3, 50, 1000, 0, 0)
    a = lib.rnnwf_workspace_bytes(C.byref(m), _lib.OP_TFIM_ELOC, 1200, 0)
    b = lib.rnnwf_workspace_bytes(C.byref(m), _lib.OP_TFIM_ELOC, 2400, 0)
    assert 0 < a < b < 2.2 * a
    # hidden-state stash: N*L*H*4 B = 600 KB per sample
    assert a / 1200 > 600e3
    par = lib.rnnwf_workspace_bytes(C.byref(m), _lib.OP_TFIM_ELOC, 1200, _lib.PARITY_SYM)
    assert 1.9 * a < par < 2.1 * a


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "rnnwavefunctions_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f


def test_lr_schedules_match_the_reference_formulas():
    """constant (1DTFIM/TrainingRNN_1DTFIM.py:221), 1/((1/lr)+it/10) (2DTFIM_1DRNN/Training1DRNN_2DTFIM.py:229),
    lr (1+it/5000)^-1 (2DTFIM_2DRNN/Training2DRNN_2DTFIM.py:228)."""
    import inspect

    import pytest

    from rnnwavefunctions_b200 import training as TR
    lr = 5e-3
    assert TR._schedule(lr, None)(1234) == lr and TR._schedule(lr, "constant")(0) == lr
    inv = TR._schedule(lr, "inverse")
    assert inv(0) == pytest.approx(lr) and inv(70) == pytest.approx(1.0 / (1.0 / lr + 7.0))
    inv5 = TR._schedule(lr, "inverse5000")
    assert inv5(0) == pytest.approx(lr) and inv5(5000) == pytest.approx(lr / 2)
    assert TR._schedule(lr, lambda it: 7.0)(3) == 7.0
    with pytest.raises(ValueError):
        TR._schedule(lr, "cosine")
    # the drivers keep the schedules the reference ships as their defaults
    assert inspect.signature(TR.run_1DTFIM).parameters["lr_schedule"].default is None
    assert inspect.signature(TR.run_2DTFIM_1DRNN).parameters["lr_schedule"].default == "inverse"
    assert inspect.signature(TR.run_2DTFIM_2DRNN).parameters["lr_schedule"].default == "inverse5000"
    assert inspect.signature(TR.run_J1J2).parameters["lr_schedule"].default is None


def test_sample_slices_fit_the_workspace_budget(lib, monkeypatch):
    """Host logic of ops.sample_chunk (no GPU needed: rnnwf_workspace_bytes is host arithmetic): the slice is the largest multiple of
    256 samples whose workspace fits RNNWF_WS_BUDGET_GB; batches that fit are not sliced."""
    from rnnwavefunctions_b200 import ops
    m = ops.make_model(num_layers=3, units=50, n_sites=1000)
    ns = 100_000
    need = ops.workspace_bytes(m, ops.OP_VMC_GRAD, ns)
    assert need > 400e9                                   # ~47 GB per 10^4 samples at cfg2: does not fit one B200
    monkeypatch.setenv("RNNWF_WS_BUDGET_GB", "100")
    step = ops.sample_chunk(m, ops.OP_VMC_GRAD, ns, 0, "cuda:0")
    assert 0 < step < ns and step % 256 == 0
    assert ops.workspace_bytes(m, ops.OP_VMC_GRAD, step) <= 100 * 2 ** 30 < ops.workspace_bytes(m, ops.OP_VMC_GRAD, step + 256)
    assert ops.sample_chunk(m, ops.OP_TFIM_ELOC, 10_000, 0, "cuda:0") == 10_000      # 6.3 GB: fits
    par = ops.sample_chunk(m, ops.OP_TFIM_ELOC, ns, ops.PARITY_SYM, "cuda:0")
    assert par < ops.sample_chunk(m, ops.OP_TFIM_ELOC, ns, 0, "cuda:0")             # both directions of every chain


def test_unequal_layer_widths_embed_into_the_padded_layout():
    """params.gru_pad_index: the real TF-order vector of a stack with unequal widths lands in the top-left blocks of the equal-width
    layout of its widest layer; everything else stays zero (host logic of the zero-padding path, 1DTFIM/RNNwavefunction.py:32)."""
    import numpy as np

    from rnnwavefunctions_b200 import params as P
    for units, heads in (([5, 3, 4], ("wf_dense",)), ([3, 6], ("wf_dense_ampl", "wf_dense_phase")), ([7, 2], ("wf_dense",))):
        H = max(units)
        real_shapes, pad_shapes = P.gru_shapes(units, heads=heads), P.gru_shapes([H] * len(units), heads=heads)
        index, padded_count = P.gru_pad_index(units, heads=heads)
        assert padded_count == P.count(pad_shapes)
        flat = np.arange(1, P.count(real_shapes) + 1, dtype=np.float64)
        padded = np.zeros(padded_count)
        padded[index] = flat
        real, pad = P.split_flat(flat, real_shapes), P.split_flat(padded, pad_shapes)
        d = 2
        for l, h in enumerate(units):
            D = 2 if l == 0 else H
            pre = f"RNNwavefunction/multi_rnn_cell/cell_{l}/cudnn_compatible_gru_cell/"
            Kg, Kp = real[pre + "gates/kernel"], pad[pre + "gates/kernel"]
            assert np.array_equal(Kp[:d, :h], Kg[:d, :h]) and np.array_equal(Kp[:d, H:H + h], Kg[:d, h:])         # input rows: r | u
            assert np.array_equal(Kp[D:D + h, :h], Kg[d:, :h]) and np.array_equal(Kp[D:D + h, H:H + h], Kg[d:, h:])   # hidden rows
            assert np.count_nonzero(Kp) == Kg.size
            bg, bp = real[pre + "gates/bias"], pad[pre + "gates/bias"]
            assert np.array_equal(bp[:h], bg[:h]) and np.array_equal(bp[H:H + h], bg[h:]) and np.count_nonzero(bp) == bg.size
            for name in ("candidate/input_projection/kernel", "candidate/hidden_projection/kernel"):
                a, b = real[pre + name], pad[pre + name]
                assert np.array_equal(b[:a.shape[0], :a.shape[1]], a) and np.count_nonzero(b) == a.size
            for name in ("candidate/input_projection/bias", "candidate/hidden_projection/bias"):
                a, b = real[pre + name], pad[pre + name]
                assert np.array_equal(b[:h], a) and np.count_nonzero(b) == a.size
            d = h
        for head in heads:
            a, b = real[f"RNNwavefunction/{head}/kernel"], pad[f"RNNwavefunction/{head}/kernel"]
            assert np.array_equal(b[:units[-1]], a) and np.count_nonzero(b) == a.size
            assert np.array_equal(pad[f"RNNwavefunction/{head}/bias"], real[f"RNNwavefunction/{head}/bias"])
