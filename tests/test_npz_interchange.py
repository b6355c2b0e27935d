"""Weight interchange keyed by TF 1.13 variable names: the reference prints `v.name` / shapes at start-up
(1DTFIM/TrainingRNN_1DTFIM.py:125-136) and saves with tf.train.Saver (:166, :217-219); a TF user dumps
`{v.name: sess.run(v)}` (scripts/tf113_dump.py) and `load_npz` / `join_named` must consume it: names carry TF's ':0' suffix,
arrive in any order, float32 or float64, one / three GRU layers, the cRNN's two heads, the MDRNN's five tensors.

CPU tests cover the host logic; the GPU tests load such a file into the CUDA wave functions and compare log psi with the oracle
evaluated from the same dictionary."""
import numpy as np
import pytest

from oracle import rnnwf_oracle as O
from rnnwavefunctions_b200 import params as P

CHEADS = ("wf_dense_ampl", "wf_dense_phase")


def tf_style_dump(shapes, seed, dtype, suffix=":0"):
    """What `{v.name: sess.run(v) for v in tf.trainable_variables()}` looks like: ':0' names, shuffled order."""
    rng = np.random.default_rng(seed)
    named = {name + suffix: rng.normal(scale=0.4, size=shape).astype(dtype) for name, shape in shapes.items()}
    keys = list(named)
    rng.shuffle(keys)
    return {k: named[k] for k in keys}


def strip(named):
    return {k[:-2] if k.endswith(":0") else k: v for k, v in named.items()}


@pytest.mark.parametrize("shapes,dtype", [
    (P.gru_shapes([10]), np.float32), (P.gru_shapes([50, 50, 50]), np.float32), (P.gru_shapes([100]), np.float64),
    (P.gru_shapes([10, 10], heads=CHEADS), np.float32), (P.mdrnn_shapes(100), np.float64)])
def test_join_named_accepts_tf_names_in_any_order(shapes, dtype, tmp_path):
    named = tf_style_dump(shapes, 1, dtype)
    path = tmp_path / "dump.npz"
    np.savez(path, samples=np.zeros((2, 3), np.int64), **named)           # extra arrays in the file are ignored
    with np.load(path) as z:
        loaded = {k: z[k] for k in z.files}
    flat = P.join_named(loaded, shapes, dtype)
    assert flat.dtype == dtype and flat.size == P.count(shapes)
    back = P.split_flat(flat, shapes)
    for name in shapes:
        assert np.array_equal(back[name], named[name + ":0"])
    # names without the suffix work too; a float64 dump narrows to the float32 model dtype
    assert np.array_equal(P.join_named(strip(named), shapes, dtype), flat)
    assert P.join_named({k: v.astype(np.float64) for k, v in named.items()}, shapes, np.float32).dtype == np.float32


def test_join_named_rejects_missing_and_misshapen_variables():
    shapes = P.gru_shapes([10])
    named = tf_style_dump(shapes, 2, np.float32)
    bad = dict(named)
    bad.pop("RNNwavefunction/wf_dense/bias:0")
    with pytest.raises(KeyError):
        P.join_named(bad, shapes, np.float32)
    bad = dict(named)
    bad["RNNwavefunction/wf_dense/kernel:0"] = np.zeros((2, 10), np.float32)
    with pytest.raises(ValueError):
        P.join_named(bad, shapes, np.float32)


def test_parameter_names_and_counts_are_the_references():
    # Tutorial_1DTFIM.ipynb#cell15: 422 parameters for 1 x GRU(10); Tutorial_1DJ1J2.ipynb#cell15: 444 for the cRNN
    assert P.count(P.gru_shapes([10])) == 422 and P.count(P.gru_shapes([10], heads=CHEADS)) == 444
    assert list(P.gru_shapes([7, 7])) == list(O.gru_param_shapes([7, 7]))
    assert list(P.gru_shapes([7], heads=CHEADS)) == list(O.gru_param_shapes([7], heads=CHEADS))
    assert list(P.mdrnn_shapes(9)) == list(O.mdrnn_param_shapes(9))
    assert P.units_from_named(tf_style_dump(P.gru_shapes([50, 50, 50]), 0, np.float32)) == [50, 50, 50]
    assert P.units_from_named(strip(tf_style_dump(P.gru_shapes([8], heads=CHEADS), 0, np.float32))) == [8]


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("kind,units,N", [("gru", [10], 12), ("gru", [50, 50, 50], 40), ("parity", [6, 6], 10)])
def test_load_npz_gru_logprob_matches_oracle(kind, units, N, tmp_path):
    import torch
    from rnnwavefunctions_b200 import wavefunction as W
    shapes = P.gru_shapes(units)
    named = tf_style_dump(shapes, 3, np.float32)
    path = str(tmp_path / "tf_dump.npz")
    s = np.random.default_rng(4).integers(0, 2, size=(60, N))
    np.savez(path, samples=s, **named)
    wf = (W.RNNwavefunctionParity if kind == "parity" else W.RNNwavefunction1D)(N, units=W.units_from_named(named), seed=1)
    before = wf.params.clone()
    wf.load_npz(path)
    assert not torch.equal(before, wf.params)
    p = {k: v for k, v in strip(named).items()}
    ref = (O.log_probability_parity if kind == "parity" else O.log_probability)(p, s)
    np.testing.assert_allclose(wf.log_probability(s).cpu().numpy(), ref, rtol=1e-5, atol=1e-6)
    # round trip: save_npz writes the TF names, a second object loads them and is bitwise the same model
    out = str(tmp_path / "saved.npz")
    wf.save_npz(out, samples=s)
    with np.load(out) as z:
        assert set(shapes) <= set(z.files)
    wf2 = type(wf)(N, units=units, seed=99)
    wf2.load_npz(out)
    assert torch.equal(wf2.params, wf.params)


@pytest.mark.gpu
def test_load_npz_f64_flat_crnn_and_mdrnn(tmp_path):
    from rnnwavefunctions_b200 import wavefunction as W
    rng = np.random.default_rng(5)
    # float64 1-D GRU over a 4 x 5 lattice (2DTFIM_1DRNN)
    shapes = P.gru_shapes([12])
    named = tf_style_dump(shapes, 6, np.float64)
    path = str(tmp_path / "flat.npz")
    np.savez(path, **named)
    wf = W.RNNwavefunction2DFlat(4, 5, units=[12])
    wf.load_npz(path)
    s = rng.integers(0, 2, size=(30, 20))
    np.testing.assert_allclose(wf.log_probability(s).cpu().numpy(), O.log_probability(strip(named), s), rtol=1e-11)
    # cRNN: two Dense heads (J1J2/ComplexRNNwavefunction.py:40-43)
    shapes = P.gru_shapes([9, 9], heads=CHEADS)
    named = tf_style_dump(shapes, 7, np.float32)
    path = str(tmp_path / "crnn.npz")
    np.savez(path, **named)
    wf = W.ComplexRNNwavefunction(10, units=[9, 9])
    wf.load_npz(path)
    s = O.crnn_sample(strip(named), 40, 10, seed=1)
    la = wf.log_amplitude(s).cpu().numpy()
    ref = O.crnn_log_amplitude(strip(named), s)
    np.testing.assert_allclose(la.real, ref.real, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(la.imag, ref.imag, rtol=1e-5, atol=1e-5)
    # MDRNN: Wh, Uh, Wv, Uv, b (2DTFIM_2DRNN/MDRNNcell.py:21-35) + Dense
    shapes = P.mdrnn_shapes(11)
    named = {k: v * 0.5 for k, v in tf_style_dump(shapes, 8, np.float64).items()}
    path = str(tmp_path / "md.npz")
    np.savez(path, **named)
    wf = W.RNNwavefunction2D(3, 4, units=[11])
    wf.load_npz(path)
    s = rng.integers(0, 2, size=(25, 3, 4))
    np.testing.assert_allclose(wf.log_probability(s).cpu().numpy(), O.mdrnn_log_probability(strip(named), s), rtol=1e-11)
    # MDRNNcell.call evaluates one cell step from the same tensors (MDRNNcell.py:51-66)
    import torch
    xl, xu = (torch.tensor(rng.normal(size=(5, 2)), device=wf.device) for _ in range(2))
    hl, hu = (torch.tensor(rng.normal(size=(5, 11)), device=wf.device) for _ in range(2))
    out, state = wf.rnn.call((xl, xu), (hl, hu))
    ref = O.mdrnn_cell(strip(named), xl.cpu().numpy(), xu.cpu().numpy(), hl.cpu().numpy(), hu.cpu().numpy())
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=1e-12)
    assert state is out
