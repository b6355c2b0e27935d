"""ctypes binding of librnnwf_b200.so (the C ABI declared in include/rnnwf.h).

There is deliberately NO fallback: if the CUDA library is missing or fails to load, importing the ops
raises.  Nothing here (or anywhere in the package) touches `oracle/`.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RNNWF_LIB") or os.path.join(HERE, "librnnwf_b200.so")   # RNNWF_LIB: A/B builds while developing

CELL_GRU, CELL_MDRNN = 0, 1
HEAD_PROB, HEAD_COMPLEX = 0, 1
F32, F64 = 0, 1
PARITY_SYM = 1
OP_SAMPLE, OP_LOGPSI, OP_TFIM_ELOC, OP_VMC_GRAD, OP_J1J2_ELOC = range(5)


class Model(C.Structure):
    """struct rnnwf_model (include/rnnwf.h)."""
    _fields_ = [(n, C.c_int32) for n in ("cell", "head", "dtype", "num_layers", "units", "n_sites", "nx", "ny")]


class RnnwfError(RuntimeError):
    pass


_P = C.c_void_p
_SIGNATURES = {
    # name: (restype, argtypes)
    "rnnwf_last_error": (C.c_char_p, []),
    "rnnwf_abi_version": (C.c_int, []),
    "rnnwf_param_count": (C.c_int64, [C.POINTER(Model)]),
    "rnnwf_workspace_bytes": (C.c_size_t, [C.POINTER(Model), C.c_int, C.c_int64, C.c_int]),
    "rnnwf_sample": (C.c_int, [C.POINTER(Model), _P, C.c_int64, C.c_uint64, C.c_uint64, _P, _P, C.c_size_t, _P]),
    "rnnwf_logpsi": (C.c_int, [C.POINTER(Model), _P, _P, C.c_int64, C.c_int, _P, _P, C.c_size_t, _P]),
    "rnnwf_tfim_eloc": (C.c_int, [C.POINTER(Model), _P, _P, C.c_int64, _P, C.c_double, C.c_int, _P, _P, _P, C.c_size_t, _P]),
    "rnnwf_tfim_flip_ratios": (C.c_int, [C.POINTER(Model), _P, _P, C.c_int64, _P, C.c_double, C.c_int, _P, _P, _P, _P, C.c_size_t, _P]),
    "rnnwf_tfim_chain_mode": (C.c_int, [C.POINTER(Model)]),
    "rnnwf_tfim_diag": (C.c_int, [C.POINTER(Model), _P, C.c_int64, _P, _P, _P]),
    "rnnwf_tfim_enumerate": (C.c_int, [_P, C.c_int64, C.c_int32, _P, _P]),
    "rnnwf_j1j2_enumerate": (C.c_int, [_P, C.c_int64, C.c_int32, _P, _P, _P, C.c_int, C.c_int, _P, _P, _P, _P]),
    "rnnwf_j1j2_eloc": (C.c_int, [C.POINTER(Model), _P, _P, C.c_int64, _P, _P, _P, C.c_int, _P, _P, _P, C.c_size_t, _P]),
    "rnnwf_vmc_grad": (C.c_int, [C.POINTER(Model), _P, _P, C.c_int64, _P, C.c_int, _P, _P, C.c_size_t, _P]),
    "rnnwf_adam_step": (C.c_int, [C.c_int, C.c_int64, _P, _P, _P, _P, C.c_double, C.c_double, C.c_double, C.c_double,
                                  C.c_double, C.c_int64, _P]),
    "rnnwf_energy_moments": (C.c_int, [_P, C.c_int64, C.c_int, _P, _P]),
    "rnnwf_profile_begin": (C.c_int, []),
    "rnnwf_profile_end": (C.c_int, [C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_double)]),
    "rnnwf_ffma_peak": (C.c_int, [C.c_int, C.POINTER(C.c_double), _P]),
    "rnnwf_fp64_peak": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_double), _P]),
    "rnnwf_umma_selftest": (C.c_int, [C.c_int, C.c_int, _P, _P, _P, C.c_int, _P]),
}
EXPORTS = tuple(_SIGNATURES)

_lib = None


def load():
    """Load the shared library (building is the job of __graft_entry__.build / rnnwavefunctions_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RnnwfError(
            f"{LIB_PATH} is missing: build it with `python -m rnnwavefunctions_b200.build` (needs nvcc). "
            "There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.rnnwf_abi_version() != 1:
        raise RnnwfError("ABI version mismatch")
    _lib = lib
    return lib


def check(code: int):
    if code != 0:
        msg = load().rnnwf_last_error()
        raise RnnwfError(f"librnnwf_b200 error {code}: {msg.decode() if msg else ''}")
