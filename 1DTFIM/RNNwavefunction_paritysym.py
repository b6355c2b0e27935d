"""Drop-in for the reference's 1DTFIM/RNNwavefunction_paritysym.py (parity-symmetrised log-probability)."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))

from rnnwavefunctions_b200.wavefunction import RNNwavefunctionParity as RNNwavefunction  # noqa: E402,F401
