"""Drop-in for the reference's 2DTFIM_2DRNN/RNNwavefunction.py (2-D RNN on the zig-zag path, float64)."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))

from rnnwavefunctions_b200.wavefunction import RNNwavefunction2D as RNNwavefunction  # noqa: E402,F401
