"""Drop-in for the reference's 2DTFIM_1DRNN/RNNwavefunction.py (1-D RNN over the flattened lattice, float64)."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))

from rnnwavefunctions_b200.wavefunction import RNNwavefunction2DFlat as RNNwavefunction  # noqa: E402,F401
