"""Drop-in for the reference's 1DTFIM/RNNwavefunction.py: same class name and methods, sm_100a kernels underneath."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))

from rnnwavefunctions_b200.wavefunction import RNNwavefunction1D as RNNwavefunction  # noqa: E402,F401
