"""Drop-in for the reference's J1J2/ComplexRNNwavefunction.py (cRNN with U(1) masking)."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))

from rnnwavefunctions_b200.wavefunction import ComplexRNNwavefunction as RNNwavefunction  # noqa: E402,F401
