"""Drop-in for the reference's 2DTFIM_2DRNN/MDRNNcell.py."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))

from rnnwavefunctions_b200.wavefunction import MDRNNcell  # noqa: E402,F401
