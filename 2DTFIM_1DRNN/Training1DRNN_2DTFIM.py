"""Drop-in for the reference's 2DTFIM_1DRNN/Training1DRNN_2DTFIM.py: Ising2D_local_energies and run_2DTFIM."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))

from rnnwavefunctions_b200.training import Ising2D_local_energies  # noqa: E402,F401
from rnnwavefunctions_b200.training import run_2DTFIM_1DRNN as run_2DTFIM  # noqa: E402,F401
