"""Drop-in for the reference's 1DTFIM/TrainingRNN_1DTFIM.py: Ising_local_energies and run_1DTFIM."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))

from rnnwavefunctions_b200.training import Ising_local_energies, run_1DTFIM  # noqa: E402,F401
