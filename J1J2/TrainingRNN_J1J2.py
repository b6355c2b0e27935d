"""Drop-in for the reference's J1J2/TrainingRNN_J1J2.py: J1J2MatrixElements, J1J2Slices, run_J1J2."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))

from rnnwavefunctions_b200.training import J1J2MatrixElements, J1J2Slices, run_J1J2  # noqa: E402,F401
