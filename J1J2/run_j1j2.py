"""Same call as the reference's J1J2/run_j1j2.py."""
from TrainingRNN_J1J2 import run_J1J2

if __name__ == "__main__":
    RNNEnergy, varRNNEnergy = run_J1J2(numsteps=3000, systemsize=10, J1_=1.0, J2_=0.2, Marshall_sign=False, num_units=10,
                                       num_layers=1, numsamples=200, learningrate=5e-4, seed=111)
