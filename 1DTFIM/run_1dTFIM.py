"""Same call as the reference's 1DTFIM/run_1dTFIM.py (run from this directory; results under ../Check_Points/1DTFIM)."""
from TrainingRNN_1DTFIM import run_1DTFIM

if __name__ == "__main__":
    RNNEnergy, varRNNEnergy = run_1DTFIM(numsteps=10 ** 3, systemsize=20, Bx=+1, num_units=50, num_layers=1, numsamples=500,
                                         learningrate=5e-3, seed=111)
