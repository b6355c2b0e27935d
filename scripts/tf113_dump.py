"""Dump the trainable variables, a batch of samples and their log-probabilities of a TF 1.13 run of the reference, keyed by
TF variable name, for the external parity check of SURVEY.md 8(c).

NOT RUNNABLE IN THIS REPOSITORY'S CONTAINER: it needs TensorFlow 1.13.1 (Python <= 3.7) and the reference checkout.  Run it
next to the reference's 1DTFIM/TrainingRNN_1DTFIM.py (after the `sess.run(init)` of :168, or after training), then check the file
here with

    python scripts/check_tf_dump.py dump.npz          # needs a B200: loads the weights, compares log-probabilities (1e-5)

The variable names are what the reference itself prints at start-up (1DTFIM/TrainingRNN_1DTFIM.py:125-136), e.g.
    RNNwavefunction/multi_rnn_cell/cell_0/cudnn_compatible_gru_cell/gates/kernel:0
`rnnwavefunctions_b200.params.join_named` accepts them with or without the ':0' suffix, in any order.
"""


def dump(path, sess, wf, samples_placeholder, log_probs_tensor, numsamples=64):
    import numpy as np
    import tensorflow as tf            # 1.13.1

    with wf.graph.as_default():
        variables = {v.name: sess.run(v) for v in tf.trainable_variables()}
        samples = sess.run(wf.sample(numsamples=numsamples, inputdim=2))
        log_probs = sess.run(log_probs_tensor, feed_dict={samples_placeholder: samples})
    np.savez(path, samples=samples, log_probs=log_probs, **variables)
