"""Data-parallel host logic of the VMC iteration (rnnwavefunctions_b200/vmc.py) on CPU: world_size-2 gloo
run vs a single process over the union of the samples.  The CUDA ops are replaced by the oracle here only
to exercise sharding, the two all-reduces and the replicated Adam update."""
import os
import socket
import types

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

N, UNITS, NS, STEPS = 6, [4], 8, 3


def fake_ops():
    from oracle import rnnwf_oracle as O
    from oracle import torch_grad as TG
    from rnnwavefunctions_b200 import ops as real
    shapes = O.gru_param_shapes(UNITS)

    def P(params):
        return O.unflatten(params.numpy(), shapes, np.float32)

    f = types.SimpleNamespace()
    f.param_count = real.param_count
    f.sample = lambda model, params, ns, seed=0, sample_offset=0: torch.tensor(
        O.sample(P(params), ns, N, seed=seed, sample_offset=sample_offset).astype(np.uint8))

    def tfim_eloc(model, params, s, jz, bx, flags=0, want_logp=True):
        e = O.ising_local_energies(jz.numpy(), bx, s.numpy().astype(np.int64), lambda c: O.log_probability(P(params), c))
        return torch.tensor(e), None
    f.tfim_eloc = tfim_eloc

    def energy_moments(e, stride=1, count=None):
        v = e[::stride] if count is None else e[:count * stride:stride]
        return torch.tensor([v.sum().item(), (v * v).sum().item(), float(v.numel())], dtype=torch.float64)
    f.energy_moments = energy_moments
    f.vmc_grad = lambda model, params, s, w, flags=0: torch.tensor(
        TG.gru_vmc_grad({k: v.astype(np.float64) for k, v in P(params).items()}, s.numpy().astype(np.int64), w.numpy()))

    def adam_step(model, theta, mom, vel, grad, t, lr, gs=1.0, b1=0.9, b2=0.999, eps=1e-8):
        th, m, v, _ = O.adam_tf1(theta.numpy().astype(np.float64), grad.numpy() * gs, mom.numpy().astype(np.float64),
                                 vel.numpy().astype(np.float64), t - 1, lr, b1, b2, eps)
        theta.copy_(torch.tensor(th.astype(np.float32)))
        mom.copy_(torch.tensor(m.astype(np.float32)))
        vel.copy_(torch.tensor(v.astype(np.float32)))
    f.adam_step = adam_step
    return f


def run_vmc(ns):
    import rnnwavefunctions_b200.vmc as V
    from rnnwavefunctions_b200.wavefunction import RNNwavefunction1D
    V.ops = fake_ops()
    wf = RNNwavefunction1D(N, units=UNITS, seed=5, device="cpu")
    opt = V.VMC(wf, V.TFIM(np.ones(N), 1.0), ns)
    hist = []
    for _ in range(STEPS):
        mean, var = opt.step(1e-2)
        hist.append((float(mean), float(var)))
    return wf.params.numpy().copy(), hist


def worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    params, hist = run_vmc(NS)
    q.put((rank, params, hist))
    dist.barrier()
    dist.destroy_process_group()


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(300)
def test_two_ranks_equal_one_rank_over_the_union():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=240) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    single_params, single_hist = run_vmc(2 * NS)
    for rank, params, hist in res:
        np.testing.assert_allclose(hist, single_hist, rtol=1e-10)          # global mean / variance of E_loc
        np.testing.assert_allclose(params, single_params, rtol=0, atol=2e-6)   # replicated, identical update
    assert np.array_equal(res[0][1], res[1][1])                            # replicas stay bitwise identical
