#!/usr/bin/env python
"""Headline benchmark: one full VMC iteration (sample + local energies + gradient + all-reduce + TF1 Adam) of a BASELINE.json
configuration.  Default: configs[1] (1-D TFIM N=1000, 3 x GRU(50), 10^4 samples per GPU, Bx=1), the configuration the metric
is quoted on; `--config` selects the others (they are parity-test cases; their lines go to profiles/).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg1|cfg2|cfg2p|cfg3|cfg4|cfg5] [--impl reference]

Own arm: `value` = samples/s through the whole iteration, aggregate over ranks, device-timed with CUDA events, max over ranks.
`e2e` = the same iteration driven through the reference-facing host API (wf.sample -> NumPy; *_local_energies(host samples) ->
NumPy; optimiser step fed from host arrays), copies included.  `roofline` = the dominant kernel (the prefix-reuse chain kernel of
the local-energy stage), timed live with CUDA events recorded inside the library around its launch, against the roofline that
bounds it: the dense 16-bit tensor peak of MEASURED_PEAKS.json for the tcgen05 3xFP16 kernel (cfg1, cfg2, cfg2p, cfg5), the FP64
peak measured in the same run (rnnwf_fp64_peak) for the float64 models (cfg3, cfg4).  HBM is never the bound (weights live in shared
memory).  `cpu_baseline` = the NumPy restatement of the reference algorithm (oracle/) on the host cores, bounded sample.
Reference arm (--impl reference): the same oracle timed as the reference's CPU implementation.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

if "reference" in sys.argv:      # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every host thread
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FP32_PEAK_THEORETICAL_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12                                              # 74.4
CHEADS = ("wf_dense_ampl", "wf_dense_phase")

# BASELINE.json configs (sizes, parameter counts P and flops per cell-stack F).  `ns`: samples per GPU (cfg1: the
# shipped run script's 500; the others 10^4 as cfg2, the weak-scaling unit of SURVEY.md 8e).
CONFIGS = {
    "cfg1": dict(kind="tfim1d", N=20, layers=1, units=50, ns=500, dtype="f32", bx=1.0, lr=5e-3, parity=False,
                 label="1D TFIM N=20 Bx=1 OBC, pRNN 1x GRU(50)"),
    "cfg2": dict(kind="tfim1d", N=1000, layers=3, units=50, ns=10_000, dtype="f32", bx=1.0, lr=5e-3, parity=False,
                 label="1D TFIM N=1000 Bx=1 OBC, pRNN 3x GRU(50)"),
    "cfg2p": dict(kind="tfim1d", N=1000, layers=3, units=50, ns=10_000, dtype="f32", bx=1.0, lr=5e-3, parity=True,
                  label="1D TFIM N=1000 Bx=1 OBC, pRNN 3x GRU(50)"),
    "cfg3": dict(kind="tfim2d_flat", nx=12, ny=12, N=144, layers=1, units=100, ns=10_000, dtype="f64", bx=3.0, lr=1e-3, parity=False,
                 label="2D TFIM 12x12 Bx=3 OBC, 1D pRNN (row-major order) 1x GRU(100) float64"),
    "cfg4": dict(kind="tfim2d_mdrnn", nx=12, ny=12, N=144, layers=1, units=100, ns=10_000, dtype="f64", bx=3.0, lr=5e-3, parity=False,
                 label="2D TFIM 12x12 Bx=3 OBC, 2D RNN MDRNNcell(100) float64, zig-zag path"),
    "cfg5": dict(kind="j1j2", N=100, layers=1, units=50, ns=10_000, dtype="f32", j2=0.2, marshall=True, lr=2.5e-4, parity=False,
                 label="1D J1-J2 N=100 J2=0.2 OBC, cRNN 1x GRU(50), U(1) zero magnetisation, Marshall sign"),
}
METRICS = {
    "tfim1d": "local-energy samples/s through one full VMC step (sample + E_loc + gradient + all-reduce + Adam), 1D TFIM N={N} GRU",
    "tfim2d_flat": "local-energy samples/s through one full VMC step (sample + E_loc + gradient + all-reduce + Adam), 2D TFIM 12x12 1D-RNN GRU(100) f64",
    "tfim2d_mdrnn": "local-energy samples/s through one full VMC step (sample + E_loc + gradient + all-reduce + Adam), 2D TFIM 12x12 MDRNN(100) f64",
    "j1j2": "local-energy samples/s through one full VMC step (sample + E_loc + gradient + all-reduce + Adam), 1D J1-J2 N=100 cRNN",
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--parity", action="store_true", help="same as --config cfg2p (parity-symmetric wave function, RNNwavefunction_paritysym)")
    ap.add_argument("--ns", type=int, default=None, help="samples per GPU (default: the config's)")
    ap.add_argument("--n-sites", type=int, default=None, help="chain length override (1-D configs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="target CPU seconds per reference-arm step")
    a = ap.parse_args()
    if a.parity:
        a.config = "cfg2p"
    c = dict(CONFIGS[a.config])
    if a.ns is not None:
        c["ns"] = a.ns
    if a.n_sites is not None and c["kind"] in ("tfim1d", "j1j2"):
        c["N"] = a.n_sites
    c["name"] = a.config
    a.cfg = c
    return a


def flop_per_cellstack(c):
    """F of SURVEY.md 8d: 2 * [sum over layers (d + h) * 3h + 2h] for the GRU stacks (the complex head adds 2h), 2 * [(2+h)*2h + 2h] for the MDRNN."""
    h, L = c["units"], c["layers"]
    if c["kind"] == "tfim2d_mdrnn":
        return 2 * ((2 + h) * 2 * h + 2 * h)
    head = 4 * h if c["kind"] == "j1j2" else 2 * h
    return 2 * ((2 + h) * 3 * h + (L - 1) * (2 * h) * 3 * h + head)


def param_count(c):
    h, L = c["units"], c["layers"]
    if c["kind"] == "tfim2d_mdrnn":
        return 2 * h * h + 4 * h + h + 2 * h + 2
    gru = lambda d: (d + h) * 2 * h + 2 * h + d * h + h + h * h + h
    return gru(2) + (L - 1) * gru(h) + (2 if c["kind"] == "j1j2" else 1) * (2 * h + 2)


def workload(c):
    return {"workload": f"{c['label']}, {c['ns']} samples per GPU" + (", parity-symmetric" if c["parity"] else ""),
            "baseline_config": c["name"], "n_sites": c["N"], "layers": c["layers"], "units": c["units"], "samples_per_gpu": c["ns"],
            "parity": bool(c["parity"]), "params": param_count(c)}


def metric_of(c):
    return METRICS[c["kind"]].format(N=c["N"])


# ------------------------------------------------------------------------------------------------
# CPU arm: the NumPy restatement of the reference algorithm (full recompute of all connected configurations
# per sample in the reference's chunks, model dtype, f64 log-sum; autograd gradient; TF1 Adam)
# ------------------------------------------------------------------------------------------------
class CpuReference:
    def __init__(self, c):
        from oracle import rnnwf_oracle as O
        from oracle import torch_grad as TG
        self.O, self.TG, self.c = O, TG, c
        try:
            import torch
            torch.set_num_threads(os.cpu_count() or 1)
        except Exception:
            pass
        k, h = c["kind"], c["units"]
        self.np_dtype = np.float32 if c["dtype"] == "f32" else np.float64
        self.units = [h] * c["layers"]
        if k == "tfim2d_mdrnn":
            self.p = O.init_mdrnn_params(h, seed=111, dtype=np.float64, scale=0.5)
            self.shapes = O.mdrnn_param_shapes(h)
        else:
            heads = CHEADS if k == "j1j2" else ("wf_dense",)
            self.p = O.init_gru_params(self.units, seed=111, dtype=self.np_dtype, heads=heads)
            self.shapes = O.gru_param_shapes(self.units, heads=heads)
        self.m = np.zeros(O.num_params(self.shapes))
        self.v = np.zeros_like(self.m)
        self.t = 0
        self.it = 0

    def step(self, ns):
        O, TG, c = self.O, self.TG, self.c
        k, N = c["kind"], c["N"]
        p64 = {key: v.astype(np.float64) for key, v in self.p.items()}
        seed = 111 + self.it
        self.it += 1
        if k == "tfim1d":
            s = O.sample(self.p, ns, N, seed=seed)
            lp = (lambda q: O.log_probability_parity(self.p, q)) if c["parity"] else (lambda q: O.log_probability(self.p, q))
            e = O.ising_local_energies(np.ones(N), c["bx"], s, lp)
            g = TG.gru_vmc_grad(p64, s, (e - e.mean()) / ns, parity=c["parity"])
        elif k == "tfim2d_flat":
            s = O.sample(self.p, ns, N, seed=seed)
            e = O.ising2d_local_energies(np.ones((c["nx"], c["ny"])), c["bx"], c["nx"], c["ny"], s, lambda q: O.log_probability(self.p, q), flat=True)
            g = TG.gru_vmc_grad(p64, s, (e - e.mean()) / ns)
        elif k == "tfim2d_mdrnn":
            s = O.mdrnn_sample(self.p, ns, c["nx"], c["ny"], seed=seed)
            e = O.ising2d_local_energies(np.ones((c["nx"], c["ny"])), c["bx"], c["nx"], c["ny"], s, lambda q: O.mdrnn_log_probability(self.p, q), flat=False)
            g = TG.mdrnn_vmc_grad(p64, s, (e - e.mean()) / ns)
        else:
            s = O.crnn_sample(self.p, ns, N, seed=seed)
            e = O.j1j2_local_energies(np.ones(N), c["j2"] * np.ones(N), np.zeros(N), s, lambda q: O.crnn_log_amplitude(self.p, q),
                                      marshall_sign=c["marshall"])
            g = TG.crnn_vmc_grad(p64, s, 2.0 * (e - e.mean()) / ns)
        theta = O.flatten(self.p).astype(np.float64)
        theta, self.m, self.v, self.t = O.adam_tf1(theta, g, self.m, self.v, self.t, c["lr"])
        self.p = O.unflatten(theta.astype(self.np_dtype), self.shapes, self.np_dtype)
        return float(np.real(e).mean())

    def calibrate(self, target_s):
        """Pick a sample count whose step takes about target_s (work is linear in the sample count)."""
        n0 = 1 if self.c["N"] >= 500 else 8
        t0 = time.perf_counter()
        self.step(n0)
        t1 = (time.perf_counter() - t0) / n0
        return max(1, min(self.c["ns"], 4096, int(target_s / max(t1, 1e-4))))


def cpu_threads():
    try:
        import torch
        return int(torch.get_num_threads())
    except Exception:
        return os.cpu_count() or 1


def cpu_sample_note(c, ns, extra=""):
    what = {"tfim1d": f"full recompute of ({c['N']}+1) configurations per sample", "tfim2d_flat": "full recompute of 145 configurations per sample",
            "tfim2d_mdrnn": "full recompute of 145 configurations per sample",
            "j1j2": "full recompute of every connected configuration (diagonal + antiparallel NN / NNN exchanges) per sample"}[c["kind"]]
    return f"{ns} samples per step{extra} of the same workload: {what}, as the reference does"


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    c = args.cfg
    ref = CpuReference(c)
    ns = ref.calibrate(args.cpu_seconds)
    for _ in range(max(0, args.warmup - 1)):      # calibrate() already ran one untimed step
        ref.step(ns)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref.step(ns)
    dt = time.perf_counter() - t0
    val = ns * args.steps / dt
    line = {"impl": "reference", "metric": metric_of(c), "value": val, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": c["dtype"], "data": "synthetic", "config": dict(workload(c), samples_timed=ns),
            "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cpu_threads(), "kind": "port",
                             "sample": cpu_sample_note(c, ns, f" x {args.steps} steps")},
            "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "NumPy/torch-CPU restatement of the reference algorithm (oracle/); TensorFlow 1.13 is not installable here. The rate is "
                    "linear in the sample count: samples_timed per step were timed, not the config's samples_per_gpu"}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                       "-i", str(index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return None
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().splitlines():
            c = [x.strip() for x in ln.split(",")]
            if len(c) < 7:
                continue
            try:
                sm.append(float(c[0])); mx.append(float(c[1])); pw.append(float(c[2]))
            except ValueError:
                continue
            for nm, val in zip(names, c[3:7]):
                if val == "Active":
                    reasons.add(nm)
        os.unlink(self.f.name)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# own arm
# ------------------------------------------------------------------------------------------------
def build_problem(c, dev):
    from rnnwavefunctions_b200.vmc import J1J2, TFIM
    from rnnwavefunctions_b200.wavefunction import (ComplexRNNwavefunction, RNNwavefunction1D, RNNwavefunction2D, RNNwavefunction2DFlat,
                                                    RNNwavefunctionParity)
    k, N, units = c["kind"], c["N"], [c["units"]] * c["layers"]
    if k == "tfim1d":
        wf = (RNNwavefunctionParity if c["parity"] else RNNwavefunction1D)(N, units=units, seed=111, device=dev)
        return wf, TFIM(np.ones(N), c["bx"])
    if k == "tfim2d_flat":
        return RNNwavefunction2DFlat(c["nx"], c["ny"], units=units, seed=111, device=dev), TFIM(np.ones((c["nx"], c["ny"])), c["bx"])
    if k == "tfim2d_mdrnn":
        return RNNwavefunction2D(c["nx"], c["ny"], units=units, seed=111, device=dev), TFIM(np.ones((c["nx"], c["ny"])), c["bx"])
    return ComplexRNNwavefunction(N, units=units, seed=111, device=dev), J1J2(np.ones(N), c["j2"] * np.ones(N), np.zeros(N), c["marshall"])


def chain_cellstacks_per_sample(c, samples_u8):
    """Algorithmic cell-stack evaluations of the flip / exchange chains per sample (prefix reuse: a chain restarts after the first
    modified site s and runs sites s+1 .. N-1).  TFIM: N(N-1)/2 exactly; J1-J2: data dependent -- only antiparallel NN / NNN pairs
    have a connected configuration (J1J2/TrainingRNN_J1J2.py:68-92) -- counted from the samples."""
    N = c["N"]
    if c["kind"] != "j1j2":
        return N * (N - 1) / 2.0
    import torch
    s = samples_u8.to(torch.int64)
    w = torch.arange(N - 1, 0, -1, device=s.device, dtype=torch.float64)          # chain length N-1-s for s = 0 .. N-2
    nn = (s[:, :-1] != s[:, 1:]).to(torch.float64) @ w
    nnn = (s[:, :-2] != s[:, 2:]).to(torch.float64) @ w[:-1] if c["j2"] != 0 else 0.0
    return float((nn + nnn).mean().item())


def run_b200_arm(args):
    import torch
    import torch.distributed as dist

    from rnnwavefunctions_b200 import ops
    from rnnwavefunctions_b200 import training as TR
    from rnnwavefunctions_b200.vmc import VMC

    c = args.cfg
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N, ns, LR = c["N"], c["ns"], c["lr"]
    wf, H = build_problem(c, dev)
    opt = VMC(wf, H, ns)
    ndir = 2 if c["parity"] else 1
    F = flop_per_cellstack(c)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ev = lambda: torch.cuda.Event(enable_timing=True)
    for _ in range(args.warmup):
        opt.step(LR)
    barrier()

    # ---- timed region: K full iterations, stage events on the launching stream ----------------------
    clocks = ClockSampler(local) if rank == 0 else None
    ops.profile_begin()
    marks = [[ev() for _ in range(5)] for _ in range(args.steps)]
    e0, e1 = ev(), ev()
    barrier()
    e0.record()
    means, chain_stacks = [], []
    for k in range(args.steps):
        marks[k][0].record()
        s = opt.draw()
        marks[k][1].record()
        e = opt.local_energies(s)
        marks[k][2].record()
        mean, var, n = opt.moments(e)           # all-reduce of [sum E, sum E^2, n]: the first sync point between ranks
        marks[k][3].record()
        g = opt.gradient(s, e, mean, n)
        opt.apply(g, LR)
        marks[k][4].record()
        means.append(mean)
        if k == args.steps - 1:
            last_samples = s
    e1.record()
    barrier()
    launches, dom_n, dom_ms = ops.profile_end()
    clk = clocks.stop() if clocks else None
    t_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    t_ms = float(t_ms.item())
    st4 = np.array([[m[i].elapsed_time(m[i + 1]) for i in range(4)] for m in marks]).mean(axis=0)
    stage = np.array([st4[0], st4[1], st4[2] + st4[3]])
    value = world * ns * args.steps / (t_ms * 1e-3)
    # per-rank attribution of the 1 -> N gap: chain-kernel time and the time spent in the moments all-reduce (waiting for the slowest rank)
    per_rank = None
    if world > 1:
        mine = torch.tensor([dom_ms / max(dom_n, 1), st4[1], st4[2], st4[3]], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = [{"rank": r, "chain_ms": float(v[0]), "local_energies_ms": float(v[1]), "sync_wait_ms": float(v[2]),
                     "gradient+adam_ms": float(v[3])} for r, v in enumerate(allr)]

    # ---- roofline of the dominant kernel (prefix-reuse chain kernel), timed live above ---------------
    stacks = chain_cellstacks_per_sample(c, last_samples)
    chain_flops = ndir * ns * stacks * F                                         # algorithmic flops per launch
    chain_ms = dom_ms / max(dom_n, 1)
    achieved = chain_flops / (chain_ms * 1e-3) / 1e12
    step_flops = ns * F * (N + ndir * (stacks + N + 3 * N))                      # sample + E_loc (base pass + chains) + gradient
    mode = ops.tfim_chain_mode(wf.model) if c["dtype"] == "f32" and c["kind"] in ("tfim1d", "j1j2") else 0
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    # DRAM traffic of the dominant kernel per launch: one recorded ncu capture of this workload (profiles/chain_kernel_traffic.json)
    traffic, traffic_note = None, "no ncu capture recorded for this workload"
    try:
        with open(os.path.join(ROOT, "profiles", "chain_kernel_traffic.json")) as f:
            tj = json.load(f)
        wl = tj["workload"]
        if c["kind"] == "tfim1d" and mode == 3 and \
                (wl["n_sites"], wl["layers"], wl["units"], wl["samples_per_gpu"], bool(wl["parity"])) == (N, c["layers"], c["units"], ns, bool(c["parity"])):
            traffic = int(tj["dram_bytes_read"]) + int(tj["dram_bytes_write"])
            traffic_note = "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum), recorded ncu capture: " + tj["how"]
    except Exception:
        pass
    common = {"achieved": achieved, "unit": "TFLOP/s", "kernel_ms": chain_ms, "kernel_share_of_step": dom_ms / t_ms, "launches_timed": dom_n,
              "flops_per_launch": chain_flops, "flop_per_cellstack": F, "chain_cellstacks_per_sample": stacks, "traffic": traffic,
              "traffic_note": traffic_note, "step_algorithmic_tflops": step_flops * args.steps / (t_ms * 1e-3) / 1e12}
    if c["dtype"] == "f64":
        dfma, dmma = ops.fp64_peak(0, 4000), ops.fp64_peak(1, 4000)
        kname = ("md_chain_kernel<double> (2-D RNN zig-zag path, prefix reuse)" if c["kind"] == "tfim2d_mdrnn"
                 else "gru_chain_kernel<double> (FP64 tile engine, prefix reuse)")
        peak = max(dfma, dmma)
        roofline = dict(common, bound="fp64", kernel=kname, peak=peak, frac=achieved / peak, fp64_dfma_peak_measured=dfma,
                        fp64_dmma_peak_measured=dmma,
                        peak_kind="FP64 peak measured in this run (rnnwf_fp64_peak: DFMA and mma.sync.m8n8k4.f64, the larger one); "
                                  "MEASURED_PEAKS.json holds only HBM and bf16-tensor peaks, neither bounds a float64 recurrence whose weights "
                                  "live in shared memory")
    elif mode == 0:
        ffma_meas = ops.ffma_peak(20000)
        roofline = dict(common, bound="fp32-ffma", kernel="gru_chain_kernel<float> (CUDA-core FFMA tile engine)", peak=ffma_meas,
                        frac=achieved / ffma_meas, fp32_ffma_peak_theoretical=FP32_PEAK_THEORETICAL_TFLOPS,
                        peak_kind="FP32 FFMA peak measured in this run (rnnwf_ffma_peak); MEASURED_PEAKS.json holds only HBM and bf16-tensor "
                                  "peaks, neither bounds a CUDA-core FP32 kernel")
    else:
        ffma_meas = ops.ffma_peak(20000)
        # executed tensor-pipe flops: split-operand passes over padded tiles: per (site, 128-row tile, layer) MMAs of
        # 2 * 128 * N * 16 flops each
        L = c["layers"]
        M_old = 128 if mode == 3 else 120
        tiles128 = -(-(ndir * (-(-ns // M_old)) * M_old) // 128)
        if mode == 3:
            # dense K packing (3e): x group 10 x N160 (layer 0: 2 one-hot MMAs), h group 112 + 64 + 9 x N160
            per_l0 = 2 * 128 * 16 * (2 * 160 + 176 + 9 * 160)
            per_l1 = 2 * 128 * 16 * (10 * 160 + 176 + 9 * 160)
            kname = ("tc16p::chain_kernel<false,%s> (tcgen05 kind::f16, 3xFP16 split operands packed densely along K (10 MMAs of K=16 per "
                     "operand group), weights resident in shared memory, merged N=160 gate-block MMAs, MMA and gate math software-pipelined "
                     "over anti-diagonals of the (site, layer) grid)" % ("true" if c["kind"] == "j1j2" else "false"))
        elif mode == 2:
            per_l0 = 2 * 128 * 16 * (2 * 160 + 64 + 12 * 160)
            per_l1 = 2 * 128 * 16 * (12 * 160 + 64 + 12 * 160)
            kname = "tc16::chain_kernel<50,false> (tcgen05 kind::f16, 3xFP16 operands, weights resident in shared memory)"
        else:
            per_l0 = 2 * 128 * 8 * (4 * 192 + 21 * (128 + 64))
            per_l1 = 2 * 128 * 8 * (21 * 192 + 21 * (128 + 64))
            kname = "gru_chain_tc_kernel<50,false> (tcgen05 kind::tf32, 3xTF32 operands)"
        executed = tiles128 * stacks * (per_l0 + (L - 1) * per_l1)
        peak = float(peaks.get("bf16_tflops_sustained", 1399.0)) / (1.0 if mode >= 2 else 2.0)
        roofline = dict(common, bound="tensor", kernel=kname, peak=peak, frac=achieved / peak,
                        fp32_ffma_peak_measured=ffma_meas, fp32_ffma_peak_theoretical=FP32_PEAK_THEORETICAL_TFLOPS,
                        achieved_over_fp32_ffma_peak=achieved / ffma_meas,
                        peak_kind=("dense 16-bit tensor peak, sustained figure of MEASURED_PEAKS.json (kernel timed inside a multi-second step)"
                                   if peaks else "fallback: 1.4 PFLOP/s sustained bf16 (B200_PROFILING.md)") +
                                  ("" if mode >= 2 else "; kind::tf32 runs at half the 16-bit rate"),
                        peak_over_passes=peak / 3.0, frac_of_peak_over_passes=achieved / (peak / 3.0),
                        tf32x3_roofline=(peak / 2.0 if mode >= 2 else peak) / 3.0,
                        tensor_pipe_tflops_executed=executed / (chain_ms * 1e-3) / 1e12,
                        tensor_pipe_frac_executed=executed / (chain_ms * 1e-3) / 1e12 / peak,
                        executed_over_algorithmic=executed / chain_flops,
                        note_rooflines="peak_over_passes = tensor peak / 3: the ceiling of any 3-pass split-operand scheme with FP32-grade accuracy on "
                                       "this pipe; tf32x3_roofline = (16-bit peak / 2) / 3: the ceiling of the 3xTF32 scheme BASELINE.json's north_star names",
                        note="achieved counts ALGORITHMIC flops (F per GRU-stack evaluation); FP32-grade accuracy costs three split-operand products "
                             "(hi*hi + hi*lo + lo*hi); packed densely along K they take 10 MMAs of K = 16 per operand group (152 of 160 K slots "
                             "used) over N = 160 (150 used): the tensor pipe executes ~3.6x the algorithmic flops.  The board runs this kernel at "
                             "its power cap (see clocks).  Ablation builds on one box (profiles/r2/ab_power_ablation.log): the gate math alone "
                             "takes the same cycle count as the whole kernel (the MMAs are hidden in cycles), the tensor pipe alone needs 64 % "
                             "of it, and the cap costs ~10 % of the clock -- the tensor pipe's energy, i.e. the executed MMA count, sets the time")

    # ---- end to end through the reference-facing host API (host buffers, copies inside the timed region) ----
    e2e = None
    if not args.no_e2e:
        k_e2e = max(1, min(args.steps, 2))
        h2d = d2h = 0
        from rnnwavefunctions_b200.wavefunction import Session
        sess = Session()
        kind = c["kind"]

        def host_step():
            nonlocal h2d, d2h
            samples = sess.run(wf.sample(ns, 2))                             # sess.run(samples_) -> NumPy (int64), as TrainingRNN_1DTFIM.py:203
            d2h += ns * N                                                     # one byte per site crosses PCIe; widened to int64 on the host
            if kind == "tfim1d":
                eloc = TR.Ising_local_energies(np.ones(N), c["bx"], samples, None, wf, None, None, None)     # host in, host out
            elif kind in ("tfim2d_flat", "tfim2d_mdrnn"):
                eloc = TR.Ising2D_local_energies(np.ones((c["nx"], c["ny"])), c["bx"], c["nx"], c["ny"], samples, None, wf, None, None, None)
            else:   # J1J2/TrainingRNN_J1J2.py:255-279: the fused device path replaces J1J2Slices + chunked log_amplitude + combine
                eloc = H.local_energies(wf, ops.as_u8_samples(samples, dev, N)).cpu().numpy()
            h2d += ns * N
            d2h += eloc.nbytes
            su8 = ops.as_u8_samples(samples, dev, N)                          # feed_dict {samp: samples, Eloc: local_energies}
            el = torch.as_tensor(eloc).to(dev)
            h2d += ns * N + eloc.nbytes
            mean, var = opt.step_from(su8, el, LR)
            m = complex(mean.item())
            d2h += 16
            return m

        host_step()
        h2d = d2h = 0
        barrier()
        a, b = ev(), ev()
        a.record()
        for _ in range(k_e2e):
            host_step()
        b.record()
        barrier()
        t2 = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        e2e = {"value": world * ns * k_e2e / (float(t2.item()) * 1e-3), "unit": "samples/s", "steps": k_e2e,
               "h2d_bytes_per_step": h2d // k_e2e, "d2h_bytes_per_step": d2h // k_e2e,
               "api": "sess.run(RNNwavefunction.sample(...)) -> NumPy int64; local energies from host samples -> NumPy; optimiser step fed from host arrays"}

    # ---- CPU baseline (rank 0, N=1 only) -------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ref = CpuReference(c)
        ns_cpu = ref.calibrate(args.cpu_seconds)
        t0 = time.perf_counter()
        ref.step(ns_cpu)
        dt = time.perf_counter() - t0
        cpu = {"value": ns_cpu / dt, "unit": "samples/s", "cores": cpu_threads(), "kind": "port",
               "sample": cpu_sample_note(c, ns_cpu, f" ({dt:.1f} s)") + "; work is linear in the sample count"}

    if rank == 0:
        l2 = ("per-step working set (6 GB hidden-state stash per GPU) exceeds L2; no flush needed" if c["kind"] == "tfim1d" and N >= 500 else
              "the workload is compute-bound with weights in shared memory and a per-step working set of restart states > L2 at 10^4 samples; "
              "no flush between iterations")
        line = {"metric": metric_of(c), "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": t_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": c["dtype"],
                "data": "synthetic", "config": dict(workload(c), l2=l2, parallelism=f"dp{world}"),
                "vmc_steps_per_s": args.steps / (t_ms * 1e-3), "eloc_samples_per_s": world * ns / (stage[1] * 1e-3),
                "stages_ms": {"sample": stage[0], "local_energies": stage[1], "moments+gradient+adam": stage[2]},
                "mean_energy_last_step": complex(means[-1].item()).real,
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clk}
        if per_rank is not None:
            line["per_rank"] = per_rank
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_b200_arm(a)
