#!/usr/bin/env python
"""Headline benchmark: one full VMC iteration of the 1-D TFIM pRNN at BASELINE.json's configs[1]
(N=1000, 3 x GRU(50), 10^4 samples per GPU, Bx=1) -- sample + local energies (all N single-flip
configurations per sample) + VMC gradient + all-reduce + TF1 Adam.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--parity] [--impl reference]

Own arm: `value` = samples/s through the whole iteration, aggregate over ranks, device-timed with CUDA events,
max over ranks.  `e2e` = the same iteration driven through the reference-facing host API (wf.sample ->
NumPy; Ising_local_energies(host samples) -> NumPy; optimiser step fed from host arrays), copies included.
`roofline` = the prefix-reuse chain kernel against the FP32 FFMA roofline (this path is CUDA-core compute
bound: weights live in shared memory, HBM traffic is negligible).  `cpu_baseline` = the NumPy
restatement of the reference algorithm (oracle/) on the host cores, bounded sample.
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

N_SITES, LAYERS, UNITS, NS_PER_GPU, BX, LR = 1000, 3, 50, 10_000, 1.0, 5e-3
FLOP_PER_CELLSTACK = 2 * ((2 + UNITS) * 3 * UNITS + (LAYERS - 1) * (2 * UNITS) * 3 * UNITS + 2 * UNITS)   # 75 800
FP32_PEAK_THEORETICAL_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12                                              # 74.4


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--parity", action="store_true", help="parity-symmetric wave function (RNNwavefunction_paritysym)")
    ap.add_argument("--ns", type=int, default=NS_PER_GPU, help="samples per GPU (default: the BASELINE config)")
    ap.add_argument("--n-sites", type=int, default=N_SITES)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="target CPU seconds per reference-arm step")
    return ap.parse_args()


def workload(args):
    return {"workload": f"1D TFIM N={args.n_sites} Bx=1 OBC, pRNN {LAYERS}x GRU({UNITS}), {args.ns} samples per GPU"
                        + (", parity-symmetric" if args.parity else ""),
            "n_sites": args.n_sites, "layers": LAYERS, "units": UNITS, "samples_per_gpu": args.ns, "parity": bool(args.parity),
            "params": 38502}


# ------------------------------------------------------------------------------------------------
# CPU arm: the NumPy restatement of the reference algorithm (full recompute of all (N+1) configurations
# per sample in <=25 000-row chunks, fp32 GRU, f64 log-sum; autograd gradient; TF1 Adam)
# ------------------------------------------------------------------------------------------------
class CpuReference:
    def __init__(self, n_sites, parity):
        from oracle import rnnwf_oracle as O
        from oracle import torch_grad as TG
        self.O, self.TG = O, TG
        try:
            import torch
            torch.set_num_threads(os.cpu_count() or 1)
        except Exception:
            pass
        self.N, self.parity = n_sites, parity
        self.units = [UNITS] * LAYERS
        self.p = O.init_gru_params(self.units, seed=111, dtype=np.float32)
        self.Jz = np.ones(n_sites)
        self.m = np.zeros(O.num_params(O.gru_param_shapes(self.units)))
        self.v = np.zeros_like(self.m)
        self.t = 0
        self.it = 0

    def step(self, ns):
        O, TG = self.O, self.TG
        s = O.sample(self.p, ns, self.N, seed=111 + self.it)
        self.it += 1
        lp = (lambda c: O.log_probability_parity(self.p, c)) if self.parity else (lambda c: O.log_probability(self.p, c))
        e = O.ising_local_energies(self.Jz, BX, s, lp)
        w = (e - e.mean()) / ns
        g = TG.gru_vmc_grad({k: v.astype(np.float64) for k, v in self.p.items()}, s, w, parity=self.parity)
        theta = O.flatten(self.p).astype(np.float64)
        theta, self.m, self.v, self.t = O.adam_tf1(theta, g, self.m, self.v, self.t, LR)
        self.p = O.unflatten(theta.astype(np.float32), O.gru_param_shapes(self.units), np.float32)
        return float(e.mean())

    def calibrate(self, target_s):
        """Pick a sample count whose step takes about target_s (work is linear in ns: (N+1) rows per sample)."""
        t0 = time.perf_counter()
        self.step(1)
        t1 = time.perf_counter() - t0
        return max(1, min(64, int(target_s / max(t1, 1e-3))))


def cpu_threads():
    try:
        import torch
        return int(torch.get_num_threads())
    except Exception:
        return os.cpu_count() or 1


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ref = CpuReference(args.n_sites, args.parity)
    ns = ref.calibrate(args.cpu_seconds)
    for _ in range(max(0, args.warmup - 1)):      # calibrate() already ran one untimed step
        ref.step(ns)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref.step(ns)
    dt = time.perf_counter() - t0
    val = ns * args.steps / dt
    sample = f"{ns} samples per step x {args.steps} steps of the same workload: full recompute of ({args.n_sites}+1) configurations per sample"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload(args),
            "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cpu_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "NumPy/torch-CPU restatement of the reference algorithm (oracle/); TensorFlow 1.13 is not installable here"}
    print(json.dumps(line), flush=True)


METRIC = "local-energy samples/s through one full VMC step (sample + E_loc + gradient + all-reduce + Adam), 1D TFIM N=1000 GRU"


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
def run_b200_arm(args):
    import torch
    import torch.distributed as dist

    from rnnwavefunctions_b200 import ops
    from rnnwavefunctions_b200.training import Ising_local_energies
    from rnnwavefunctions_b200.vmc import TFIM, VMC
    from rnnwavefunctions_b200.wavefunction import RNNwavefunction1D, RNNwavefunctionParity

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N, ns = args.n_sites, args.ns
    cls = RNNwavefunctionParity if args.parity else RNNwavefunction1D
    wf = cls(N, units=[UNITS] * LAYERS, seed=111, device=dev)
    H = TFIM(np.ones(N), BX)
    opt = VMC(wf, H, ns)
    ndir = 2 if args.parity else 1

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
    marks = [[ev() for _ in range(4)] for _ in range(args.steps)]
    e0, e1 = ev(), ev()
    barrier()
    e0.record()
    means = []
    for k in range(args.steps):
        marks[k][0].record()
        s = opt.draw()
        marks[k][1].record()
        e = opt.local_energies(s)
        marks[k][2].record()
        mean, var, n = opt.moments(e)
        g = opt.gradient(s, e, mean, n)
        opt.apply(g, LR)
        marks[k][3].record()
        means.append(mean)
    e1.record()
    barrier()
    launches, dom_n, dom_ms = ops.profile_end()
    clk = clocks.stop() if clocks else None
    t_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    t_ms = float(t_ms.item())
    stage = np.array([[m[i].elapsed_time(m[i + 1]) for i in range(3)] for m in marks]).mean(axis=0)
    value = world * ns * args.steps / (t_ms * 1e-3)

    # ---- roofline of the dominant kernel (prefix-reuse chain kernel), timed live above ---------------
    chain_flops = ndir * ns * (N * (N - 1) / 2.0) * FLOP_PER_CELLSTACK       # algorithmic flops per launch
    chain_ms = dom_ms / max(dom_n, 1)
    achieved = chain_flops / (chain_ms * 1e-3) / 1e12
    ffma_meas = ops.ffma_peak(20000)
    step_flops = ns * FLOP_PER_CELLSTACK * (N + ndir * (N * (N + 1) / 2.0 + 3 * N))   # sample + E_loc + gradient
    mode = ops.tfim_chain_mode(wf.model)
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
        if (wl["n_sites"], wl["layers"], wl["units"], wl["samples_per_gpu"], bool(wl["parity"])) == (N, LAYERS, UNITS, ns, bool(args.parity)) \
                and ops.tfim_chain_mode(wf.model) == 3:
            traffic = int(tj["dram_bytes_read"]) + int(tj["dram_bytes_write"])
            traffic_note = "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum), recorded ncu capture: " + tj["how"]
    except Exception:
        pass
    common = {"achieved": achieved, "unit": "TFLOP/s", "kernel_ms": chain_ms, "kernel_share_of_step": dom_ms / t_ms, "launches_timed": dom_n,
              "flops_per_launch": chain_flops, "traffic": traffic, "traffic_note": traffic_note, "fp32_ffma_peak_measured": ffma_meas,
              "fp32_ffma_peak_theoretical": FP32_PEAK_THEORETICAL_TFLOPS, "achieved_over_fp32_ffma_peak": achieved / ffma_meas,
              "step_algorithmic_tflops": step_flops * args.steps / (t_ms * 1e-3) / 1e12}
    if mode == 0:
        roofline = dict(common, bound="fp32-ffma", kernel="gru_chain_kernel<float> (CUDA-core FFMA tile engine)", peak=ffma_meas,
                        frac=achieved / ffma_meas,
                        peak_kind="FP32 FFMA peak measured in this run (rnnwf_ffma_peak); MEASURED_PEAKS.json holds only HBM and bf16-tensor "
                                  "peaks, neither bounds a CUDA-core FP32 kernel")
    else:
        # executed tensor-pipe flops: 3 operand passes, padded tiles: per (site, 128-row tile, layer) MMAs of
        # 2*128*N*K flops each
        tiles128 = -(-(ndir * (-(-ns // 120)) * 120) // 128)
        if mode == 3:
            per_l0 = 2 * 128 * 16 * ((2 + 11) * 160 + 176)   # x group: 2 one-hot MMAs, h group: 12 (the first split 112 + 64), N = 160
            per_l1 = 2 * 128 * 16 * (23 * 160 + 176)
            kname = ("tc16p::chain_kernel<false,false> (tcgen05 kind::f16, 3xFP16 operands, weights resident in shared memory, merged "
                     "N=160 gate-block MMAs, MMA and gate math software-pipelined over anti-diagonals of the (site, layer) grid, one "
                     "specialised step copy per layer)")
        elif mode == 2:
            per_l0 = 2 * 128 * 16 * (2 * 160 + 64 + 12 * 160)
            per_l1 = 2 * 128 * 16 * (12 * 160 + 64 + 12 * 160)
            kname = "tc16::chain_kernel<50,false> (tcgen05 kind::f16, 3xFP16 operands, weights resident in shared memory)"
        else:
            per_l0 = 2 * 128 * 8 * (4 * 192 + 21 * (128 + 64))
            per_l1 = 2 * 128 * 8 * (21 * 192 + 21 * (128 + 64))
            kname = "gru_chain_tc_kernel<50,false> (tcgen05 kind::tf32, 3xTF32 operands)"
        executed = tiles128 * (N * (N - 1) / 2.0) * (per_l0 + (LAYERS - 1) * per_l1)
        peak = float(peaks.get("bf16_tflops_sustained", 1399.0)) / (1.0 if mode >= 2 else 2.0)
        roofline = dict(common, bound="tensor", kernel=kname, peak=peak, frac=achieved / peak,
                        peak_kind=("dense 16-bit tensor peak, sustained figure of MEASURED_PEAKS.json (kernel timed inside a multi-second step)"
                                   if peaks else "fallback: 1.4 PFLOP/s sustained bf16 (B200_PROFILING.md)") +
                                  ("" if mode >= 2 else "; kind::tf32 runs at half the 16-bit rate"),
                        peak_over_passes=peak / 3.0, frac_of_peak_over_passes=achieved / (peak / 3.0),
                        tf32x3_roofline=(peak / 2.0 if mode >= 2 else peak) / 3.0,
                        tensor_pipe_tflops_executed=executed / (chain_ms * 1e-3) / 1e12,
                        tensor_pipe_frac_executed=executed / (chain_ms * 1e-3) / 1e12 / peak,
                        note_rooflines="peak_over_passes = tensor peak / 3: the ceiling of any 3-pass split-operand scheme with FP32-grade accuracy on "
                                       "this pipe; tf32x3_roofline = (16-bit peak / 2) / 3: the ceiling of the 3xTF32 scheme BASELINE.json's north_star names",
                        note="achieved counts ALGORITHMIC flops (75 800 per GRU-stack evaluation); FP32-grade accuracy costs 3 tensor passes "
                             "over padded tiles (K 51 -> 64, N 150 -> 160), so the tensor pipe executes ~3.6-4x the algorithmic flops; the "
                             "instructions are N = 160 wide because one with a new A chunk pays ~81 cycles of TMEM operand fetch whatever its N "
                             "(measured, scripts/mma_probe2.py); the recurrence leaves one (site, layer) step of look-ahead and a single "
                             "accumulator set fits TMEM, so the gate math (XU pipe ~70 % busy, ncu) and the MMAs (tensor pipe ~71 %) overlap "
                             "step against step")

    # ---- end to end through the reference-facing host API (host buffers, copies inside the timed region) ----
    e2e = None
    if not args.no_e2e:
        Jz = np.ones(N)
        k_e2e = max(1, min(args.steps, 2))
        h2d = d2h = 0

        def host_step():
            nonlocal h2d, d2h
            samples = wf.sample(ns, 2).cpu().numpy()                         # sess.run(samples_) -> NumPy (int64)
            d2h += samples.nbytes
            eloc = Ising_local_energies(Jz, BX, samples, None, wf, None, None, None)   # host in, host out
            h2d += ns * N
            d2h += eloc.nbytes
            su8 = ops.as_u8_samples(samples, dev, N)                          # feed_dict {samp: samples, Eloc: local_energies}
            el = torch.as_tensor(eloc).to(dev)
            h2d += ns * N + eloc.nbytes
            mean, var = opt.step_from(su8, el, LR)
            m = float(mean.item())
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
               "api": "RNNwavefunction.sample -> NumPy; Ising_local_energies(host samples) -> NumPy; optimiser step fed from host arrays"}

    # ---- CPU baseline (rank 0, N=1 only) -------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ref = CpuReference(N, args.parity)
        ns_cpu = ref.calibrate(args.cpu_seconds)
        t0 = time.perf_counter()
        ref.step(ns_cpu)
        dt = time.perf_counter() - t0
        cpu = {"value": ns_cpu / dt, "unit": "samples/s", "cores": cpu_threads(), "kind": "port",
               "sample": f"{ns_cpu} samples of the same workload in {dt:.1f} s (work is linear in the sample count: N+1 configurations "
                         f"per sample, full recompute as the reference does)"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": t_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": dict(workload(args), l2="per-step working set (6 GB hidden-state stash per GPU) exceeds L2; no flush needed",
                                                     parallelism=f"dp{world}"),
                "vmc_steps_per_s": args.steps / (t_ms * 1e-3), "eloc_samples_per_s": world * ns / (stage[1] * 1e-3),
                "stages_ms": {"sample": stage[0], "local_energies": stage[1], "moments+gradient+adam": stage[2]},
                "mean_energy_last_step": float(means[-1].item()),
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clk}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_b200_arm(a)
