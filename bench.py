#!/usr/bin/env python
"""bench.py — Kalman filter + RTS smoother + ELBO forward and explicit-adjoint backward on B200.

    python bench.py --gpus N --steps K --warmup W [--workload cfg2|cfg3|cfg4|cfg5]     # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ... [--workload ...]            # the reference algorithm, host CPU

Workloads (BASELINE.json `configs`, shapes in kalman_vae_b200/synthetic.py):
  cfg2 (default, the configuration the metric is quoted on): B=8192 sequences PER GPU (weak scaling), T=20, z_dim=4,
        a_dim=2, u_dim=4, K=3; one step = smooth (filter+smoother) + elbo + backward (all gradients).
  cfg3  imputation: T=1000, 65 536 sequences IN TOTAL sharded over the ranks (strong scaling), Bernoulli-masked missing
        observations, forward only (smooth), no collective.
  cfg4  SKVAE switching dynamics, K=8, z_dim=16, a_dim=8, T=200, B=16 384 per GPU, fwd+bwd (FMA bound: both the HBM and the
        fp32-FMA fraction are reported).
  cfg5  full KVAE training step: the reference's KVAE module (PyTorch conv encoder/decoder, kvae/model/model.py) with this
        package's Kalman block swapped in, Adam + grad clipping as kvae/train/train.py:32-58,
        batch 32 x ranks, NCCL all-reduce of the gradients.  Needs the reference sources (baseline/_ref or /root/reference).

One JSON line on stdout (rank 0):
  value      sequence-steps/s, whole job, inputs resident in HBM, CUDA-event timed, max over ranks (the median of the
             timed blocks is listed too)
  e2e        the same metric through the reference-shaped API (KalmanFilter.smooth -> .elbo -> autograd.grad) with pinned
             HOST inputs: every step's inputs cross PCIe and its loss + parameter gradients come back, all inside the timed
             region (median of >= 3 blocks); e2e_engine = the same through engine.HostPipeline
  roofline   dominant kernel's algorithmic HBM bytes / its CUDA-event duration vs MEASURED_PEAKS.json
  cpu_baseline  the CPU port of the reference algorithm (oracle/) timed on this box's host cores
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from kalman_vae_b200.synthetic import CONFIGS, Shape, make_case  # noqa: E402

# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures (profiles/)
NCU_TRAFFIC_BYTES = {
    ("cfg2", "k_bwd"): 54.6e6,              # profiles/r01j_ncu_full_summary_cfg2.csv (lane-group adjoint, B=8192)
    ("cfg2", "k_filter_smooth"): 20.3e6,    # profiles/r02_ncu_summary.txt
}

UNIT = "sequence-steps/s"
METRICS = {
    "cfg2": "kalman_filter_smoother_fwd_bwd_sequence_steps_per_sec",
    "cfg3": "kalman_filter_smoother_fwd_sequence_steps_per_sec",
    "cfg4": "kalman_filter_smoother_fwd_bwd_sequence_steps_per_sec",
    "cfg5": "kvae_training_step_sequence_steps_per_sec",
}


def workload_text(name: str) -> str:
    """The `config.workload` string: identical in the CUDA arm and the reference arm."""
    if name == "cfg5":
        return ("cfg5: full KVAE training step (reference KVAE module, 32x32 frames, T=20, a_dim=2, z_dim=4, K=3, batch 32 per "
                "GPU), forward + loss + backward + grad clip + Adam")
    s = CONFIGS[name]
    what = {"cfg2": "smooth+elbo forward and backward", "cfg3": "smooth forward (imputation, Bernoulli-masked observations)",
            "cfg4": "smooth+elbo forward and backward (switching dynamics: Q per mode, shared C)"}[name]
    per = "in total, sharded over the GPUs" if name == "cfg3" else "per GPU"
    return f"{name}: B={s.B} sequences {per}, T={s.T}, n={s.n}, p={s.p}, m={s.m}, K={s.K}; {what}"


def algorithmic_bytes(shape: Shape):
    """bytes per sequence-step with every tensor of the public contract touched once."""
    n, p, m, K = shape.n, shape.p, shape.m, shape.K
    c = 0 if shape.c_shared else 1
    fwd = 4 * (p + m + 1 + K + 3 * n + 4 * n * n + n * m + c * p * n)
    elbo = 4 * n
    bwd = 4 * (2 * p + m + 1 + 2 * K + n + 3 * n + 3 * n * n)
    return dict(fwd=fwd, elbo=elbo, bwd=bwd, total=fwd + elbo + bwd)


def flops_per_seq_step(shape: Shape):
    """SURVEY.md §8(d) estimate (for the FMA-bound cfg4): fwd = 8n^3 + 8n^2 p (filter) + 8.3 n^3 (smoother) + mixing;
    fwd+bwd ~ 3.5x."""
    n, p, m, K = shape.n, shape.p, shape.m, shape.K
    fwd = 8 * n ** 3 + 8 * n * n * p + 8.3 * n ** 3 + 2 * K * (2 * n * n + n * m + p * n)
    return dict(fwd=fwd, total=3.5 * fwd)


FP32_FMA_PEAK_TFLOPS = 72.0   # 148 SMs x 128 lanes x 2 flop x 1.965 GHz (tools/microbench/ub_stream.cu measures 72.4)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU legs run on rank 0 alone and may use the whole host."""
    n = os.cpu_count() or 1
    if torch.get_num_threads() < n:
        torch.set_num_threads(n)


# ----------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the CPU port of the reference algorithm (oracle/), all host threads
# ----------------------------------------------------------------------------------------------------
def cpu_case(name: str, sample_B: int):
    s = CONFIGS[name]
    mk = "bernoulli" if name == "cfg3" else "ones"
    return make_case(Shape(sample_B, s.T, s.n, s.p, s.m, s.K, s.q_per_mode, s.c_shared), seed=10, mask_kind=mk)


def cpu_one_pass(name: str, case):
    """One pass of the workload's hot path with the oracle (the one place bench.py may execute oracle/)."""
    from oracle import kalman_oracle as ko
    if name == "cfg3":
        ko.run_case(case, torch.float32, want_grads=False, with_elbo=False)
    else:
        ko.smooth_elbo_fwd_bwd(case, torch.float32, backward=True)


def cpu_steps_per_sec(name: str, sample_B: int, reps: int, warmup: int = 1):
    case = cpu_case(name, sample_B)
    for _ in range(warmup):
        cpu_one_pass(name, case)
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        cpu_one_pass(name, case)
        times.append(time.perf_counter() - t0)
    return sample_B * CONFIGS[name].T / min(times), times


CPU_SAMPLE_B = {"cfg2": 8192, "cfg3": 128, "cfg4": 32}   # sequences per CPU pass (cfg2: the full batch)


def run_reference(args, rank, world):
    """Times the reference algorithm (CPU port, all host threads).  Each step is a bounded sample of the workload's
    batch, sized from a short calibration so that warmup+steps finish in about two minutes."""
    if rank != 0:
        return
    name = args.workload
    if name == "cfg5":
        return run_reference_cfg5(args)
    shape = CONFIGS[name]
    use_all_host_threads()
    threads = torch.get_num_threads()
    t0 = time.perf_counter()
    cal_B = {"cfg2": 512, "cfg3": 32, "cfg4": 8}[name]
    calib = cpu_case(name, cal_B)
    cpu_one_pass(name, calib)
    t1 = time.perf_counter()
    cpu_one_pass(name, calib)
    per_seq = (time.perf_counter() - t1) / cal_B
    budget = 170.0
    lo, q = {"cfg2": (1024, 256), "cfg3": (32, 16), "cfg4": (8, 4)}[name]
    sample_B = int(min(shape.B, max(lo, budget / ((args.steps + args.warmup) * per_seq))))
    sample_B = max(lo, (sample_B // q) * q)
    case = cpu_case(name, sample_B)
    for _ in range(args.warmup):
        cpu_one_pass(name, case)
    tt = []
    for _ in range(args.steps):
        t1 = time.perf_counter()
        cpu_one_pass(name, case)
        tt.append(time.perf_counter() - t1)
    ms = 1e3 * sum(tt) / len(tt)
    value = sample_B * shape.T / (ms * 1e-3)
    line = {
        "impl": "reference", "metric": METRICS[name], "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if name == "cfg3" else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_text(name),
                   "reference_arm": "reference algorithm on the host CPU (oracle/ port: torch CPU ops in the reference's op "
                                    "order + autograd), per-step sample of the workload's batch"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample_B} of {shape.B} sequences x T={shape.T} per step, mean of {args.steps} steps; "
                                   f"os.cpu_count()={os.cpu_count()}, torch threads={threads}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t0,
    }
    emit(line)


def run_reference_cfg5(args):
    """The UNMODIFIED reference KVAE training step on the host CPU (all threads), batch 32."""
    from kalman_vae_b200 import kvae_step
    use_all_host_threads()
    threads = torch.get_num_threads()
    t0 = time.perf_counter()
    try:
        stepper = kvae_step.ReferenceTrainStep(device=torch.device("cpu"), drop_in=False, seed=10, dynamics_model=args.dynamics)
    except Exception as err:
        emit({"impl": "reference", "unavailable": f"reference sources not importable: {type(err).__name__}: {err}"[:300]})
        return
    B, T = stepper.batch, stepper.T
    x = stepper.synthetic_batch(seed=1)
    for _ in range(args.warmup):
        stepper.step(x)
    tt = []
    for _ in range(args.steps):
        t1 = time.perf_counter()
        stepper.step(x)
        tt.append(time.perf_counter() - t1)
    ms = 1e3 * sum(tt) / len(tt)
    value = B * T / (ms * 1e-3)
    emit({"impl": "reference", "metric": METRICS["cfg5"], "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
          "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
          "dtype": "f32", "data": "synthetic",
          "config": {"workload": workload_text("cfg5"), "dynamics_model": args.dynamics,
                     "reference_arm": "the unmodified reference KVAE + its own Kalman filter, torch CPU, all host threads"},
          "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference",
                           "sample": f"batch {B} x T={T} per step, mean of {args.steps} steps"},
          "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
          "wall_s": time.perf_counter() - t0})


# ----------------------------------------------------------------------------------------------------
# CUDA arm: shared helpers
# ----------------------------------------------------------------------------------------------------
class Ctx:
    def __init__(self, args, rank, local_rank, world):
        import torch.distributed as dist
        from kalman_vae_b200 import capi
        capi.lib()   # fail loudly if the CUDA library is missing
        self.args, self.rank, self.world, self.dist = args, rank, world, dist
        self.dev = torch.device("cuda", local_rank)
        torch.cuda.set_device(self.dev)
        if world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.stream = torch.cuda.current_stream(self.dev)
        self.peak, self.peak_src = peaks()
        self.sampler = ClockSampler(local_rank)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, ms):
        t = torch.tensor([ms], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t)

    def timed_blocks(self, fn, steps, blocks=1, stream=None):
        """`blocks` timed regions of `steps` calls each: barrier + synchronize on both sides, CUDA events on the
        launching stream, max over ranks.  Returns ms per step of every block."""
        stream = stream or self.stream
        out = []
        for _ in range(blocks):
            self.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for i in range(steps):
                fn(i)
            e1.record(stream)
            self.barrier()
            out.append(self.max_over_ranks(e0.elapsed_time(e1)) / steps)
        return out

    def median_call_seconds(self, fn, reps):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize(self.dev)
        evs = []
        for i in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(self.stream)
            fn(i)
            b.record(self.stream)
            evs.append((a, b))
        torch.cuda.synchronize(self.dev)
        return statistics.median(x.elapsed_time(y) for x, y in evs) * 1e-3

    def finish(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def to_dev(case, dev):
    return {k: (v.to(dev).contiguous() if torch.is_tensor(v) else v) for k, v in case.items()}


def make_problem(g, shape, lanes, with_u=True):
    from kalman_vae_b200.functional import Problem
    return Problem(g["Y"], g["U"] if with_u else None, g["mask"], g["alpha"], g["A"], g["B"], g["C"], g["Q"], g["R"], g["mu0"],
                   g["Sigma0"], shape.q_per_mode, shape.c_shared, lanes=lanes)


def cpu_baseline_block(name, shape, dev, gpu_context=True):
    use_all_host_threads()
    reps = 3 if name == "cfg2" else 2
    sample_B = CPU_SAMPLE_B[name]
    cpu_val, _ = cpu_steps_per_sec(name, sample_B, reps)
    nthreads = torch.get_num_threads()
    one_B = {"cfg2": 1024, "cfg3": 16, "cfg4": 4}[name]
    torch.set_num_threads(1)
    try:
        cpu_1t, _ = cpu_steps_per_sec(name, one_B, 1)
    finally:
        torch.set_num_threads(nthreads)
    gpu_ref = None
    if gpu_context and name == "cfg2":
        # context: the same op sequence (stock ATen ops + autograd, what the reference executes) on THIS GPU
        try:
            from oracle import kalman_oracle as ko
            gcase = to_dev(make_case(shape, seed=10), dev)
            ko.smooth_elbo_fwd_bwd(gcase, torch.float32, backward=True)
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            ko.smooth_elbo_fwd_bwd(gcase, torch.float32, backward=True)
            torch.cuda.synchronize(dev)
            gpu_ref_s = time.perf_counter() - t0
            del gcase
            gpu_ref = {"value": shape.B * shape.T / gpu_ref_s, "unit": UNIT, "ms_per_step": gpu_ref_s * 1e3,
                       "note": "the port's torch op sequence (the reference's stock-ATen path incl. autograd) run with CUDA "
                               "tensors on this B200, full cfg2 batch, wall clock of the second run: kernel-launch bound "
                               "(~10^4 launches per step)"}
        except Exception as err:   # context only: never let it break the bench line
            gpu_ref = {"value": None, "note": f"not measured: {type(err).__name__}: {err}"[:300]}
    what = "smooth forward" if name == "cfg3" else "smooth+elbo fwd+bwd"
    return {"value": cpu_val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "reference_ops_on_this_gpu": gpu_ref, "value_1_thread": cpu_1t,
            "sample_1_thread": f"{one_B} sequences x T={shape.T}, torch.set_num_threads(1), one run after warm-up",
            "sample": f"{sample_B} of {shape.B} sequences x T={shape.T}, {what} with the oracle (torch CPU ops in the reference's "
                      f"op order), best of {reps}, scaled linearly; os.cpu_count()={os.cpu_count()}"}


# ----------------------------------------------------------------------------------------------------
# cfg2 / cfg4: the training step (smooth + elbo + backward)
# ----------------------------------------------------------------------------------------------------
def run_train(cx: Ctx, name: str):
    from kalman_vae_b200 import KalmanFilter, capi
    from kalman_vae_b200.dyn_param import PrecomputedWeights
    from kalman_vae_b200.engine import HostPipeline, KalmanStep
    args, rank, world, dev, dist = cx.args, cx.rank, cx.world, cx.dev, cx.dist
    shape = CONFIGS[name]
    ab = algorithmic_bytes(shape)
    lanes = args.lanes
    big = name == "cfg4"
    steps = args.steps if args.steps > 0 else (10 if big else 2000)
    warmup = args.warmup

    # ---- device-resident buffer sets, rotated so that consecutive steps never re-use L2-resident data
    #      (cfg4: one set is 3.4 GB of states alone, far beyond L2)
    nsets = 1 if big else args.buffer_sets
    sets = []
    for s in range(nsets):
        g = to_dev(make_case(shape, seed=10 + 97 * rank + s), dev)
        sets.append(KalmanStep(make_problem(g, shape, lanes), g["eps"], use_graphs=not args.no_graphs, need_dU=False))
    lanes_used = capi.pick_lanes(sets[0].pb.dims) if lanes == 0 else lanes
    k0 = sets[0]
    set_bytes = sum(t.numel() * t.element_size() for t in
                    [k0.pb.Y, k0.pb.U, k0.pb.alpha, k0.eps, k0.st.mus_filt, k0.st.Sigmas_filt, k0.st.mus_pred, k0.st.Sigmas_pred,
                     k0.st.mus_smooth, k0.st.Sigmas_smooth, k0.A_list, k0.B_list, k0.ws_bwd] + ([k0.C_list] if k0.C_list is not None else []))

    if rank == 0:
        cx.sampler.start()          # samples through warm-up, the timed region and the per-kernel timing (all under load)
    for i in range(max(warmup, 3)):
        sets[i % nsets].step()
    nblocks = 1 if big else 3
    blocks = cx.timed_blocks(lambda i: sets[i % nsets].step(), steps, blocks=nblocks)
    ms_step = blocks[0]                       # the contract's number: the FIRST block of exactly K steps
    value = world * shape.B * shape.T / (ms_step * 1e-3)
    for s in sets:
        s.check()                             # a pivot failure or an exchange time-out would have left garbage

    # ---- data-parallel correctness (N > 1): identical parameter gradients on every rank, and agreement with the
    #      torch.distributed (NCCL) all-reduce path on the same data
    dp_check = None
    if world > 1:
        dp_check = dp_consistency(cx, shape, lanes, sets[0])

    # ---- per-kernel durations (CUDA events around each C-ABI call, rotating sets), for the roofline
    kt = {}
    if rank == 0:
        from kalman_vae_b200 import capi as _c
        fwd = lambda i: (lambda s: _c.filter_smooth_fwd(s.pb.dims, s._inputs, s._states, s.A_list, s.B_list, s.C_list, s.info, s.dev))(sets[i % nsets])
        bwd = lambda i: (lambda s: _c.bwd(s.dims_bwd, s._inputs, s._states, s.eps, s.jitter, s.g_elbo, s.terms, None, s.grads,
                                          s.ws_bwd, s.info, s.dev))(sets[i % nsets])
        reps = 5 if big else args.kernel_reps
        fam = "k_seq" if lanes_used == 1 and shape.n == 4 and shape.T % 4 == 0 else "k"
        kt[f"{fam}_filter_smooth" if fam == "k" else "k_seq_fwd"] = cx.median_call_seconds(fwd, reps)
        kt["k_bwd(+bwd_final)" if fam == "k" else "k_seq_bwd(+bwd_final)"] = cx.median_call_seconds(bwd, reps)

    # ---- the same step in the throughput regime (the machine full: B = 262 144 sequences), for the roofline discussion:
    #      cfg2 itself is 1 024 warps = 11 % of the warp slots and is latency bound
    thr = None
    if rank == 0 and world == 1 and not args.no_throughput and name == "cfg2":
        bshape = Shape(262144, shape.T, shape.n, shape.p, shape.m, shape.K, shape.q_per_mode, shape.c_shared)
        gb = to_dev(make_case(bshape, seed=77), dev)
        ksb = KalmanStep(make_problem(gb, bshape, lanes), gb["eps"], use_graphs=not args.no_graphs, need_dU=False)
        for _ in range(3):
            ksb.step()
        ms_big = statistics.median(cx.timed_blocks(lambda i: ksb.step(), 20, blocks=3))
        ms_fwd = statistics.median(cx.timed_blocks(lambda i: ksb.forward_only(), 20, blocks=3))
        ksb.check()
        n_big = bshape.B * bshape.T
        thr = {"B": bshape.B, "T": bshape.T, "lanes_per_sequence": capi.pick_lanes(ksb.pb.dims) if lanes == 0 else lanes,
               "kernels": "k_seq_fwd, k_seq_bwd (thread per sequence, TMA-staged streams)" if (capi.pick_lanes(ksb.pb.dims) if lanes == 0 else lanes) == 1
                          else "k_filter_smooth, k_bwd (lane groups)",
               "ms_per_step": ms_big, "value": n_big / (ms_big * 1e-3), "unit": UNIT,
               "whole_step_frac": ab["total"] * n_big / (ms_big * 1e-3) / 1e9 / cx.peak,
               "forward_ms": ms_fwd, "forward_frac": ab["fwd"] * n_big / (ms_fwd * 1e-3) / 1e9 / cx.peak,
               "note": "32x the cfg2 batch (4 GB touched per step > L2): where HBM is the binding roofline; median of 3 blocks "
                       "of 20 steps"}
        del ksb, gb
        torch.cuda.empty_cache()
    clocks = cx.sampler.stop() if rank == 0 else None

    # ---- end to end with pinned HOST inputs
    e2e = e2e_engine = None
    if not args.no_e2e and not big:
        e2e, e2e_engine = e2e_train(cx, shape, lanes, steps)

    if rank == 0:
        dom = max(kt, key=kt.get)
        dom_bytes = (ab["fwd"] if "fwd" in dom or "filter" in dom else ab["bwd"]) * shape.B * shape.T
        achieved = dom_bytes / kt[dom] / 1e9
        roof = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": cx.peak, "unit": "GB/s", "frac": achieved / cx.peak,
                "traffic": NCU_TRAFFIC_BYTES.get((name, dom.split("(")[0])), "peak_source": cx.peak_src,
                "algorithmic_bytes_per_seq_step": ab, "kernel_seconds": kt,
                "whole_step_frac": ab["total"] * world * shape.B * shape.T / (ms_step * 1e-3) / 1e9 / (cx.peak * world),
                "throughput_regime": thr}
        if big:   # n = 16 is fp32-FMA bound, not HBM bound: report both fractions
            fl = flops_per_seq_step(shape)
            tfl = fl["total"] * shape.B * shape.T / (ms_step * 1e-3) / 1e12
            roof["fma"] = {"flops_per_seq_step_estimate": fl, "achieved_tflops": tfl, "peak_tflops": FP32_FMA_PEAK_TFLOPS,
                           "frac": tfl / FP32_FMA_PEAK_TFLOPS,
                           "note": "~39 flop per algorithmic byte: the fp32 FMA pipes, not HBM, bound this shape"}
        cpu = None if args.no_cpu else cpu_baseline_block(name, shape, dev)
        line = {
            "metric": METRICS[name], "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "value_median_of_blocks": world * shape.B * shape.T / (statistics.median(blocks) * 1e-3), "ms_per_step_blocks": blocks,
            "config": {"workload": workload_text(name), "lanes_per_sequence": lanes_used, "cuda_graphs": not args.no_graphs,
                       "l2": (f"{nsets} rotating buffer sets of {set_bytes / 2**20:.0f} MiB each (> 126 MB L2 in total)" if nsets > 1
                              else f"one buffer set of {set_bytes / 2**20:.0f} MiB (inputs larger than L2)"),
                       "sharding": "batch dimension, contiguous per rank; ONE exchange per step of a flat buffer [parameter gradients | 5 ELBO sums]",
                       "collective": sets[0].collective},
            "e2e": e2e, "e2e_eager": (e2e or {}).get("eager"), "e2e_engine": e2e_engine, "dp_check": dp_check,
            "gpu_launches": sets[0].kernel_launches_per_step * steps,
            "roofline": roof, "cpu_baseline": cpu, "clocks": clocks,
        }
        emit(line)
    for s in sets:
        s.close()


def dp_consistency(cx: Ctx, shape, lanes, ks_peer):
    """N > 1: (1) the parameter gradients / ELBO terms left by the peer-memory exchange are bit-identical on every
    rank; (2) one step through the torch.distributed all-reduce path on the same data agrees with them."""
    from kalman_vae_b200.engine import KalmanStep
    dist, dev, world = cx.dist, cx.dev, cx.world
    ks_peer.step()
    torch.cuda.synchronize(dev)
    flat = ks_peer.flat[:ks_peer.n_reduce].clone()
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    identical = all(torch.equal(gathered[0], g) for g in gathered[1:])
    ks_nccl = KalmanStep(ks_peer.pb, ks_peer.eps, use_graphs=False, need_dU=False, collective="nccl")
    ks_nccl.step()
    torch.cuda.synchronize(dev)
    ref = ks_nccl.flat[:ks_nccl.n_reduce]
    rel = float((flat.double() - ref.double()).norm() / ref.double().norm().clamp_min(1e-30))
    rel_dy = float((ks_peer.grads["dY"].double() - ks_nccl.grads["dY"].double()).norm() / ks_nccl.grads["dY"].double().norm().clamp_min(1e-30))
    finite = bool(torch.isfinite(flat).all())
    t = torch.tensor([rel, rel_dy, 0.0 if (identical and finite) else 1.0], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = bool(t[2] == 0) and float(t[0]) <= 2e-5 and float(t[1]) <= 2e-5
    return {"ok": ok, "bit_identical_across_ranks": bool(t[2] == 0), "rel_vs_nccl_allreduce_params_and_elbo": float(t[0]),
            "rel_vs_nccl_allreduce_dY": float(t[1]), "collective": ks_peer.collective, "tolerance": 2e-5}


def e2e_train(cx: Ctx, shape, lanes, steps):
    """(1) e2e: the reference-shaped calls KalmanFilter.smooth -> .elbo -> torch.autograd.grad on pinned HOST inputs; the
    next step's inputs are uploaded on a side stream while this step computes (what a prefetching data loader does), the
    loss and the parameter gradients come back to pinned host memory every step.  (2) e2e_engine: engine.HostPipeline."""
    from kalman_vae_b200 import KalmanFilter
    from kalman_vae_b200.dyn_param import PrecomputedWeights
    from kalman_vae_b200.engine import HostPipeline
    args, rank, world, dev, dist = cx.args, cx.rank, cx.world, cx.dev, cx.dist
    case = make_case(shape, seed=1234 + rank)
    params = {k: case[k].to(dev).contiguous() for k in ("A", "B", "C", "Q", "R", "mu0", "Sigma0")}
    names = ("Y", "U", "mask", "alpha", "eps")

    # ---------------- (2) engine
    pipe = HostPipeline((shape.B, shape.T, shape.n, shape.p, shape.m, shape.K), params, shape.q_per_mode, shape.c_shared,
                        lanes=lanes, device=dev)
    slab, views = pipe.host_slab()
    for k in names:
        views[k].copy_(case[k])
    state = {"prev": None, "elbo": None}

    def pipe_step(i):
        k = pipe.step_packed(slab)
        if state["prev"] is not None:
            state["elbo"], _ = pipe.result(state["prev"])
        state["prev"] = k

    for i in range(max(args.warmup, 10)):
        pipe_step(i)
    blocks = cx.timed_blocks(pipe_step, steps, blocks=3, stream=pipe.compute_stream)
    elbo_last, _ = pipe.result(state["prev"])
    ms = statistics.median(blocks)
    e2e_engine = {"value": world * shape.B * shape.T / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": pipe.h2d_bytes_per_step,
                  "d2h_bytes_per_step": pipe.d2h_bytes_per_step, "ms_per_step": ms, "ms_per_step_blocks": blocks,
                  "elbo_last_step": elbo_last,
                  "api": "engine.HostPipeline.step_packed(pinned host slab holding Y,U,mask,alpha,eps) + .result(): ONE H2D copy "
                         "per step on a copy stream overlapping the previous step, fwd+ELBO+bwd graph, D2H of parameter "
                         "gradients + ELBO terms + status word; median of 3 blocks"}
    pipe.close()

    # ---------------- (1) reference-shaped API
    dyn = PrecomputedWeights(case["A"], case["B"], case["C"], case["Q"] if shape.q_per_mode else None, switching=shape.q_per_mode)
    kf = KalmanFilter(0.02 ** 0.5, 0.03 ** 0.5, case["mu0"], case["Sigma0"], dyn, lanes=lanes).to(dev)
    host = {k: case[k].pin_memory() for k in names}
    h2d = sum(v.numel() * 4 for v in host.values())
    params_l = list(dyn.parameters())
    out_host = [torch.empty(1 + sum(p.numel() for p in params_l)).pin_memory() for _ in range(2)]
    d2h = out_host[0].numel() * 4
    copy_stream = torch.cuda.Stream(device=dev)
    main = cx.stream
    slots = [{k: torch.empty_like(v, device=dev) for k, v in host.items()} for _ in range(2)]
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    used = [False, False]

    cur_eps = [None]
    kf._draw_eps = lambda B, T, n, like: cur_eps[0]   # the rsample draw of kalman_filter.py:351, supplied from the host

    def upload(i):
        k = i % 2
        with torch.cuda.stream(copy_stream):
            if used[k]:
                copy_stream.wait_event(ev_free[k])
            for nm, src in host.items():
                slots[k][nm].copy_(src, non_blocking=True)
            ev_in[k].record(copy_stream)

    def auto_step(i):
        k = i % 2
        upload(i + 1)                      # next step's inputs cross PCIe while this step computes
        main.wait_event(ev_in[k])
        d = slots[k]
        Y = d["Y"].requires_grad_(True)
        dyn.set_weights(d["alpha"].requires_grad_(True))
        cur_eps[0] = d["eps"]
        outs = kf.smooth(Y, d["U"], d["mask"])
        val = kf.elbo(outs[0], outs[1], Y, d["U"], outs[6], outs[7], outs[8], mask=d["mask"])
        grads = torch.autograd.grad(val, [Y, dyn.alpha] + params_l)
        if world > 1:
            from kalman_vae_b200.dist import allreduce_param_grads
            allreduce_param_grads(list(grads[2:]))
        flat = torch.cat([val.detach().reshape(1)] + [g.reshape(-1) for g in grads[2:]])
        out_host[k].copy_(flat, non_blocking=True)
        ev_free[k].record(main)
        used[k] = True
        d["Y"].requires_grad_(False)
        d["alpha"].requires_grad_(False)

    n_auto = min(steps, 500)
    upload(0)
    base = [0]

    def run(i):
        auto_step(base[0] + i)

    for i in range(max(min(args.warmup, 20), 10)):   # lets the caching allocator reach its steady state
        auto_step(i)
    base[0] = max(min(args.warmup, 20), 10)
    blocks = []
    for _ in range(3):
        blocks += cx.timed_blocks(run, n_auto, blocks=1)
        base[0] += n_auto
    torch.cuda.synchronize(dev)
    ms = statistics.median(blocks)
    e2e_eager = {"value": world * shape.B * shape.T / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                 "ms_per_step": ms, "ms_per_step_blocks": blocks, "steps_per_block": n_auto,
                 "elbo_last_step": float(out_host[(base[0] - 1) % 2][0]),
                 "api": "KalmanFilter.smooth -> .elbo -> torch.autograd.grad (reference signatures), eager: host-bound (~0.4 ms of "
                        "Python / autograd bookkeeping per step); pinned host Y,U,mask,alpha,eps uploaded every step on a side "
                        "stream (double-buffered), loss + parameter gradients copied back to pinned host memory every step; "
                        "median of 3 blocks"}

    # ---------------- (1b) the SAME reference-shaped calls, captured once per input slot with torch.cuda.graph and replayed
    # (what a trainer that wants the kernels' speed does with a fixed-shape step; the module makes no host reads under
    # capture).  Per step: the upload of the next inputs, one replay, the copy-back of loss + parameter gradients.
    try:
        torch.cuda.synchronize(dev)
        kf.check_info = False
        graphs, flats = [], []
        for k in range(2):
            d = slots[k]
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                Y = d["Y"].requires_grad_(True)
                dyn.set_weights(d["alpha"].requires_grad_(True))
                cur_eps[0] = d["eps"]
                outs = kf.smooth(Y, d["U"], d["mask"])
                val = kf.elbo(outs[0], outs[1], Y, d["U"], outs[6], outs[7], outs[8], mask=d["mask"])
                grads = torch.autograd.grad(val, [Y, dyn.alpha] + params_l)
                flat = torch.cat([val.detach().reshape(1)] + [gr.reshape(-1) for gr in grads[2:]])
            d["Y"].requires_grad_(False)
            d["alpha"].requires_grad_(False)
            graphs.append(g)
            flats.append(flat)
        used[0] = used[1] = False

        def graph_step(i):
            k = i % 2
            upload(i + 1)
            main.wait_event(ev_in[k])
            graphs[k].replay()
            if world > 1:   # one NCCL all-reduce of the parameter gradients (what allreduce_param_grads does on the eager route)
                dist.all_reduce(flats[k][1:], op=dist.ReduceOp.SUM)
            out_host[k].copy_(flats[k], non_blocking=True)
            ev_free[k].record(main)
            used[k] = True

        upload(0)
        for i in range(10):
            graph_step(i)
        base[0] = 10
        blocks_g = []
        for _ in range(3):
            blocks_g += cx.timed_blocks(lambda i: graph_step(base[0] + i), n_auto, blocks=1)
            base[0] += n_auto
        torch.cuda.synchronize(dev)
        ms_g = statistics.median(blocks_g)
        e2e = {"value": world * shape.B * shape.T / (ms_g * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": ms_g, "ms_per_step_blocks": blocks_g, "steps_per_block": n_auto,
               "elbo_last_step": float(out_host[(base[0] - 1) % 2][0]),
               "api": "KalmanFilter.smooth -> .elbo -> torch.autograd.grad (reference signatures) captured once per input slot with "
                      "torch.cuda.graph and replayed; per step: pinned host Y,U,mask,alpha,eps uploaded on a side stream "
                      "(double-buffered), one replay, [N > 1: one NCCL all-reduce of the parameter gradients,] loss + parameter "
                      "gradients copied back to pinned host memory; median of 3 blocks.  The same calls without capture: see e2e_eager",
               "eager": e2e_eager}
    except Exception as err:   # a capture problem must not cost the line: report the eager measurement and say why
        torch.cuda.synchronize(dev)
        e2e = dict(e2e_eager)
        e2e["eager"] = e2e_eager
        e2e["graph_capture_failed"] = f"{type(err).__name__}: {err}"[:300]
    return e2e, e2e_engine


# ----------------------------------------------------------------------------------------------------
# cfg3: imputation (forward only), the batch sharded over the ranks
# ----------------------------------------------------------------------------------------------------
def run_impute(cx: Ctx):
    from kalman_vae_b200 import KalmanFilter, capi
    from kalman_vae_b200 import functional as F
    from kalman_vae_b200.dist import shard_bounds
    from kalman_vae_b200.dyn_param import PrecomputedWeights
    args, rank, world, dev = cx.args, cx.rank, cx.world, cx.dev
    name = "cfg3"
    full = CONFIGS[name]
    lo, hi = shard_bounds(full.B, rank, world)
    shape = Shape(hi - lo, full.T, full.n, full.p, full.m, full.K)
    ab = algorithmic_bytes(shape)
    steps = args.steps if args.steps > 0 else 20
    lanes = args.lanes
    case = make_case(shape, seed=10 + 97 * rank, mask_kind=args.mask)   # bernoulli(0.5) | block = "observe 4, hide 12" tiled
    g = to_dev(case, dev)
    pb = make_problem(g, shape, lanes)
    B, T, n, p, m, K = pb.shape
    e = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
    st = F.States(e(B, T, n, 1), e(B, T, n, n), e(B, T, n, 1), e(B, T, n, n), e(B, T, n, 1), e(B, T, n, n))
    Al, Bl, Cl = e(B, T, n, n), e(B, T, n, m), e(B, T, p, n)
    info = F.info_word(dev)
    info.zero_()
    inputs, states = pb.inputs(), st.c_struct()
    fwd = lambda i: capi.filter_smooth_fwd(pb.dims, inputs, states, Al, Bl, Cl, info, dev)
    lanes_used = capi.pick_lanes(pb.dims) if lanes == 0 else lanes
    if rank == 0:
        cx.sampler.start()
    for i in range(max(args.warmup, 3)):
        fwd(i)
    blocks = cx.timed_blocks(fwd, steps, blocks=3)
    ms_step = blocks[0]
    value = full.B * full.T / (ms_step * 1e-3)
    # bit-exact mask handling at full size
    miss = g["mask"] == 0
    ok_mu = bool(torch.equal(st.mus_filt[miss], st.mus_pred[miss]))
    sp = st.Sigmas_pred[miss]
    ok_sig = bool(torch.equal(st.Sigmas_filt[miss], 0.5 * (sp + sp.mT)))
    del sp, miss
    finite = bool(torch.isfinite(st.mus_smooth).all())
    code = int(info.item())
    clocks = cx.sampler.stop() if rank == 0 else None

    # end to end: pinned host Y / mask / alpha in, imputed observations a_t = C_t mu_{t|T} out (KVAE.impute, model.py:280-281)
    e2e = None
    if not args.no_e2e:
        del st, Al, Bl, Cl
        torch.cuda.empty_cache()
        dyn = PrecomputedWeights(case["A"], case["B"], case["C"], None, switching=False)
        kf = KalmanFilter(0.02 ** 0.5, 0.03 ** 0.5, case["mu0"], case["Sigma0"], dyn, lanes=lanes).to(dev)
        host = {k: case[k].pin_memory() for k in ("Y", "mask", "alpha")}
        out_host = torch.empty(B, T, p).pin_memory()
        dslots = {k: torch.empty_like(v, device=dev) for k, v in host.items()}

        def imp(i):
            for k, v in host.items():
                dslots[k].copy_(v, non_blocking=True)
            dyn.set_weights(dslots["alpha"])
            a_imp, _, _, _ = kf.impute_observations(dslots["Y"], None, dslots["mask"])
            out_host.copy_(a_imp, non_blocking=True)

        for i in range(2):
            imp(i)
        eb = cx.timed_blocks(imp, max(3, min(steps, 5)), blocks=3)
        ms_mono = statistics.median(eb)
        # the same call chunked over the batch (engine.ImputePipeline): upload of chunk c+1, launch of chunk c and
        # copy-back of chunk c-1 run on three streams
        from kalman_vae_b200.engine import ImputePipeline
        del dslots
        torch.cuda.empty_cache()
        pipe = ImputePipeline(kf, chunk=16384 if B >= 16384 else B, device=dev)
        piped = lambda i: pipe.run(host["Y"], host["mask"], alpha=host["alpha"], out_imputed=out_host)
        for i in range(2):
            piped(i)
        eb2 = cx.timed_blocks(piped, max(3, min(steps, 5)), blocks=3)
        ms_e = statistics.median(eb2)
        e2e = {"value": full.B * full.T / (ms_e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": sum(v.numel() * 4 for v in host.values()),
               "d2h_bytes_per_step": out_host.numel() * 4, "ms_per_step": ms_e, "ms_per_step_blocks": eb2,
               "api": "engine.ImputePipeline(kf).run(Y, mask, alpha): KalmanFilter.impute_observations (the Kalman part of "
                      "KVAE.impute) on pinned host Y, mask, alpha in chunks of 16384 sequences -- upload, launch and copy-back of "
                      "consecutive chunks overlap on three streams; the imputed observations [B,T,p] land in pinned host "
                      "memory; median of 3 blocks",
               "monolithic": {"ms_per_step": ms_mono, "ms_per_step_blocks": eb,
                              "api": "one KalmanFilter.impute_observations(Y, None, mask) call per step: upload, launch, copy-back "
                                     "one after the other"}}
    if rank == 0:
        kern = "k_seq_fwd" if lanes_used == 1 else "k_filter_smooth"
        achieved = ab["fwd"] * full.B * full.T / (ms_step * 1e-3) / 1e9 / world
        cpu = None if args.no_cpu else cpu_baseline_block(name, full, dev, gpu_context=False)
        emit({
            "metric": METRICS[name], "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "value_median_of_blocks": full.B * full.T / (statistics.median(blocks) * 1e-3),
            "ms_per_step_blocks": blocks,
            "config": {"workload": workload_text(name), "mask": args.mask, "lanes_per_sequence": lanes_used, "sequences_per_gpu": B,
                       "l2": f"inputs + outputs of one step = {(ab['fwd'] * B * T) / 2**30:.1f} GiB per GPU (far larger than L2)",
                       "sharding": "batch dimension, contiguous per rank; forward only: no collective", "collective": "none"},
            "checks": {"mask0_mu_filt_equals_mu_pred_bit_exact": ok_mu, "mask0_sigma_filt_equals_sym_sigma_pred_bit_exact": ok_sig,
                       "finite": finite, "status_word": code},
            "e2e": e2e, "gpu_launches": steps,
            "roofline": {"bound": "hbm", "kernel": kern, "achieved": achieved, "peak": cx.peak, "unit": "GB/s",
                         "frac": achieved / cx.peak, "traffic": None, "peak_source": cx.peak_src,
                         "algorithmic_bytes_per_seq_step": ab,
                         "note": "the step is one launch of this kernel: achieved = 440 B x sequence-steps / step time, per GPU"},
            "cpu_baseline": cpu, "clocks": clocks,
        })


# ----------------------------------------------------------------------------------------------------
# cfg5: full KVAE training step with the Kalman block swapped in
# ----------------------------------------------------------------------------------------------------
def run_kvae(cx: Ctx):
    from kalman_vae_b200 import kvae_step
    args, rank, world, dev, dist = cx.args, cx.rank, cx.world, cx.dev, cx.dist
    steps = args.steps if args.steps > 0 else 50
    res = {}
    if rank == 0:
        cx.sampler.start()
    for label, drop_in in (("drop_in_graphed", True), ("drop_in", True), ("reference_ops", False)):
        graphed = label == "drop_in_graphed"
        if graphed:   # the whole step captured in CUDA graphs (kvae_step.GraphedTrainStep)
            stepper = kvae_step.GraphedTrainStep(device=dev, seed=10 + rank, distributed=world > 1, dynamics_model=args.dynamics)
        else:
            stepper = kvae_step.ReferenceTrainStep(device=dev, drop_in=drop_in, seed=10 + rank, distributed=world > 1,
                                                   dynamics_model=args.dynamics)
        xs_host = [stepper.synthetic_batch(seed=100 * rank + i).pin_memory() for i in range(4)]
        if graphed:
            try:
                stepper.capture(xs_host[0])
            except Exception as err:   # a capture problem must not cost the line: the eager arm below becomes the value
                res[label] = {"failed": f"{type(err).__name__}: {err}"[:300]}
                torch.cuda.synchronize(dev)
                del stepper
                torch.cuda.empty_cache()
                continue
            out_host = torch.zeros(1).pin_memory()

            def run(i):   # per step: H2D of the frames into the static buffer, replay, D2H of the loss
                loss = stepper.step(xs_host[i % 4])
                out_host.copy_(loss.detach().reshape(1), non_blocking=True)
        else:
            run = lambda i: stepper.step(xs_host[i % 4].to(dev, non_blocking=True))
        n = steps if drop_in else max(3, min(steps, 10))
        for i in range(max(3, min(args.warmup, 5))):
            run(i)
        blocks = cx.timed_blocks(run, n, blocks=3 if drop_in else 1)
        res[label] = {"ms_per_step": statistics.median(blocks), "ms_per_step_blocks": blocks, "steps_per_block": n,
                      "loss_last": float(stepper.last_loss.detach())}
        B, T = stepper.batch, stepper.T
        h2d = xs_host[0].numel() * 4
        del stepper
        torch.cuda.empty_cache()
    clocks = cx.sampler.stop() if rank == 0 else None
    if rank == 0:
        best = "drop_in_graphed" if "ms_per_step" in res["drop_in_graphed"] else "drop_in"
        ms = res[best]["ms_per_step_blocks"][0]
        value = world * B * T / (ms * 1e-3)
        emit({
            "metric": METRICS["cfg5"], "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_text("cfg5"), "dynamics_model": args.dynamics,
                       "what": "reference KVAE module (conv encoder / decoder in PyTorch) with kalman_vae_b200.KalmanFilter and "
                               "DynamicsParameter swapped in; per step: H2D of the frames, forward, compute_loss, backward, NCCL "
                               "all-reduce of all gradients (N > 1), clip_grad_norm_(10), Adam(lr 0.007).  value = the step captured "
                               "in CUDA graphs (kvae_step.GraphedTrainStep: forward + loss + backward [+ eager NCCL all-reduce] + "
                               "clip + Adam replayed; compute_loss's two diagnostic host reads left out); kvae_step.drop_in = the "
                               "same step eager, kvae_step.reference_ops = eager with the reference's own Kalman ops",
                       "collective": "nccl all-reduce of the flattened gradients" if world > 1 else "none"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": ms,
                    "api": "the training-step body of kvae/train/train.py:32-58 on the reference KVAE with the drop-in Kalman block; "
                           "the batch of frames comes from pinned host memory every step and the loss is read back"},
            "kvae_step": res,
            "speedup_over_reference_ops_same_gpu": res["reference_ops"]["ms_per_step"] / res[best]["ms_per_step"],
            "speedup_over_reference_ops_same_gpu_eager": res["reference_ops"]["ms_per_step"] / res["drop_in"]["ms_per_step"],
            "gpu_launches": None, "roofline": None, "cpu_baseline": None, "clocks": clocks,
        })


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line of the contract, written to the process's real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # libraries print to fd 1 behind Python's back (NCCL's "NCCL version ..." banner at communicator creation):
    # send everything except the final JSON line to stderr
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0, help="timed steps (0 = the workload's default: cfg2 2000, cfg3 20, cfg4 10, cfg5 50)")
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--dynamics", default="lstm", choices=["lstm", "switching"],
                    help="cfg5: the dynamics parameter network of the KVAE (lstm = KVAE, switching = SKVAE with the regime sampler; "
                         "the reference's shipped config.yaml:44 selects switching, its KVAEConfig default too)")
    ap.add_argument("--mask", default="bernoulli", choices=["bernoulli", "block"],
                    help="cfg3: the missing-observation pattern (SURVEY 8d: Bernoulli(0.5) per (b,t), or the imputation "
                         "pattern 'observe 4, hide 12' of kvae/utils/imputation.py:4-25 tiled over T)")
    ap.add_argument("--buffer-sets", type=int, default=6)
    ap.add_argument("--no-graphs", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-throughput", action="store_true", help="skip the B=262144 throughput-regime measurement")
    ap.add_argument("--kernel-reps", type=int, default=200, help="launches per kernel for the roofline's per-kernel timing "
                    "(use a small number under ncu so that the launch list keeps the step's own proportions)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        if args.steps <= 0:
            args.steps = {"cfg2": 2000, "cfg3": 20, "cfg4": 10, "cfg5": 50}[args.workload]
        run_reference(args, rank, world)
        return
    cx = Ctx(args, rank, local_rank, world)
    try:
        if args.workload in ("cfg2", "cfg4"):
            run_train(cx, args.workload)
        elif args.workload == "cfg3":
            run_impute(cx)
        else:
            run_kvae(cx)
    finally:
        cx.finish()


if __name__ == "__main__":
    main()
