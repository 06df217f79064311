#!/usr/bin/env python
"""bench.py -- headline benchmark of the BaM hot path on B200 (contract: see the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ...]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1], the configuration the metric is quoted on): BaRatin_SPD multi-period
rating curves -- 4096 chains x 100 000 synthetic gaugings PER GPU (weak scaling: chains shard across
ranks, observations are replicated, no data-path collective).  A "step" = one batched log-posterior
evaluation of every chain of the rank against all observations.  metric = chain x obs evaluations / s,
whole job.  `value`: inputs resident in HBM, CUDA-event timed per step with an L2 flush between steps.
`e2e`: the same metric through the C-ABI call a host makes (bamgpu_logpost) with pinned HOST buffers,
H2D of the parameter vectors and D2H of (fx, flags) inside the timed region.
Also reported (same JSON line, key "prediction"): BASELINE's second metric, prediction samples x inputs / s
on configs[4] (BaRatin spaghetti with structural-error noise), scaled to fit the default run time.

--impl reference: the reference executable cannot be built here or on the GPU box (no Fortran compiler,
BMSL/miniDMSL un-vendored); this arm times the CPU restatement (oracle, kind "port") of the reference's own
algorithm on all host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "log-posterior evals/sec (chains x obs)"
UNIT = "chain*obs/s"
NOBS = 100_000
CHAINS_PER_GPU = 4096
# FP64-pipe instructions per chain x obs of logpost_main<BaRatin,1,1> on this workload, from the ncu capture in
# profiles/ (smsp__inst_executed_pipe_fp64 / units); see DESIGN.md "roofline accounting".
FP64_INST_PER_UNIT = None  # filled from profiles/fp64_inst_per_unit.json when present
NOMINAL_FLOPS_PER_UNIT = 24.6  # 3 range compares + 5 per active control (1.51 avg) + 14 likelihood tail
ALGO_BYTES_PER_OBS = 28  # H, Q, uQ (3 x 8 B) + int32 period


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


class ClockSampler:
    """SM clock / throttle-reason sampler that runs DURING the timed region.

    The timed region of the headline workload is tens of milliseconds, shorter than one `nvidia-smi -lms`
    period, so the sampling is done in-process through NVML (nvidia_ml_py) from a thread every ~2 ms; the
    nvidia-smi subprocess of the profiling recipe is the fallback when NVML cannot be loaded."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
               ("sw_power_cap", 0x4))

    def __init__(self, device_index: int):
        import threading

        self.rows = []
        self.stop_flag = threading.Event()
        self.thread = None
        self.proc = None
        self.path = None
        self.max_mhz = None
        self.source = None
        try:
            import pynvml

            pynvml.nvmlInit()
            h = None
            try:
                import torch

                uuid = str(torch.cuda.get_device_properties(device_index).uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

            def loop():
                while not self.stop_flag.is_set():
                    try:
                        sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                        rs = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                        pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
                        self.rows.append((time.perf_counter(), float(sm), int(rs), pw))
                    except Exception:
                        pass
                    time.sleep(0.002)

            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
            self.source = "nvml"
        except Exception:
            f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.path = f.name
            q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap")
            try:
                self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms",
                                              "20", "-i", str(device_index)], stdout=f, stderr=subprocess.DEVNULL)
                self.source = "nvidia-smi"
            except OSError:
                self.proc = None

    def n(self):
        if self.thread is not None:
            return len(self.rows)
        try:
            return sum(1 for _ in open(self.path))
        except Exception:
            return 0

    def stop(self, t_begin=None, t_end=None):
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "source": self.source}
        self.stop_flag.set()
        if self.thread is not None:
            self.thread.join(timeout=2)
            rows = [r for r in self.rows if (t_begin is None or r[0] >= t_begin) and (t_end is None or r[0] <= t_end)]
            out["samples_in_timed_region"] = len(rows)
            if not rows:
                rows = self.rows
            if rows:
                out["sm_mhz"] = float(np.median([r[1] for r in rows]))
                out["power_w_max"] = max(r[3] for r in rows)
                mask = 0
                for r in rows:
                    mask |= r[2]
                out["reasons"] = [nm for nm, bit in self.REASONS if mask & bit]
                out["samples"] = len(rows)
            return out
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        try:
            rows = [ln.strip().split(",") for ln in open(self.path) if ln.strip()]
            rows = [r for r in rows if len(r) >= 9]
            if rows:
                out["sm_mhz"] = float(np.median([float(r[1]) for r in rows]))
                out["sm_max_mhz"] = float(rows[0][2])
                out["power_w_max"] = max(float(r[3]) for r in rows)
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                for j, nm in enumerate(names):
                    if any("Active" in r[5 + j] and "Not" not in r[5 + j] for r in rows):
                        out["reasons"].append(nm)
                out["samples"] = len(rows)
        except Exception:
            pass
        finally:
            try:
                os.unlink(self.path)
            except OSError:
                pass
        return out


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def fp64_inst_per_unit():
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "fp64_inst_per_unit.json")))
    except Exception:
        return {}


def cpu_baseline(prob, x, seconds_target=12.0):
    """oracle (port) on all host cores on a bounded sample of the same workload"""
    from oracle.oracle_api import Oracle

    o = Oracle(prob)
    cores = os.cpu_count() or 1
    n = prob.nobs
    k = min(len(x), cores)
    t0 = time.perf_counter()
    o.logpost(x[:k], nthreads=cores)          # calibration: one vector per core
    dt = max(time.perf_counter() - t0, 1e-6)
    kk = int(min(len(x), max(cores, round(seconds_target / dt) * cores)))
    t0 = time.perf_counter()
    o.logpost(x[:kk], nthreads=cores)
    dt = time.perf_counter() - t0
    t1 = time.perf_counter()
    o.logpost(x[:2], nthreads=1)
    dt1 = time.perf_counter() - t1
    return {"value": kk * n / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{kk} chains x {n} obs, oracle/bam_oracle.c (C restatement; the Fortran reference cannot be built "
                      f"here), OpenMP over chains", "value_1core": 2 * n / dt1, "seconds": dt}


def headline_config(K, n, flush=True):
    """`config` of the JSON line: identical for both arms (the driver compares them)"""
    return {"workload": f"BaRatin_SPD (3 controls, 5 periods, VAR k1/k2, Gaussian-copula prior 19x19 ON): {K} chains x {n} "
                        f"gaugings per GPU (configs[1])",
            "chains_per_gpu": K, "nobs": n, "P": 19}


def run_reference(args):
    """The reference's own CPU implementation of the path, on all host cores, SAME batch as our arm: one step = the
    log-posterior of 4096 chains x 100 000 gaugings.  The Fortran executable cannot be built (no compiler, BMSL / miniDMSL
    un-vendored), so this is the C restatement (oracle/, -O3 -ffp-contract=off), kind "port"."""
    rank, world, local = dist_env()
    if rank != 0:
        return 0
    from workloads import synth
    from oracle.oracle_api import Oracle

    K, n = args.chains, args.nobs
    prob, truth = synth.baratin_spd(nobs=n, copula=True)
    cores = os.cpu_count() or 1
    x = synth.feasible_starts(truth, K, seed=synth.SEED + 1, check=synth.spd_start_check)
    o = Oracle(prob)
    for _ in range(max(args.warmup, 1)):
        o.logpost(x[: 4 * cores], nthreads=cores)  # warm-up: page in the tables, spin up the OpenMP team
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.logpost(x, nthreads=cores)
    dt = time.perf_counter() - t0
    val = args.steps * K * n / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": headline_config(K, n),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{K} chains x {n} obs per step (the full batch of our arm), oracle/bam_oracle.c "
                                       f"-O3 -ffp-contract=off, OpenMP over chains"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def prediction_e2e(g, lib, mc, mode, xspag, doRemnant, pairs, calls=3):
    """the prediction metric through the C-ABI call a host makes (bamgpu_predict): HOST sample matrix and inputs in,
    envelopes out, copies inside the timed region (wall clock around the blocking call)"""
    mcf = np.asfortranarray(mc)
    lib.bamgpu_pin(mcf.ctypes.data, mcf.nbytes)
    try:
        g.predict(mcf, mode, xspag, 1, doRemnant, seed=1)  # warm-up (allocations)
        dts = []
        for _ in range(calls):
            t0 = time.perf_counter()
            env, _ = g.predict(mcf, mode, xspag, 1, doRemnant, seed=1)
            dts.append(time.perf_counter() - t0)
        dt = float(np.median(dts))  # a single host hiccup (page-locked copy competing with another tenant) is not the rate
    finally:
        lib.bamgpu_unpin(mcf.ctypes.data)
    return {"value": pairs / dt, "unit": "sample*input/s", "ms_per_call": 1e3 * dt,
            "h2d_bytes_per_step": int(mcf.nbytes + sum(np.asarray(x).nbytes for x in xspag)), "d2h_bytes_per_step": int(env.nbytes),
            "ms_per_call_all": [1e3 * v for v in dts],
            "path": "bamgpu_predict (C ABI): pinned host sample matrix + inputs in, envelopes out, wall clock, median of the calls"}


def prediction_cpu_baseline(prob, mc, mode, xspag, doRemnant, nY_pred, seconds_target=10.0):
    """oracle (port) on all host cores, OpenMP over prediction inputs, on a bounded sample of the same experiment"""
    from oracle.oracle_api import Oracle

    o = Oracle(prob)
    cores = os.cpu_count() or 1
    nrep = min(mc.shape[0], 20_000)
    nobs = min(np.asarray(xspag[0]).shape[0], 64 * cores)
    xs = [np.asarray(x)[:nobs] for x in xspag]
    t0 = time.perf_counter()
    o.predict(mc[:nrep], mode, xs, 1, doRemnant, seed=1, nthreads=cores)
    dt = max(time.perf_counter() - t0, 1e-6)
    scale = int(max(1, min(np.asarray(xspag[0]).shape[0] // nobs, round(seconds_target / dt))))
    nobs2 = nobs * scale
    xs = [np.asarray(x)[:nobs2] for x in xspag]
    t0 = time.perf_counter()
    o.predict(mc[:nrep], mode, xs, 1, doRemnant, seed=1, nthreads=cores)
    dt = time.perf_counter() - t0
    return {"value": nrep * nobs2 * nY_pred / dt, "unit": "sample*input/s", "cores": cores, "kind": "port", "seconds": dt,
            "sample": f"{nrep} samples x {nobs2} inputs of the same experiment, oracle/bam_oracle.c (propagation + noise + "
                      f"sorted-column envelopes), OpenMP"}


def prediction_leg(device, fp64_peak, hbm_peak, steps=2):
    """BASELINE's second metric on configs[4] (BaRatin spaghetti + structural-error noise), bounded size.

    Two kernels dominate: the propagation kernel (FP64-pipe bound) and the envelope selection (HBM bound: it reads
    the materialised tile twice).  Both are timed with CUDA events on the launch stream and reported against their
    own roofline; the instruction count per (sample, input) comes from the committed ncu capture."""
    from workloads import synth
    from bam_b200.capi import BamGpu

    Nrep, nobs = 1_000_000, 100_000
    prob, _ = synth.baratin_2c(nobs=38)
    mc, mode, Hs = synth.baratin_spaghetti(Nrep=Nrep, nobs=nobs)
    g = BamGpu(prob, device)
    g.predict_upload(mc, mode, [Hs])
    g.predict_resident(1, [1], seed=1)
    g.sync()
    g.kernel_times()
    g.step_times()
    for _ in range(steps):
        g.step_begin()
        g.predict_resident(1, [1], seed=1)
        g.step_end()
    g.sync()
    ms = g.step_times()
    km = g.kernel_times()  # [propagation, envelopes] per step
    env = g.predict_download()
    ok = bool(np.isfinite(env).all())
    pairs = Nrep * nobs
    out = {"metric": "prediction samples x inputs/sec", "value": pairs / (ms.mean() * 1e-3), "unit": "sample*input/s",
           "ms_per_step": float(ms.mean()), "workload": f"BaRatin spaghetti {Nrep} samples x {nobs} inputs, parametric + "
           f"remnant noise, exact envelopes (configs[4], full size: {8 * Nrep * nobs / 1e9:.0f} GB of spaghetti materialised in "
           f"tiles of <= 48 GiB, far larger than L2)",
           "finite": ok}
    try:
        from bam_b200.capi import load_library
        out["e2e"] = prediction_e2e(g, load_library(), mc, mode, [Hs], [1], pairs)
        out["cpu_baseline"] = prediction_cpu_baseline(prob, mc, mode, [Hs], [1], 1)
    except Exception as e:
        out["e2e_error"] = str(e)
    if km.size >= 2 * steps:
        kp, ke = float(km[0::2].sum()) / steps, float(km[1::2].sum()) / steps  # summed over the tiles of a step
        ipu = fp64_inst_per_unit().get("pred_baratin_fast")
        out["kernels"] = {
            "propagation": {"kernel": "pred_baratin_fast<2,Linear,noise>", "ms": kp, "bound": "fp64",
                            "fp64_inst_per_unit": ipu,
                            "achieved": (2.0 * ipu * pairs / (kp * 1e-3) / 1e12) if ipu else None, "peak": fp64_peak,
                            "unit": "TFLOP/s", "frac": (2.0 * ipu * pairs / (kp * 1e-3) / 1e12 / fp64_peak) if ipu else None,
                            "hbm_write_gbs": 8.0 * pairs / (kp * 1e-3) / 1e9},
            "envelopes": {"kernel": "envelope_select3", "ms": ke, "bound": "hbm",
                          "algorithmic_bytes": 16 * pairs, "achieved": 16.0 * pairs / (ke * 1e-3) / 1e9, "peak": hbm_peak,
                          "unit": "GB/s", "frac": 16.0 * pairs / (ke * 1e-3) / 1e9 / hbm_peak},
        }
    return out


def sediment_prediction_leg(device, hbm_peak, Nrep=10_000_000, steps=2):
    """configs[3] at full size: Sediment Camenen + Soulsby (d, s as inputs), 10^7-sample posterior propagated through
    the 812 inputs of the reference workspace's shape, remnant noise on, exact envelopes."""
    from workloads import synth
    from bam_b200.capi import BamGpu

    nobs = 812
    prob, truth = synth.sediment_camenen(nobs=nobs)
    mc = synth.sediment_posterior(Nrep)
    X = [np.asarray(prob.X)[:, j].copy() for j in range(4)]
    g = BamGpu(prob, device)
    g.predict_upload(mc, truth, X)
    g.predict_resident(1, [1], seed=1)
    g.sync()
    g.step_times(); g.kernel_times()
    for _ in range(steps):
        g.step_begin()
        g.predict_resident(1, [1], seed=1)
        g.step_end()
    g.sync()
    ms = g.step_times()
    km = g.kernel_times()
    pairs = Nrep * nobs
    kp, ke = float(km[0::2].sum()) / steps, float(km[1::2].sum()) / steps
    extra = {}
    try:
        from bam_b200.capi import load_library
        extra["e2e"] = prediction_e2e(g, load_library(), mc, truth, X, [1], pairs)
        extra["cpu_baseline"] = prediction_cpu_baseline(prob, mc, truth, X, [1], 1)
    except Exception as e:
        extra["e2e_error"] = str(e)
    return {**extra, "metric": "prediction samples x inputs/sec", "value": pairs / (ms.mean() * 1e-3), "unit": "sample*input/s",
            "ms_per_step": float(ms.mean()),
            "workload": f"Sediment Camenen+Soulsby, {Nrep} samples x {nobs} inputs, parametric + remnant noise, exact envelopes "
                        f"(configs[3], full size; {8 * pairs / 1e9:.0f} GB of spaghetti materialised in tiles, inputs larger than L2)",
            "kernels": {"propagation": {"kernel": "pred_sediment_fast<Linear,noise>", "ms": kp,
                                        "hbm_write_gbs": 8.0 * pairs / (kp * 1e-3) / 1e9},
                        "envelopes": {"kernel": "envelope_select3<2048,1024>", "ms": ke, "bound": "hbm",
                                      "algorithmic_bytes": 16 * pairs, "achieved": 16.0 * pairs / (ke * 1e-3) / 1e9,
                                      "peak": hbm_peak, "unit": "GB/s", "frac": 16.0 * pairs / (ke * 1e-3) / 1e9 / hbm_peak}},
            "finite": bool(np.isfinite(g.predict_download()).all())}


def vegetation_prediction_leg(device, Nrep=10_000_000, nobs=1095, steps=1):
    """configs[3], second law: Vegetation (Growth_Yin3, 3 years of daily inputs, 5 outputs: Q, g, cos, sin, g0), posterior
    sample propagated through the per-observation Newton solve, remnant noise on Q, exact envelopes of all outputs.
    Generic kernels (pred_main, libdevice pow); the tie-heavy growth-index columns (exactly zero off season) exercise the
    4-read and the radix selection.  Full size: 10^7 samples."""
    from workloads import synth
    from bam_b200.capi import BamGpu

    prob, truth = synth.vegetation(nobs=nobs, growth="Growth_Yin3", nY=5)
    rng = np.random.default_rng(3)
    mc = rng.standard_normal((Nrep, truth.size))  # in place: 10^7 x P doubles once
    mc *= 0.002
    mc += 1.0
    mc *= truth
    X = [np.asarray(prob.X)[:, j].copy() for j in range(prob.nX)]
    g = BamGpu(prob, device)
    g.predict_upload(mc, truth, X)
    dor = [1, 0, 0, 0, 0]
    g.predict_resident(1, dor, seed=1)
    g.sync()
    g.step_times(); g.kernel_times()
    for _ in range(steps):
        g.step_begin()
        g.predict_resident(1, dor, seed=1)
        g.step_end()
    g.sync()
    ms, km = g.step_times(), g.kernel_times()
    pairs = Nrep * nobs
    return {"metric": "prediction samples x inputs/sec", "value": pairs / (ms.mean() * 1e-3), "unit": "sample*input/s",
            "ms_per_step": float(ms.mean()),
            "workload": f"Vegetation Growth_Yin3, {Nrep} samples x {nobs} inputs x 5 outputs, parametric + remnant noise on Q, "
                        f"exact envelopes ({8 * 5 * pairs / 1e9:.0f} GB of spaghetti materialised in tiles; configs[3], full size)",
            "kernels": {"propagation": {"kernel": "pred_main<Vegetation,4,5>", "ms": float(km[0::2].sum()) / steps},
                        "envelopes": {"kernel": "envelope_select3 / envelope_select2 / envelope_radix", "ms": float(km[1::2].sum()) / steps}},
            "finite": bool(np.isfinite(g.predict_download()).all())}


def small_sampler_leg(device, cpu=True):
    """configs[0]: the reference's own tests/BaM_BaRatin workspace (38 gaugings, P = 8), adaptive Metropolis with the
    persistent warp-per-chain kernel; K = 1 is like-for-like with the reference's single chain."""
    from bam_b200.capi import BamGpu
    from tests.helpers import problem_from_fixture

    prob, d = problem_from_fixture("baratin")
    out = {"workload": "tests/BaM_BaRatin (38 gaugings, 8 components), Config_MCMC shape: cycles of 100 iterations"}
    for K, nC in ((4096, 50), (1, 200)):
        g = BamGpu(prob, device)
        s = np.tile(np.asarray(d["start"]), (K, 1))
        sd = 0.1 * np.abs(s)
        g.fit(s, sd, nAdapt=10, nCycles=1, want_rows=False)
        g.step_times()
        g.fit(s, sd, nAdapt=100, nCycles=nC, seed=3, want_rows=False)
        ms = float(g.step_times()[-1])
        nprop = K * 100 * nC * g.P
        out[f"K{K}"] = {"iterations": 100 * nC, "ms": ms, "proposals_per_s": nprop / (ms * 1e-3),
                        "chain_obs_per_s": nprop * prob.nobs / (ms * 1e-3)}
    if cpu:  # the same single chain through the oracle on ONE host core (the reference's own mode of operation)
        import time
        from oracle.oracle_api import Oracle

        o = Oracle(prob)
        s = np.tile(np.asarray(d["start"]), (1, 1))
        sd = 0.1 * np.abs(s)
        t0 = time.perf_counter()
        o.fit(s, sd, nAdapt=100, nCycles=50, seed=3)
        dt = time.perf_counter() - t0
        v = 100 * 50 * o.P / dt
        out["cpu_1core"] = {"kind": "port", "proposals_per_s": v, "seconds": dt,
                            "note": "38 gaugings per evaluation: a single GPU chain is a serial latency chain and does NOT beat a "
                                    "CPU core here; the GPU's case is many chains (K4096) or many observations (small_k)"}
        out["K1"]["vs_cpu_1core"] = out["K1"]["proposals_per_s"] / v
        out["K4096"]["vs_cpu_1core"] = out["K4096"]["proposals_per_s"] / v
    return out


def small_k_leg(g, x, n, steps=10):
    """configs[1] problem (100 000 gaugings) with FEW chains: one resident log-posterior step at K = 1, 8, 64 (launch
    latency and tail bound; the headline configuration is K = 4096)."""
    out = {"workload": f"BaRatin_SPD log-posterior, {n} gaugings, K = 1 / 8 / 64 chains (resident, L2 flushed between steps)"}
    for K in (1, 8, 64):
        g.chains_upload(x[:K])
        for _ in range(3):
            g.logpost_resident()
        ms, km, pm, fm = timed_resident_logpost(g, steps)
        out[f"K{K}"] = {"ms_per_step": float(ms.mean()), "value": K * n / (float(ms.mean()) * 1e-3), "unit": UNIT}
    return out


def cfg2_sampler_leg(g, x, K, n, iters=5):
    """configs[1] as the user runs it: the batched adaptive-Metropolis sampler on the headline problem (4096 chains x
    100 000 gaugings, 19 components).  A proposal on a VAR parameter (k1 / k2 of one period) only re-evaluates the
    gaugings of that period; the per-combination sums of the current state cover the rest."""
    sd = 0.002 * np.abs(x) + 1e-6
    out = {"workload": f"BaRatin_SPD sampler, {K} chains x {n} gaugings, {g.P} components per iteration"}
    for name, flag in (("full_reevaluation", 1), ("partial_reevaluation", 0)):
        g.set_option("no_partial", flag)
        g.fit(x, sd, nAdapt=1, nCycles=1, seed=1, want_rows=False)
        g.step_times()
        g.fit(x, sd, nAdapt=iters, nCycles=1, seed=2, want_rows=False)
        ms = float(g.step_times()[-1]) / iters
        out[name] = {"ms_per_iteration": ms, "proposals_per_s": K * g.P / (ms * 1e-3),
                     "chain_obs_per_s_equivalent": K * g.P * n / (ms * 1e-3)}
    g.set_option("no_partial", 0)
    return out


def latent_logpost_leg(device, hbm_peak, n=1_000_000, K=64, steps=10):
    """configs[2] (Regression_X2Y2 shape with input uncertainty, 10^6 observations): log-posterior of K chains whose
    vectors each carry 2n latent "true" inputs.  The path must read 16 B of chain-private latent inputs per
    (chain, observation), so it is reported against the HBM roofline; L2 is flushed between timed steps."""
    from workloads import synth
    from bam_b200.capi import BamGpu

    prob, truth = synth.linear_latent(nobs=n)
    rng = np.random.default_rng(1)
    x = np.tile(truth, (K, 1))
    x[:, :8] *= 1.0 + 0.01 * rng.standard_normal((K, 8))
    x[:, 8:] += 0.05 * rng.standard_normal((K, 2 * n)).astype(np.float64)
    g = BamGpu(prob, device)
    g.chains_upload(x)
    for _ in range(3):
        g.logpost_resident()
    g.sync(); g.step_times(); g.kernel_times()
    for _ in range(steps):
        g.flush_l2()
        g.step_begin(); g.logpost_resident(); g.step_end()
    g.sync()
    ms, km = g.step_times(), g.kernel_times()
    fx, feas, _ = g.chains_download()
    k_ms = float(km.mean())
    algo = 16.0 * K * n + 72.0 * n  # latent inputs + the observation tables (X, 1/Xu, Y, Yu, hyper constant) once: L2 serves the chain tiles
    return {"metric": "log-posterior evals/sec (chains x obs)", "value": K * n / (float(ms.mean()) * 1e-3), "unit": "chain*obs/s",
            "workload": f"Linear 2x2, {n} obs, every input cell latent (P = {g.P}), {K} chains (configs[2] shape)",
            "ms_per_step": float(ms.mean()), "feasible": bool(feas.all()),
            "kernel": {"name": "logpost_linear_latent<2,2,Linear>", "ms": k_ms, "bound": "hbm", "algorithmic_bytes": algo,
                       "achieved": algo / (k_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                       "frac": algo / (k_ms * 1e-3) / 1e9 / hbm_peak}}


def other_models_leg(device, fp64_peak, hbm_peak, steps=10):
    """log-posterior of the other starred models at their BASELINE shapes: Sediment Camenen (812 rows of the reference's
    workspace shape) x 4096 chains, the TextFile twin of configs[2] (a*T+b with a latent T per observation, 10^6
    observations) x 64 chains, Vegetation (1095 daily rows) x 4096 chains.  Each with the roofline that bounds it: nominal
    flops per chain x obs / kernel time / DFMA peak, or algorithmic bytes / kernel time / HBM peak."""
    from workloads import synth
    from bam_b200.capi import BamGpu

    out = {}

    def run(name, prob, x, nominal_flops, algo_bytes_per_unit, kernel):
        g = BamGpu(prob, device)
        K, n = x.shape[0], prob.nobs
        g.chains_upload(x)
        for _ in range(3):
            g.logpost_resident()
        ms, km, pm, fm = timed_resident_logpost(g, steps)
        fx, feas, _ = g.chains_download()
        k_ms = float(km.mean())
        units = K * n
        e = {"workload": f"{name}: {K} chains x {n} observations", "value": units / (float(ms.mean()) * 1e-3), "unit": UNIT,
             "ms_per_step": float(ms.mean()), "feasible": bool(feas.all()),
             "breakdown_ms": {"parameter_table": float(pm.mean()), "likelihood_kernel": k_ms, "final_sum": float(fm.mean())},
             "kernel": {"name": kernel, "ms": k_ms, "nominal_flops_per_unit": nominal_flops,
                        "fp64": {"achieved": nominal_flops * units / (k_ms * 1e-3) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                                 "frac": nominal_flops * units / (k_ms * 1e-3) / 1e12 / fp64_peak},
                        "hbm": {"achieved": algo_bytes_per_unit * units / (k_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                "frac": algo_bytes_per_unit * units / (k_ms * 1e-3) / 1e9 / hbm_peak}}}
        out[name] = e

    try:
        prob, truth = synth.sediment_camenen(nobs=812)
        x = synth.feasible_starts(truth, 4096, rel=0.01, seed=synth.SEED + 21)
        run("sediment_camenen", prob, x, 27 + 14, 0.0, "logpost_sediment_cb<Linear> (per-row invariants hoisted)")
    except Exception as e:
        out["sediment_camenen"] = {"error": str(e)}
    try:
        prob, truth = synth.textfile_latent(nobs=1_000_000)
        rng = np.random.default_rng(2)
        x = np.tile(truth, (64, 1))
        x[:, :4] *= 1.0 + 0.01 * rng.standard_normal((64, 4))
        run("textfile_twin", prob, x, 2 + 14 + 7, 8.0, "logpost_tile<TextFile,1,1> (register-resident sums over observation tiles)")
    except Exception as e:
        out["textfile_twin"] = {"error": str(e)}
    try:
        prob, truth = synth.vegetation(nobs=1095, growth="Growth_Yin3", nY=5)
        x = synth.feasible_starts(truth, 4096, rel=0.002, seed=synth.SEED + 22)
        run("vegetation_yin3", prob, x, 60 + 5 * 14, 0.0, "logpost_chain<Vegetation,4,5> (thread = chain)")
    except Exception as e:
        out["vegetation_yin3"] = {"error": str(e)}
    return out


def latent_sampler_leg(device, n=1_000_000, K=64, iters=4):
    """configs[2] (Regression with input uncertainty): adaptive Metropolis over theta, gamma AND 2n latent inputs per
    chain; the latent inputs are swept by one kernel launch with O(1) local posterior differences.  Bounded size."""
    from workloads import synth
    from bam_b200.capi import BamGpu

    prob, truth = synth.linear_latent(nobs=n)
    g = BamGpu(prob, device)
    s = np.tile(truth, (K, 1))
    sd = np.tile(np.concatenate([0.01 * np.abs(truth[:8]), np.full(2 * n, 0.1)]), (K, 1))
    g.fit(s, sd, nAdapt=1, nCycles=1, seed=1, want_rows=False)
    g.step_times()
    g.fit(s, sd, nAdapt=iters, nCycles=1, seed=2, want_rows=False)
    per_iter = float(g.step_times()[-1]) * 1e-3 / iters  # CUDA events around the iteration loop (bamgpu_fit)
    P = g.P
    return {"workload": f"Linear 2x2, {n} obs, input uncertainty on every cell: P = {P} components per chain, {K} chains",
            "ms_per_iteration": 1e3 * per_iter, "component_updates_per_s": K * P / per_iter,
            "reference_equivalent_chain_obs_per_s": K * P * n / per_iter,
            "note": "one iteration = one-at-a-time update of all P components of every chain; the reference needs one "
                    "full log-posterior (n observations) per component, the sweep kernel needs O(1) per latent input"}


def timed_resident_logpost(g, steps, flush=True):
    """`steps` resident log-posterior evaluations, each bracketed by CUDA events, with the three kernels of a step timed
    separately: parameter table (priors, copula, continuity pre-pass), dominant kernel, fixed-order final sum."""
    g.sync(); g.step_times(); g.kernel_times(); g.phase_times()
    for _ in range(steps):
        if flush:
            g.flush_l2(256 << 20)
        g.step_begin(); g.logpost_resident(); g.step_end()
    g.sync()
    ms, km = g.step_times(), g.kernel_times()
    pm, fm = g.phase_times()
    return ms, km, pm, fm


def strong_scaling_leg(device, rank, world, dist, torch, total_chains, n, steps, prob, truth):
    """BASELINE configs[1] as written: 4096 chains x 100 000 gaugings SHARDED over the GPUs (total_chains / world per GPU,
    observations replicated, no collective on the data path).  value = total chain x obs / max-over-ranks step time; the
    per-kernel breakdown names what limits the curve."""
    from workloads import synth
    from bam_b200.capi import BamGpu

    Kloc = total_chains // world
    xall = synth.feasible_starts(truth, total_chains, seed=synth.SEED + 1, check=synth.spd_start_check)
    g = BamGpu(prob, device)
    g.set_chain_offset(rank * Kloc)
    g.chains_upload(xall[rank * Kloc:(rank + 1) * Kloc])
    for _ in range(3):
        g.logpost_resident()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms, km, pm, fm = timed_resident_logpost(g, steps)
    fx, feas, _ = g.chains_download()
    t = torch.tensor([float(ms.sum()), float(km.sum()), float(pm.sum()), float(fm.sum())], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tt = [float(v) / steps for v in t.tolist()]
    return {"scaling": "strong", "workload": f"BaRatin_SPD {total_chains} chains x {n} gaugings sharded over {world} GPU(s): "
            f"{Kloc} chains per GPU (configs[1] as written)", "chains_per_gpu": Kloc,
            "value": total_chains * n / (tt[0] * 1e-3), "unit": UNIT, "ms_per_step": tt[0],
            "breakdown_ms": {"parameter_table": tt[2], "likelihood_kernel": tt[1], "final_sum": tt[3],
                             "launch_gaps": max(tt[0] - tt[1] - tt[2] - tt[3], 0.0)},
            "feasible": bool(feas.all()), "timing": "CUDA events per step, L2 flushed between steps, max over ranks"}


def multi_gpu_legs(device, rank, world, dist, torch):
    """N > 1: the two places where the path exchanges data, through NCCL over NVLink INSIDE the
    library (bam_b200/csrc/bam_comm.cu; torch.distributed only carries the 128-byte communicator id between the ranks).
    (1) prediction inputs shard across ranks ([t0,t1) slices, samples replicated); every rank owns whole columns, so
        the exact quantiles need no exchange; the envelope slices are all-gathered (bamgpu_envelope_allgather).
        STRONG scaling: the experiment is fixed (10^6 samples x 10^5 inputs), time = max over ranks (CUDA events) + the
        gather.
    (2) chains shard across ranks; per-chain moments {count, mean, M2} are packed into one message, all-gathered and
        the Gelman-Rubin statistic is computed redundantly on every rank (bamgpu_moments_allgather)."""
    from workloads import synth
    from bam_b200.capi import BamGpu, Comm
    from tests.helpers import problem_from_fixture

    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(Comm.unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    comm = Comm(world, rank, bytes(uid.cpu().numpy().tobytes()), device)
    out = {"nccl_version": comm.nccl_version()}
    # ---- (1) sharded prediction
    Nrep, nobs = 1_000_000, 100_000
    prob, _ = synth.baratin_2c(nobs=38)
    mc, mode, Hs = synth.baratin_spaghetti(Nrep=Nrep, nobs=nobs)
    per = (nobs + world - 1) // world
    t0, t1 = min(rank * per, nobs), min(nobs, (rank + 1) * per)
    g = BamGpu(prob, device)
    g.predict_upload(mc, mode, [Hs])
    g.predict_resident(1, [1], seed=1, t0=t0, t1=t1)
    g.envelope_allgather(comm, nobs)  # warm-up: NCCL channels, buffers
    g.sync(); g.step_times(); g.kernel_times()
    dist.barrier(); torch.cuda.synchronize()
    g.step_begin()
    g.predict_resident(1, [1], seed=1, t0=t0, t1=t1)
    g.step_end()
    g.sync()
    ms = float(g.step_times()[-1])
    full, ms_gather = g.envelope_allgather(comm, nobs)
    tt = torch.tensor([ms, ms_gather], device="cuda", dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    out["prediction"] = {"metric": "prediction samples x inputs/sec", "value": Nrep * nobs / ((tt[0].item() + tt[1].item()) * 1e-3),
                         "unit": "sample*input/s", "scaling": "strong", "ms_compute_max_over_ranks": tt[0].item(),
                         "ms_envelope_allgather_nccl": tt[1].item(),
                         "workload": f"BaRatin spaghetti {Nrep} samples x {nobs} inputs sharded by input over {world} GPUs",
                         "finite": bool(np.isfinite(full).all()), "median_checksum": float(full[0, 0].sum())}
    del g
    # ---- (1b) configs[3]: Sediment, 10^7 samples x 812 inputs, inputs sharded (101 columns per GPU at N = 8)
    try:
        Nrep, nobs = 10_000_000, 812
        prob, truth = synth.sediment_camenen(nobs=nobs)
        mc = synth.sediment_posterior(Nrep)
        X = [np.asarray(prob.X)[:, j].copy() for j in range(4)]
        per = (nobs + world - 1) // world
        t0, t1 = min(rank * per, nobs), min(nobs, (rank + 1) * per)
        g = BamGpu(prob, device)
        g.predict_upload(mc, truth, X)
        g.predict_resident(1, [1], seed=1, t0=t0, t1=t1)
        g.envelope_allgather(comm, nobs)
        g.sync(); g.step_times(); g.kernel_times()
        dist.barrier(); torch.cuda.synchronize()
        g.step_begin()
        g.predict_resident(1, [1], seed=1, t0=t0, t1=t1)
        g.step_end()
        g.sync()
        ms = float(g.step_times()[-1])
        full, ms_gather = g.envelope_allgather(comm, nobs)
        tt = torch.tensor([ms, ms_gather], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        out["prediction_sediment"] = {"metric": "prediction samples x inputs/sec", "value": Nrep * nobs / ((tt[0].item() + tt[1].item()) * 1e-3),
                                      "unit": "sample*input/s", "scaling": "strong", "ms_compute_max_over_ranks": tt[0].item(),
                                      "ms_envelope_allgather_nccl": tt[1].item(),
                                      "workload": f"Sediment Camenen {Nrep} samples x {nobs} inputs sharded by input over {world} GPUs "
                                                  f"({per} columns per GPU)", "finite": bool(np.isfinite(full).all()),
                                      "median_checksum": float(full[0, 0].sum())}
        del g, mc
    except Exception as e:
        out["prediction_sediment"] = {"error": str(e)}
    # ---- (1c) few inputs, many samples: SAMPLES sharded, exact quantiles from digit histograms summed over the ranks
    try:
        Nrep, nobs = 8_000_000, 181  # the 181-point stage grid of tests/BaM_BaRatin (Hgrid.txt)
        prob, _ = synth.baratin_2c(nobs=38)
        mc, mode, _ = synth.baratin_spaghetti(Nrep=Nrep, nobs=8)
        Hgrid = np.linspace(-1.0, 8.0, nobs)
        per = Nrep // world
        lo, hi = rank * per, (Nrep if rank == world - 1 else (rank + 1) * per)
        g = BamGpu(prob, device)
        g.predict_upload(mc[lo:hi], mode, [Hgrid])
        g.predict_sample_sharded(comm, lo, Nrep, 1, [1], seed=1)  # warm-up
        g.sync(); g.step_times(); g.kernel_times()
        dist.barrier(); torch.cuda.synchronize()
        g.step_begin()
        g.predict_sample_sharded(comm, lo, Nrep, 1, [1], seed=1)
        g.step_end()
        g.sync()
        tt = torch.tensor([float(g.step_times()[-1])], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        env = g.predict_download()
        out["prediction_sample_sharded"] = {"metric": "prediction samples x inputs/sec", "value": Nrep * nobs / (tt[0].item() * 1e-3),
                                            "unit": "sample*input/s", "scaling": "strong", "ms_max_over_ranks": tt[0].item(),
                                            "workload": f"BaRatin {Nrep} samples x {nobs}-point stage grid, SAMPLES sharded over {world} GPUs "
                                                        f"(6 all-reduces of digit histograms + 3 of sums inside the timed region)",
                                            "finite": bool(np.isfinite(env).all()), "median_checksum": float(env[0, 0].sum())}
        del g, mc
    except Exception as e:
        out["prediction_sample_sharded"] = {"error": str(e)}
    # ---- (2) sharded sampler + NCCL all-gather of the moments
    prob, d = problem_from_fixture("baratin")
    K = 4096
    g = BamGpu(prob, device)
    g.set_chain_offset(rank * K)
    s0 = np.tile(np.asarray(d["start"]), (K, 1))
    g.fit(s0, 0.1 * np.abs(s0), nAdapt=100, nCycles=20, seed=3, want_rows=False)
    g.moments_allgather(comm)  # warm-up
    dist.barrier(); torch.cuda.synchronize()
    rhat, ms_m = g.moments_allgather(comm)
    tm = torch.tensor([ms_m], device="cuda", dtype=torch.float64)
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    out["sampler"] = {"workload": f"tests/BaM_BaRatin, {K} chains per GPU x 2000 iterations, chains sharded over {world} GPUs",
                      "ms_fit_loop": float(g.step_times()[-1]), "ms_moments_allgather_nccl": tm.item(),
                      "chains_total": world * K, "rhat_max": float(np.nanmax(rhat))}
    comm.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chains", type=int, default=CHAINS_PER_GPU)
    ap.add_argument("--nobs", type=int, default=NOBS)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-prediction", action="store_true")
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling leg")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="which leg is the headline `value` (the other one is reported beside it)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    rank, world, local = dist_env()
    import torch  # plumbing only: barrier / max-over-ranks / synchronize

    if world > 1:
        import torch.distributed as dist

        # NCCL prints its version banner on STDOUT; the contract is ONE JSON line there -> park fd 1 on stderr
        # while the communicator is created (first collective), then restore it.
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    device = local if world > 1 else 0
    torch.cuda.set_device(device)

    from workloads import synth
    from bam_b200.capi import BamGpu, load_library

    lib = load_library()
    K, n = args.chains, args.nobs
    prob, truth = synth.baratin_spd(nobs=n, copula=True)
    x = synth.feasible_starts(truth, K, seed=synth.SEED + 1 + rank, check=synth.spd_start_check)
    g = BamGpu(prob, device)
    g.set_chain_offset(rank * K)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        g.sync()

    # ---- resident path (value) ----
    g.chains_upload(x)
    for _ in range(args.warmup):
        g.logpost_resident()
    g.sync()
    g.step_times(); g.kernel_times()
    launches0 = g.launch_count()
    sampler = ClockSampler(device)
    # keep the same load running until the sampler has seen it (clock ramp), then time K steps
    t_ramp = time.perf_counter()
    while sampler.n() < 20 and time.perf_counter() - t_ramp < 2.0:
        g.logpost_resident()
        g.sync()
    g.step_times(); g.kernel_times(); g.phase_times()
    launches0 = g.launch_count()
    barrier()
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        if not args.no_flush:
            g.flush_l2(256 << 20)
        g.step_begin()
        g.logpost_resident()
        g.step_end()
    barrier()
    wall1 = time.perf_counter()
    wall = wall1 - wall0
    clocks = sampler.stop(wall0, wall1)
    step_ms = g.step_times()
    kern_ms = g.kernel_times()
    launches = g.launch_count() - launches0
    fx, feas, isnull = g.chains_download()
    assert feas.all() and np.isfinite(fx).all(), "benchmark vectors must be feasible"
    t_total = float(step_ms.sum()) * 1e-3
    if world > 1:
        tt = torch.tensor([t_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_total = float(tt.item())
    value = world * K * n * args.steps / t_total

    pm_ms, fm_ms = g.phase_times()
    # ---- configs[1] as written: 4096 chains in total, sharded (strong scaling) ----
    strong = None
    if args.chains % world == 0 and not args.no_strong:
        try:
            strong = strong_scaling_leg(device, rank, world, dist if world > 1 else None, torch, args.chains, n,
                                        max(args.steps, 10), prob, truth)
        except Exception as e:
            strong = {"error": str(e)}

    # ---- e2e path: host buffers through the C ABI ----
    xh = np.ascontiguousarray(x)
    lib.bamgpu_pin(xh.ctypes.data, xh.nbytes)
    g.logpost(xh)
    barrier()
    e0 = time.perf_counter()
    esteps = max(3, min(args.steps, 10))
    for _ in range(esteps):
        fx2, feas2, isnull2 = g.logpost(xh)
    barrier()
    e_t = time.perf_counter() - e0
    lib.bamgpu_unpin(xh.ctypes.data)
    if world > 1:
        tt = torch.tensor([e_t], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e_t = float(tt.item())
    e2e_value = world * K * n * esteps / e_t
    assert (fx2 == fx).all()

    multi = None
    if world > 1 and not args.no_prediction:
        try:
            multi = multi_gpu_legs(device, rank, world, dist, torch)
        except Exception as e:  # every rank must reach the barriers below
            multi = {"error": str(e)}
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (logpost_main), rank 0 ----
    peaks = measured_peaks()
    tfl = np.zeros(1)
    import ctypes as C
    mb = C.create_string_buffer(250)
    lib.bamgpu_fp64_peak(device, tfl.ctypes.data_as(C.POINTER(C.c_double)), mb)
    fp64_peak = float(tfl[0])
    k_ms = float(kern_ms.mean())
    units = K * n
    ipu = fp64_inst_per_unit().get("logpost_main_baratin_spd")
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    # algorithmic bytes of one launch: every observation table once + the per-chain table + one (lkh, hyper) partial
    # per (observation tile, chain)
    n_tiles = (n + 255) // 256
    algo_bytes = n * ALGO_BYTES_PER_OBS + K * 8 * (2 + 5 * 12) + n_tiles * K * 16
    ach = NOMINAL_FLOPS_PER_UNIT * units / (k_ms * 1e-3) / 1e12
    roof = {
        "bound": "fp64",
        "kernel": "logpost_baratin_cb2<3,Linear> (BaRatin fast path, generation 3.5: Taylor sub-tiles)",
        "unit": "TFLOP/s",
        # ALGORITHMIC (nominal) flops per chain x obs -- +,-,*,/,compare and each pow/log/exp count 1 --
        # x units per launch / the kernel's CUDA-event time; independent of how the kernel is written
        "achieved": ach,
        "peak": fp64_peak,
        "frac": ach / fp64_peak,
        "algorithmic_flops_per_unit": NOMINAL_FLOPS_PER_UNIT,
        "peak_source": "builder-measured: DFMA-chain micro-benchmark (bamgpu_fp64_peak) run live in this process; "
                       "MEASURED_PEAKS.json has no FP64 figure; 2 flop per DFMA; nominal 148 SM x 64 lanes x 2 x 1.965 GHz = 37.2",
        "kernel_ms": k_ms,
        "kernel_share_of_step": k_ms / float(step_ms.mean()),
        "traffic": None,
        "hbm": {"achieved": algo_bytes / (k_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": algo_bytes / (k_ms * 1e-3) / 1e9 / hbm_peak, "algorithmic_bytes": algo_bytes,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback"},
    }
    prof = fp64_inst_per_unit()
    if ipu:
        # NOT the roofline fraction: how busy the FP64 pipe is with the instructions THIS kernel executes (ncu count)
        roof["fp64_pipe_util"] = 2.0 * ipu * units / (k_ms * 1e-3) / 1e12 / fp64_peak
        roof["fp64_inst_per_unit_executed"] = ipu
        roof["traffic"] = prof.get("dram_bytes_per_launch")
    # implementation-independent cost model at the accuracy the parity bar needs (DESIGN.md "roofline accounting"):
    # DFMA-equivalent instructions of a minimal table-driven evaluation, 2 flop each
    mde = prof.get("min_dfma_equiv_per_unit")
    if mde:
        roof["min_dfma_equiv_per_unit"] = mde
        roof["frac_dfma_equiv"] = 2.0 * mde * units / (k_ms * 1e-3) / 1e12 / fp64_peak

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": headline_config(K, n),
        "l2": "flushed between timed steps (256 MiB write)" if not args.no_flush else "not flushed",
        "timing": "sum of per-step CUDA-event intervals on the launch stream, max over ranks",
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(xh.nbytes),
                "d2h_bytes_per_step": int(K * (8 + 4)), "steps": esteps,
                "path": "bamgpu_logpost (C ABI) with pinned host buffers, wall clock around the blocking calls"},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "wall_s_timed_region": wall,
    }
    line["breakdown_ms"] = {"parameter_table": float(pm_ms.mean()) if pm_ms.size else None, "likelihood_kernel": k_ms,
                            "final_sum": float(fm_ms.mean()) if fm_ms.size else None}
    if strong is not None:
        line["strong"] = strong
        if args.scaling == "strong" and "value" in strong:  # the sharded configuration as the headline, the weak one beside it
            line["weak"] = {"value": line["value"], "ms_per_step": line["ms_per_step"], "chains_per_gpu": K}
            line.update({"value": strong["value"], "ms_per_step": strong["ms_per_step"], "scaling": "strong"})
    if multi is not None:
        line["multi_gpu"] = multi
    if not args.no_prediction and world == 1:
        try:
            line["prediction"] = prediction_leg(device, fp64_peak, hbm_peak)
        except Exception as e:  # the headline metric must still be printed
            line["prediction"] = {"error": str(e)}
        try:
            line["prediction_sediment"] = sediment_prediction_leg(device, hbm_peak)
        except Exception as e:
            line["prediction_sediment"] = {"error": str(e)}
        try:
            line["prediction_vegetation"] = vegetation_prediction_leg(device)
        except Exception as e:
            line["prediction_vegetation"] = {"error": str(e)}
        try:
            line["sampler_cfg2"] = cfg2_sampler_leg(g, x, K, n)
        except Exception as e:
            line["sampler_cfg2"] = {"error": str(e)}
        try:
            line["sampler_small"] = small_sampler_leg(device, cpu=not args.no_cpu)
        except Exception as e:
            line["sampler_small"] = {"error": str(e)}
        try:
            line["small_k"] = small_k_leg(g, x, n)
            g.chains_upload(x)
        except Exception as e:
            line["small_k"] = {"error": str(e)}
        try:
            line["latent_logpost"] = latent_logpost_leg(device, hbm_peak)
        except Exception as e:
            line["latent_logpost"] = {"error": str(e)}
        try:
            line["latent_sampler"] = latent_sampler_leg(device)
        except Exception as e:
            line["latent_sampler"] = {"error": str(e)}
        try:
            line["logpost_other_models"] = other_models_leg(device, fp64_peak, hbm_peak)
        except Exception as e:
            line["logpost_other_models"] = {"error": str(e)}
    if not args.no_cpu and world == 1:  # last: the OpenMP team of the oracle must not compete with the launch thread above
        line["cpu_baseline"] = cpu_baseline(prob, x)
        v1 = line["cpu_baseline"].get("value_1core")
        if v1 and isinstance(line.get("small_k"), dict):
            for kk, e in line["small_k"].items():
                if isinstance(e, dict) and "value" in e:
                    e["vs_cpu_1core"] = e["value"] / v1
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
