"""GPU: the exchanges of the path on an NCCL communicator INSIDE the library and the multi-GPU driver.

* one rank (runs on any GPU box): the communicator loads NCCL, the packed all-gather of the per-chain moments returns
  the Gelman-Rubin statistic of bamgpu_rhat on the same chains, the envelope gather returns the resident envelopes;
* two ranks, one process per GPU (needs 2 GPUs): chains shard with a global chain offset, prediction inputs shard as
  [t0, t1) slices; every rank ends with the statistic / the envelopes of the single-GPU run;
* `bam_gpu --gpus 2` (one process, ncclCommInitAll, one host thread per GPU) writes the files of the single-GPU run.
"""
import multiprocessing as mp
import os
import subprocess

import numpy as np
import pytest

from bam_b200.capi import BamGpu, Comm
from tests.helpers import load_fixture, problem_from_fixture
from tests.workspace_writer import write_workspace

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "bam_b200", "bam_gpu")


def gpu_count():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout
        return sum(1 for ln in out.splitlines() if ln.startswith("GPU "))
    except OSError:
        return 0


def start_std(start, K, rel=0.02):
    rng = np.random.default_rng(4)
    s = np.tile(np.asarray(start, dtype=np.float64), (K, 1))
    s *= 1.0 + rel * rng.standard_normal(s.shape)
    return s, 0.1 * np.abs(s)


def pred_case():
    prob, d = problem_from_fixture("baratin")
    rng = np.random.default_rng(2)
    mode = np.asarray(d["start"], dtype=np.float64)
    mc = mode * (1.0 + 0.02 * rng.standard_normal((400, mode.size)))
    return prob, mc, mode, np.linspace(-1, 6, 203)


def test_one_rank_communicator():
    prob, d = problem_from_fixture("baratin")
    K = 16
    s, sd = start_std(d["start"], K)
    g = BamGpu(prob)
    g.fit(s, sd, nAdapt=10, nCycles=4, seed=5, want_rows=False)
    comm = Comm(1, 0, Comm.unique_id(), 0)
    assert comm.nccl_version() >= 20000
    rhat, ms = g.moments_allgather(comm)
    cnt, mean, m2 = g.moments()
    assert (rhat == g.rhat(cnt, mean, m2)).all() and ms >= 0.0
    prob, mc, mode, H = pred_case()
    g2 = BamGpu(prob)
    g2.predict_upload(mc, mode, [H])
    g2.predict_resident(1, [1], seed=3)
    env = g2.predict_download()
    env_all, _ = g2.envelope_allgather(comm, H.size)
    assert (env_all == env).all()
    # sample-sharded selection with a single rank: the radix walk over (trivially) summed histograms
    g3 = BamGpu(prob)
    g3.predict_upload(mc, mode, [H])
    g3.predict_sample_sharded(comm, 0, mc.shape[0], 1, [1], seed=3)
    env_s = g3.predict_download()
    assert (env_s[:, :5] == env[:, :5]).all()  # order statistics are exact on every path
    sd = env[:, 6]
    assert np.allclose(env_s[:, 6], sd, rtol=1e-12) and (np.abs(env_s[:, 5] - env[:, 5]) <= 1e-12 * np.maximum(sd, np.abs(env[:, 5]))).all()
    comm.close()


def _rank(rank, world, uid, q):
    try:
        prob, d = problem_from_fixture("baratin")
        K = 16
        s, sd = start_std(d["start"], K)
        per = K // world
        g = BamGpu(prob, rank)
        g.set_chain_offset(rank * per)
        rows = g.fit(s[rank * per:(rank + 1) * per], sd[rank * per:(rank + 1) * per], nAdapt=10, nCycles=4, seed=5)
        comm = Comm(world, rank, uid, rank)
        rhat, _ = g.moments_allgather(comm)
        prob, mc, mode, H = pred_case()
        g2 = BamGpu(prob, rank)
        n = H.size
        pt = (n + world - 1) // world
        g2.predict_upload(mc, mode, [H])
        g2.predict_resident(1, [1], seed=3, t0=min(rank * pt, n), t1=min((rank + 1) * pt, n))
        env_all, _ = g2.envelope_allgather(comm, n)
        # the same prediction with the SAMPLES sharded instead: this rank's rows, every input
        half = mc.shape[0] // world
        lo, hi = rank * half, (mc.shape[0] if rank == world - 1 else (rank + 1) * half)
        mc_bad = mc.copy()
        mc_bad[37, 3] = -2.0  # an infeasible replication (k2 < k1) in rank 0's shard: dropped by the < 75 % rule
        g3 = BamGpu(prob, rank)
        g3.predict_upload(mc_bad[lo:hi], mode, [H])
        g3.predict_sample_sharded(comm, lo, mc.shape[0], 1, [1], seed=3)
        env_s = g3.predict_download()
        comm.close()
        q.put((rank, rows, rhat, env_all, None, env_s))
    except Exception as e:  # the parent must not wait for ever
        q.put((rank, None, None, None, repr(e), None))


@pytest.mark.skipif(gpu_count() < 2, reason="needs 2 GPUs")
def test_two_ranks_one_process_per_gpu():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    uid = Comm.unique_id()
    ps = [ctx.Process(target=_rank, args=(r, world, uid, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted((q.get(timeout=300) for _ in range(world)), key=lambda t: t[0])
    for p in ps:
        p.join(timeout=60)
    assert all(r[4] is None for r in res), [r[4] for r in res]
    # single-GPU run of the same experiment
    prob, d = problem_from_fixture("baratin")
    s, sd = start_std(d["start"], 16)
    g = BamGpu(prob)
    rows = g.fit(s, sd, nAdapt=10, nCycles=4, seed=5)
    cnt, mean, m2 = g.moments()
    rhat = g.rhat(cnt, mean, m2)
    assert (np.concatenate([res[0][1], res[1][1]], axis=0) == rows).all()  # sharding is invisible (global chain index)
    assert (res[0][2] == rhat).all() and (res[1][2] == rhat).all()
    prob, mc, mode, H = pred_case()
    g2 = BamGpu(prob)
    env, _ = g2.predict(mc, mode, [H], 1, [1], seed=3)
    assert (res[0][3] == env).all() and (res[1][3] == env).all()
    # sample-sharded: exact order statistics, mean / sd to rounding (the cross-rank order of the sums is NCCL's)
    mc_bad = mc.copy()
    mc_bad[37, 3] = -2.0
    envb, _ = g2.predict(mc_bad, mode, [H], 1, [1], seed=3)
    for r in range(world):
        es = res[r][5]
        assert (es[:, :5] == envb[:, :5]).all()
        sd = envb[:, 6]
        assert np.allclose(es[:, 6], sd, rtol=1e-12) and (np.abs(es[:, 5] - envb[:, 5]) <= 1e-12 * np.maximum(sd, np.abs(envb[:, 5]))).all()


def read_table(path):
    lines = open(path).read().splitlines()
    return np.array([[float(ln[i:i + 15]) for i in range(0, len(ln), 15)] for ln in lines[1:] if ln.strip()])


@pytest.mark.skipif(gpu_count() < 2, reason="needs 2 GPUs")
def test_driver_gpus2_writes_the_single_gpu_files(tmp_path):
    fx = load_fixture("baratin")
    outs = []
    for name, extra in (("ONE", []), ("TWO", ["--gpus", "2"])):
        cfg, ws = write_workspace(fx, str(tmp_path), name=name, nAdapt=20, nCycles=10, do_pred=False)
        r = subprocess.run([EXE, "-cf", cfg, "-sd", "7", "--chains", "8"] + extra, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        outs.append((read_table(os.path.join(ws, "Results_MCMC_Cooked.txt")), read_table(os.path.join(ws, "Results_Rhat.txt"))))
    assert (outs[0][0] == outs[1][0]).all()  # the pooled cooked sample of the 8 chains
    assert outs[0][1].shape == (1, 8) and np.allclose(outs[0][1], outs[1][1], rtol=1e-12)
