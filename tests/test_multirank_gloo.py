"""N > 1 host-side logic on CPU: two ranks over gloo (world_size 2), as bench.py / the driver shard the work.

* chains shard across ranks with a global chain offset -> every rank reproduces exactly the rows a single
  process would produce for its chains (the RNG streams are keyed by the GLOBAL chain index);
* per-chain moments are all-gathered and the Gelman-Rubin statistic computed redundantly on every rank
  (bamgpu_rhat, a host function of the C ABI) equals the single-process value;
* prediction inputs shard as [t0, t1) slices -> concatenated envelopes equal the single-process ones.
The compute on each rank is the ORACLE here (no GPU in this container); the GPU path is checked by
tests/test_gpu_sampler.py::test_chain_offset_makes_sharding_invisible and
tests/test_gpu_predict.py::test_input_sharding_is_invisible.
"""
import ctypes
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bam_b200.capi import _ptr_d, load_library
    from oracle.oracle_api import Oracle
    from tests.helpers import problem_from_fixture

    prob, d = problem_from_fixture("baratin")
    o = Oracle(prob)
    K, P = 8, o.P
    start = np.tile(np.asarray(d["start"]), (K, 1))
    sd = 0.1 * np.abs(start)
    kloc = K // world
    sl = slice(rank * kloc, (rank + 1) * kloc)
    rows, _, _ = o.fit(start[sl], sd[sl], nAdapt=10, nCycles=3, seed=5, chain_offset=rank * kloc)
    x = rows[:, :, :P]
    cnt = torch.full((kloc,), float(x.shape[1]), dtype=torch.float64)
    mean = torch.from_numpy(x.mean(axis=1))
    m2 = torch.from_numpy(((x - x.mean(axis=1, keepdims=True)) ** 2).sum(axis=1))
    g_cnt = torch.empty(K, dtype=torch.float64)
    g_mean = torch.empty(K, P, dtype=torch.float64)
    g_m2 = torch.empty(K, P, dtype=torch.float64)
    dist.all_gather_into_tensor(g_cnt, cnt)
    dist.all_gather_into_tensor(g_mean, mean.contiguous())
    dist.all_gather_into_tensor(g_m2, m2.contiguous())
    lib = load_library()
    rhat = np.empty(P)
    mess = ctypes.create_string_buffer(250)
    assert lib.bamgpu_rhat(K, P, _ptr_d(g_cnt.numpy()), _ptr_d(np.ascontiguousarray(g_mean.numpy())),
                           _ptr_d(np.ascontiguousarray(g_m2.numpy())), _ptr_d(rhat), mess) == 0
    # prediction: shard the inputs
    H = np.linspace(-1, 6, 40)
    mc = np.asarray(d["start"]) * (1 + 0.02 * np.random.default_rng(1).standard_normal((64, P)))
    mc[:, 6:] = np.abs(mc[:, 6:])
    nT = len(H) // world
    env, _ = o.predict(mc, mc[0], [H], 1, [1], seed=3, t0=rank * nT, t1=(rank + 1) * nT)
    g_env = [torch.empty_like(torch.from_numpy(env)) for _ in range(world)]
    dist.all_gather(g_env, torch.from_numpy(env))
    # max over ranks of a timing, as bench.py does
    tt = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0:
        np.savez(out, rows=rows, rhat=rhat, env=np.concatenate([e.numpy() for e in g_env], axis=2), tmax=tt.numpy())
    else:
        np.savez(out + ".r1", rows=rows)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gloo(tmp_path):
    from bam_b200 import build

    build.build_library()
    out = str(tmp_path / "rank0.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r0 = np.load(out)
    r1 = np.load(out + ".r1.npz")
    sys.path.insert(0, ROOT)
    from bam_b200.capi import _ptr_d, load_library
    from oracle.oracle_api import Oracle
    from tests.helpers import problem_from_fixture

    prob, d = problem_from_fixture("baratin")
    o = Oracle(prob)
    K, P = 8, o.P
    start = np.tile(np.asarray(d["start"]), (K, 1))
    rows, _, _ = o.fit(start, 0.1 * np.abs(start), nAdapt=10, nCycles=3, seed=5)
    assert (rows[:4] == r0["rows"]).all() and (rows[4:] == r1["rows"]).all()  # sharding is invisible
    x = rows[:, :, :P]
    lib = load_library()
    rhat = np.empty(P)
    mess = ctypes.create_string_buffer(250)
    cnt = np.full(K, float(x.shape[1]))
    mean = np.ascontiguousarray(x.mean(axis=1))
    m2 = np.ascontiguousarray(((x - x.mean(axis=1, keepdims=True)) ** 2).sum(axis=1))
    lib.bamgpu_rhat(K, P, _ptr_d(cnt), _ptr_d(mean), _ptr_d(m2), _ptr_d(rhat), mess)
    assert np.allclose(rhat, r0["rhat"], rtol=1e-12, equal_nan=True)
    H = np.linspace(-1, 6, 40)
    mc = np.asarray(d["start"]) * (1 + 0.02 * np.random.default_rng(1).standard_normal((64, P)))
    mc[:, 6:] = np.abs(mc[:, 6:])
    env, _ = o.predict(mc, mc[0], [H], 1, [1], seed=3)
    assert (env == r0["env"]).all()
    assert r0["tmax"][0] == 2.0
