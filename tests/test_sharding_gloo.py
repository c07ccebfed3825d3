"""world_size-2 check of the rollout-sharding arithmetic on CPU (gloo): each rank builds the record its GPU would
produce for its half of the rollouts — (beta_r, eta_r, sum w^2_r, V_r) against its LOCAL baseline — the records are
all-gathered, and the merge (product code: mppib_host_merge_records, the CPU twin of K2) must reproduce the
single-process solve. Costs/samples come from the oracle; this test covers the N>1 host logic: slice bounds, global-index
special cases (sample 0, pure-noise tail), the log-sum-exp merge and the rendezvous plumbing bench.py uses."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_record(costs, samples, lam, TC):
    """What K1 + the non-normalising K2 emit for one rank: local baseline, eta, sum w^2, V = sum_n w_n u_n."""
    D = costs.shape[0]
    rec = np.zeros((D, 4 + TC), np.float32)
    for d in range(D):
        c = costs[d].astype(np.float32)
        beta = c.min()
        w = np.exp(np.float32(-1.0 / lam) * (c - beta)).astype(np.float32)
        rec[d, 0], rec[d, 1], rec[d, 2] = beta, w.astype(np.float64).sum(), (w.astype(np.float64) ** 2).sum()
        rec[d, 4:] = np.einsum("n,nk->k", w.astype(np.float64), samples[d].reshape(len(w), TC).astype(np.float64))
    return rec


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, ROOT)
    import oracle
    import mppi_generic_b200 as m
    from mppi_generic_b200 import workloads as W
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = W.double_integrator_tube(512, 24)
    w.x0[1] = [1.9, 0.1, 0.05, 1.1]
    w.sampler.setControlCostCoeff([0.3, 0.2])
    C_ = w.dyn.CONTROL_DIM
    TC = w.T * C_
    eps = oracle.curand_normal(5, 0, w.N * TC).reshape(w.N, w.T, C_)
    full = oracle.solve(w.dyn.DYN_ID, w.cost.COST_ID, w.dyn.params, w.cost.params, w.sampler.params, None, None, w.N,
                        w.T, w.D, C_, w.dt, w.lambda_, w.alpha, w.x0, w.U0, eps, want_samples=True)
    per = w.N // world
    lo, hi = rank * per, (w.N if rank == world - 1 else (rank + 1) * per)
    rec = _rank_record(full["costs"][:, lo:hi], full["samples"][:, lo:hi], w.lambda_, TC)
    t = torch.from_numpy(rec)
    gathered = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    records = np.stack([g.numpy() for g in gathered])  # [world][D][pstride]
    merged = m.host.merge_records(records, w.lambda_, normalize=True)
    ok = True
    for d in range(w.D):
        ok &= bool(merged[d, 0] == np.float32(full["baseline"][d]))
        ok &= bool(abs(merged[d, 1] - full["normalizer"][d]) <= 2e-6 * full["normalizer"][d])
        ok &= bool(np.allclose(merged[d, 4:].reshape(w.T, C_), full["U"][d], atol=2e-5, rtol=1e-5))
    # the unique-id style broadcast bench.py uses
    ids = [b"x" * 128 if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ok &= ids[0] == b"x" * 128
    q.put((rank, ok, float(merged[0, 0])))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_two_rank_sharded_merge_matches_single_process(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert len({b for _, _, b in res}) == 1  # every rank ends with the same baseline
