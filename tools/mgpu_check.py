"""Launched under torchrun (one rank per GPU): the rollout-sharded solve must reproduce the single-GPU solve.
Rank 0 prints 'MGPU_OK' on success. Used by tests/test_gpu_multi.py and by hand under `gpurun --gpus N`."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mppi_generic_b200 as m  # noqa: E402
from mppi_generic_b200 import workloads as W  # noqa: E402

H = m.host


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for name, N, T in (("cartpole", 8192, 100), ("double_integrator_tube", 16384, 150), ("autorally", 8192, 50)):
        w = W.by_name(name, N, T)
        rng = np.random.RandomState(0)
        w.U0 = rng.uniform(-0.3, 0.3, w.U0.shape).astype(np.float32)
        e = H.Engine(w.dyn, w.cost, w.sampler, w.N, w.T, w.D, device=local, rank=rank, world_size=world)
        e.set_solver(w.dt, w.lambda_, w.alpha)
        e.seed(w.seed, 0)
        ids = [H.Engine.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        e.comm_init(ids[0])
        if os.environ.get("MPPIB_NO_P2P") is None:
            e.p2p_setup(dist)  # peer-memory exchange on every rank, or NCCL on every rank
        outs = []
        # open loop: the same three nominal sequences go to the sharded and to the single-GPU engine, so the comparison
        # sees one solve's reduction-order differences and not their closed-loop amplification
        U_ins = [w.U0.copy(), (0.5 * w.U0 + 0.05).astype(np.float32), (-0.7 * w.U0).astype(np.float32)]
        for it in range(3):
            U, stats = e.solve(w.x0, U_ins[it])
            outs.append((U.copy(), stats))
        # every rank must hold the same result
        t = torch.from_numpy(outs[-1][0].copy()).cuda()
        ref = t.clone()
        dist.broadcast(ref, src=0)
        ok &= bool(torch.equal(t, ref))
        if rank == 0:
            s = H.Engine(w.dyn, w.cost, w.sampler, w.N, w.T, w.D, device=local)
            s.set_solver(w.dt, w.lambda_, w.alpha)
            s.seed(w.seed, 0)
            for it in range(3):
                Us, sstats = s.solve(w.x0, U_ins[it])
                scale = max(1.0, float(np.abs(Us).max()))
                good = np.allclose(outs[it][0], Us, atol=2e-5 * scale, rtol=1e-4)
                for d in range(w.D):
                    good &= outs[it][1][d][0] == sstats[d][0]
                    good &= abs(outs[it][1][d][1] - sstats[d][1]) <= 1e-5 * abs(sstats[d][1])
                if not good:
                    print(f"MISMATCH {name} iter {it}: max |dU| {np.abs(outs[it][0] - Us).max()} stats {outs[it][1]} vs {sstats}")
                ok &= bool(good)
            s.close()
        e.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MGPU_OK" if int(flag) == 1 else "MGPU_FAIL", f"world={world}")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag) == 1 else 1)


if __name__ == "__main__":
    main()
