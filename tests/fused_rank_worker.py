"""One rank of the multi-process check of the fused ensemble run (tests/test_gpu_parity.py::test_run_fused_two_ranks).
Every rank runs on the GPU `--device` (two ranks may share one GPU: CUDA IPC works within a device and the time-sliced
persistent kernels still make progress), the process group is gloo (the handle exchange needs nothing else)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rank", type=int, required=True)
    ap.add_argument("--world", type=int, required=True)
    ap.add_argument("--port", type=int, required=True)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--out", required=True)
    ap.add_argument("--P", type=int, default=8192)
    ap.add_argument("--S", type=int, default=24)
    ap.add_argument("--lag", type=int, default=1)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist

    import physicsbasedbayesianinference_b200 as E
    from physicsbasedbayesianinference_b200.parallel import shard_range

    torch.cuda.set_device(a.device)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{a.port}", rank=a.rank, world_size=a.world)
    KB = 1.380649e-23
    D, L = 10, 6
    lo, hi = shard_range(a.P, a.rank, a.world)
    rng = np.random.RandomState(5)
    q0 = rng.standard_normal((D, a.P)).astype(np.float32)
    ens = E.Ensemble(D, hi - lo, dtype=np.float32, device=f"cuda:{a.device}", seed=9, particleOffset=lo)
    ens.q.copy_(torch.tensor(q0[:, lo:hi]))
    hmc = E.HMC(ens, L * 0.05 + 1e-9, 0.05, None, potential=E.FunnelPotential(D, 3.0), seed=9, bugCompat=False)
    r = hmc.run(a.S, 1 / KB, adapt=True, targetAccept=0.8, adaptIterations=a.S - 4, keepNumSteps=True, fused=True,
                group=dist.group.WORLD if a.world > 1 else None, traceParticles=8, adaptLag=a.lag)
    torch.cuda.synchronize()
    np.savez(a.out, q=ens.q.cpu().numpy(), stepSize=np.array(r["stepSize"]), acceptRate=np.array(r["acceptRate"]),
             meanH=np.array(r["meanH"]), mean=r["mean"].numpy(), var=r["var"].numpy(), lo=lo, hi=hi,
             final=hmc.stepSize, trace=r["trace"].cpu().numpy())
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
