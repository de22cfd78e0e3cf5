"""Per-iteration cost of the fused adaptive ensemble run (k_small_ens) at the 8-GPU shard size of config 5
(2^19 particles per GPU) on ONE GPU, i.e. without any communication, against the bare trajectory kernel."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import physicsbasedbayesianinference_b200 as E

KB = 1.380649e-23
D, L, h = 10, int(sys.argv[2]) if len(sys.argv) > 2 else 20, 0.05
import os

ctx = E._lib.Context.get()
if os.environ.get("EHMC_ENS_DEBUG"):
    ctx.set_option("ens_debug", float(os.environ["EHMC_ENS_DEBUG"]))
if os.environ.get("EHMC_ENS_SSHIFT"):
    ctx.set_option("ens_sshift", float(os.environ["EHMC_ENS_SSHIFT"]))  # sub-batches per queue item = 2^value
if os.environ.get("EHMC_ENS_GROUPS"):
    ctx.set_option("ens_groups", float(os.environ["EHMC_ENS_GROUPS"]))  # groups of batches per iteration at most
LAG = int(os.environ.get("EHMC_ADAPT_LAG", "2"))
for logP in (19, 20, 22):
    P = 1 << logP
    ens = E.Ensemble(D, P, dtype=np.float32, device="cuda", seed=1)
    ens.setPosition(1.0)
    hmc = E.HMC(ens, L * h + 1e-9, h, None, potential=E.FunnelPotential(D, 3.0), seed=1, bugCompat=False)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
    hmc.run(50, 1 / KB, adapt=True, keepNumSteps=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = hmc.run(n, 1 / KB, adapt=False, keepNumSteps=True, adaptLag=LAG)
    e1.record()
    torch.cuda.synchronize()
    fused = e0.elapsed_time(e1) / n
    if os.environ.get("EHMC_ENS_DEBUG"):
        ctx.set_option("ens_debug_dump", float(os.environ["EHMC_ENS_DEBUG"]) - 12)
    # bare kernel, no statistics, same step size
    for _ in range(20):
        hmc.step(1 / KB)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(300):
        hmc.step(1 / KB)
    e1.record()
    torch.cuda.synchronize()
    bare = e0.elapsed_time(e1) / 300
    st = torch.zeros(2 * D + 3, dtype=torch.float64, device="cuda")
    e0.record()
    for _ in range(300):
        hmc.step(1 / KB, stats=st)
    e1.record()
    torch.cuda.synchronize()
    with_stats = e0.elapsed_time(e1) / 300
    print(f"P=2^{logP} L={L} h={hmc.stepSize:.4f}: fused run {1e3 * fused:.2f} us/iteration | per-launch kernel without "
          f"statistics {1e3 * bare:.2f} us, with statistics (2 launches) {1e3 * with_stats:.2f} us", flush=True)
