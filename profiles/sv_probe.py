import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import physicsbasedbayesianinference_b200 as E, bench
KB = 1.380649e-23
D, P, L, h = 100, 1 << 20, 50, 0.05
pot = E.GaussianPotential(precision=bench.make_precision(D))
ctx = E._lib.Context.get()
for method in ("Leapfrog", "Stormer-Verlet"):
    for path in (0, 1):
        ctx.set_option("dense_path", path)
        ens = E.Ensemble(D, P, dtype=np.float32, device="cuda", seed=1)
        ens.setPosition(1.0)
        hmc = E.HMC(ens, L * h + 1e-9, h, None, potential=pot, seed=1, bugCompat=False, method=method)
        acc = torch.empty(P, dtype=torch.uint8, device="cuda")
        for _ in range(3):
            hmc.step(1 / KB, accept=acc)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            hmc.step(1 / KB, accept=acc)
        e1.record()
        torch.cuda.synchronize()
        print(f"{method:15s} dense_path={path}: {e0.elapsed_time(e1) / 10:.3f} ms / iteration, acceptance {acc.float().mean().item():.3f}")
ctx.set_option("dense_path", 0)
