import sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
import physicsbasedbayesianinference_b200 as E, bench
KB = 1.380649e-23
def soak(name, n):
    cfg = bench.CONFIGS[name]
    D, P, L, h = cfg["D"], cfg["P"], cfg["L"], cfg["h"]
    pot = bench.make_potential(E, name, D)
    ens = E.Ensemble(D, P, dtype=np.float32, device="cuda", seed=3)
    ens.setPosition(1.0)
    hmc = E.HMC(ens, L * h + 1e-9, h, None, potential=pot, seed=3, bugCompat=False)
    t0 = time.perf_counter()
    r = hmc.run(n, 1 / KB, adapt=name.startswith("c5"), keepNumSteps=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ok = bool(torch.isfinite(ens.q).all().item())
    print(f"{name}: {n} iterations in {dt:.1f} s ({dt / n * 1e3:.3f} ms each), finite={ok}, acceptance first/last "
          f"{r['acceptRate'][0]:.3f}/{r['acceptRate'][-1]:.3f}, meanH last {r['meanH'][-1]:.3f}", flush=True)
    assert ok
soak("c2", 3000)
soak("c5", 8000)
soak("c1", 5000)
soak("c3", 60)
soak("c4", 40)
print("soak ok")
