import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import physicsbasedbayesianinference_b200 as E
KB = 1.380649e-23
D, h = 100, 0.05
rng = np.random.RandomState(1)
A = rng.standard_normal((D, D)); prec = A @ A.T / D + np.eye(D)
ctx = E._lib.Context.get()
pot = E.GaussianPotential(precision=prec).handle(32, ctx)
P = 148 * 128
q = torch.randn(D, P, device="cuda"); mass = torch.ones(P, device="cuda")
z = torch.randn(D, P, device="cuda"); u = torch.rand(P, device="cuda")
def t(L, dbg, fed, hmc=True):
    ctx.set_option("tc_debug", dbg)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(23):
        if i == 3: e0.record()
        args = E._lib.make_args(h, h * h, L, KB, 1 / KB, seed=1, iteration=i)
        E._lib.hmc_iter(ctx, pot, q, mass, args, z=z if fed else None, u=u if fed else None)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 20 * 1e3
for fed in (False, True):
    for dbg in (3, 0):
        a, b = t(0, dbg, fed), t(200, dbg, fed)
        print(f"fed={fed} dbg={dbg}: per evaluation {(b - a) / 200 * 1e3:.0f} ns")
ctx.set_option("tc_debug", 0)
