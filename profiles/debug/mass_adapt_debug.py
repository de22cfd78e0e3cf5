"""Debug aid for test_run_adapt_mass[diag10]: step sizes, acceptance and per-dimension ESS with / without mass adaptation."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import physicsbasedbayesianinference_b200 as E
from physicsbasedbayesianinference_b200 import diagnostics

KB = 1.380649e-23
rng = np.random.RandomState(41)
P, warm, S = 8192, 200, 300
D = 10
sd = np.logspace(-1.5, 1.0, D)
for L in (16, 5, 3):
    for am in (False, True):
        pot, h0 = E.HarmonicPotential(1.0 / sd**2), 0.01
        ens = E.Ensemble(D, P, dtype=np.float32, device="cuda", seed=3)
        ens.q.copy_(torch.tensor(rng.standard_normal((D, P)) * sd[:, None] * 0.5, dtype=torch.float32))
        hmc = E.HMC(ens, L * h0 + 1e-9, h0, None, potential=pot, seed=3, bugCompat=False)
        rw = hmc.run(warm, 1 / KB, adapt=True, adaptMass=am, keepNumSteps=True)
        r = hmc.run(S, 1 / KB, traceParticles=128)
        torch.cuda.synchronize()
        tr = r["trace"].double().cpu()
        per = [diagnostics.ess(tr[d].transpose(0, 1)) for d in range(D)]
        ess = min(per)
        print(f"L={L} adaptMass={am}: h={hmc.stepSize:.4f} L*h={L * hmc.stepSize:.3f} acc={np.mean(r['acceptRate']):.3f} "
              f"ess/S={ess / S:.3f}")
        print("   scale", None if hmc.massScale is None else np.round(hmc.massScale / sd, 3))
        print("   per-dim ess/S", np.round(np.asarray(per) / S, 2))
        print("   std/sd", np.round(ens.q.double().std(dim=1).cpu().numpy() / sd, 3))
