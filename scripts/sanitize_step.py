"""Small run of every step path for compute-sanitizer (development aid): fused DiagonalEK1, one-launch KroneckerEK0 (aligned and
unaligned grids), three-launch KroneckerEK0 on a chain, the device-resident controller, the host pipeline."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from tornadox_b200 import ek0, ek1, ivp, step

for prob in (ivp.fhn_2d(dx=1 / 64, y0=np.random.default_rng(2).uniform(size=2 * 64 * 64), tmax=2e-4),
             ivp.fhn_2d(dx=1 / 7, y0=np.random.default_rng(2).uniform(size=2 * 7 * 7), tmax=2e-3),
             ivp.lorenz96(num_variables=700, tmax=0.05), ivp.burgers_1d(dx=1 / 300, tmax=1e-3)):
    for cls in (ek1.DiagonalEK1, ek0.KroneckerEK0):
        for dl in (False, True):
            sol = cls(num_derivatives=4, steprule=step.AdaptiveSteps())
            sol.device_loop = dl
            sol.device_loop_batch = 3
            st, info = sol.simulate_final_state(prob)
            assert bool(st.y.mean.isfinite().all()), (prob.spec.name, cls.__name__, dl)
sol = ek1.DiagonalEK1(num_derivatives=3, steprule=step.AdaptiveSteps())
p = ivp.fhn_2d(dx=1 / 32, y0=np.random.default_rng(2).uniform(size=2 * 32 * 32))
st = sol.initialize(*p)
from tornadox_b200 import odefilter, rv
host = odefilter.ODEFilterState(st.t, rv.BatchedMultivariateNormal(st.y.mean.cpu().pin_memory(), st.y.cov_sqrtm.cpu().pin_memory()), None, None)
sol.engine.dek1_attempt_host(0.0, 1e-4, host.y.mean, host.y.cov_sqrtm, 1e-4, 1e-2, nchunks=3)
torch.cuda.synchronize()
print("SANITIZE_RUN_OK")
