"""TEST INFRASTRUCTURE: tests/golden/reference_vectors_ek0_const.npz -- ConstantSteps solves of the reference's KroneckerEK0
(the unmodified sources of /root/reference/tornadox on the NumPy-backed jax shim, like oracle/make_golden.py; build container only).

    python oracle/make_golden_ek0_const.py

Five steps of ``ek0.KroneckerEK0(num_derivatives=3, steprule=ConstantSteps(dt))`` from the Taylor-mode initial state for the six
problems of reference_vectors.npz (same problem instances, same dt = ``dense/<name>/dt``): final mean, final factor, final time,
step counters (tests/test_ek0.py:70-82 of the reference checks exactly these counters), and every accepted time."""

import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle.make_golden import problems  # noqa: E402
from oracle.run_reference import import_reference  # noqa: E402

OUT = ROOT / "tests" / "golden" / "reference_vectors_ek0_const.npz"


def main():
    tornadox = import_reference()
    import jax.numpy as jnp

    base = np.load(ROOT / "tests" / "golden" / "reference_vectors.npz")
    rng = np.random.default_rng(13)
    g = {}
    for name, (ivp, _) in problems(tornadox).items():
        dt = float(base[f"dense/{name}/dt"])
        ivp5 = ivp._replace(tmax=ivp.t0 + 5.0 * dt - 1e-12 * dt)
        sol = tornadox.ek0.KroneckerEK0(num_derivatives=3, steprule=tornadox.step.ConstantSteps(dt))
        ts, last, info = [], None, None
        for state, info in sol.solution_generator(ivp5):
            ts.append(float(state.t))
            last = state
        key = f"solve/{name}/ek0_const"
        g[key + "/dt"] = np.asarray(dt)
        g[key + "/ts"] = np.asarray(ts)
        g[key + "/final_mean"] = np.asarray(last.y.mean)
        g[key + "/final_cov_sqrtm"] = np.asarray(last.y.cov_sqrtm_2)
        g[key + "/info"] = np.asarray([info[k] for k in ("num_f_evaluations", "num_df_evaluations", "num_df_diagonal_evaluations",
                                                          "num_steps", "num_attempted_steps")])
        # conditioning probe: how far do the reference's OWN final mean and covariance move when y0 changes by ~1 ulp?  (after an
        # exact Taylor init the first calibration is rounding noise; brusselator: 6e-8 / 9e-4)  largest of four draws
        m_ref, S_ref = np.asarray(last.y.mean), np.asarray(last.y.cov_sqrtm_2)
        wm = ws = 0.0
        for _ in range(4):
            y = np.asarray(ivp5.y0)
            yw = jnp.array(y * (1.0 + 2.0 ** -52 * rng.uniform(-1, 1, size=y.shape)))
            st2, _ = tornadox.ek0.KroneckerEK0(num_derivatives=3, steprule=tornadox.step.ConstantSteps(dt)).simulate_final_state(ivp5._replace(y0=yw))
            m2, S2 = np.asarray(st2.y.mean), np.asarray(st2.y.cov_sqrtm_2)
            wm = max(wm, float(np.max(np.abs(m2 - m_ref)) / np.max(np.abs(m_ref))))
            ws = max(ws, float(np.max(np.abs(S2 @ S2.T - S_ref @ S_ref.T)) / np.max(np.abs(S_ref @ S_ref.T))))
        g[key + "/mean_noise"], g[key + "/cov_noise"] = np.asarray(wm), np.asarray(ws)
    OUT.parent.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT, **g)
    print(f"wrote {OUT} with {len(g)} arrays, {OUT.stat().st_size / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
