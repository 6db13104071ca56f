"""TEST INFRASTRUCTURE: generate tests/golden/reference_vectors.npz by running the UNMODIFIED reference
(/root/reference/tornadox) on the NumPy-backed jax shim (oracle/jaxshim).

    python oracle/make_golden.py          # only works in the build container, where /root/reference exists
    python oracle/make_golden.py --small  # reference_vectors_small.npz: vanderpol_julia, threebody, pleiades, lorenz96_loop

The fixtures pin (i) the NumPy restatement in oracle/ref_np and (ii) the CUDA kernels against outputs of the
reference's own code: prior constants, vector fields and Jacobian diagonals, Taylor-mode initial states,
sequences of ``attempt_step`` calls of ek1.DiagonalEK1 / ek0.KroneckerEK0 / ek0.DiagonalEK0 and complete
adaptive solves (accepted-step sequences).  Inputs are fully explicit (no jax.random).
"""

import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle.run_reference import import_reference  # noqa: E402

OUT = ROOT / "tests" / "golden" / "reference_vectors.npz"
OUT_SMALL = ROOT / "tests" / "golden" / "reference_vectors_small.npz"      # python oracle/make_golden.py --small


def problems(tornadox):
    """name -> (reference ivp, dict of constructor facts the tests need to rebuild it elsewhere)"""
    rng = np.random.default_rng(2)
    out = {}
    out["vanderpol"] = (tornadox.ivp.vanderpol(t0=0.0, tmax=1.0, stiffness_constant=1.0), dict(mu=1.0))
    out["lorenz96"] = (tornadox.ivp.lorenz96(t0=0.0, tmax=0.5, num_variables=10, forcing=8.0), dict(N=10, forcing=8.0))
    out["brusselator"] = (tornadox.ivp.brusselator(N=7, t0=0.0, tmax=0.05), dict(N=7))
    y0 = rng.uniform(size=2 * 5 * 5)
    import jax.numpy as jnp

    out["fhn_2d"] = (tornadox.ivp.fhn_2d(t0=0.0, tmax=0.05, y0=jnp.array(y0), dx=0.2), dict(dx=0.2, nx=5))
    out["burgers_1d"] = (tornadox.ivp.burgers_1d(t0=0.0, tmax=0.05, dx=0.1), dict(dx=0.1))
    out["wave_1d"] = (tornadox.ivp.wave_1d(t0=0.0, tmax=0.05, dx=0.1), dict(dx=0.1))
    return out


def small_problems(tornadox):
    """The reference's remaining (small, dense) problems; kept in a file of their own (``--small``)."""
    out = {}
    out["vanderpol_julia"] = (tornadox.ivp.vanderpol_julia(t0=0.0, tmax=0.5, stiffness_constant=2.0), dict(mu=2.0))
    out["threebody"] = (tornadox.ivp.threebody(tmax=0.02), dict())
    out["pleiades"] = (tornadox.ivp.pleiades(t0=0.0, tmax=0.2), dict())
    out["lorenz96_loop"] = (tornadox.ivp.lorenz96_loop(t0=0.0, tmax=0.3, num_variables=6, forcing=8.0), dict(N=6, forcing=8.0))
    return out


def _ulp_perturbed(tornadox, jnp, state, sname, rng):
    """The same filter state with every entry of mean and factor multiplied by (1 + u * 2^-52), u ~ U(-1, 1)."""
    eps = 2.0 ** -52
    wiggle = lambda x: jnp.array(np.asarray(x) * (1.0 + eps * rng.uniform(-1, 1, size=np.shape(x))))
    y = state.y
    if sname == "ek0":
        y2 = tornadox.rv.LeftIsotropicMatrixNormal(wiggle(y.mean), y.d, wiggle(y.cov_sqrtm_2))
    else:
        y2 = tornadox.rv.BatchedMultivariateNormal(wiggle(y.mean), wiggle(y.cov_sqrtm))
    return tornadox.odefilter.ODEFilterState(state.t, y2, state.error_estimate, state.reference_state)


def _outputs(a, b, sname):
    fac = (lambda s: s.y.cov_sqrtm_2) if sname == "ek0" else (lambda s: s.y.cov_sqrtm)
    return [("mean", a.y.mean, b.y.mean), ("cov_sqrtm", fac(a), fac(b)),
            ("error_estimate", a.error_estimate, b.error_estimate),
            ("reference_state", a.reference_state, b.reference_state)]


def main(small=False):
    tornadox = import_reference()
    import jax.numpy as jnp

    g = {}
    rng = np.random.default_rng(12 if small else 11)

    # ---- prior constants (iwp.py:19-37, 65-73) ---------------------------------------------------
    for nu in (() if small else (1, 2, 3, 4, 5)):
        iwp = tornadox.iwp.IntegratedWienerTransition(wiener_process_dimension=1, num_derivatives=nu)
        A, Ql = iwp.preconditioned_discretize_1d
        g[f"prior/nu{nu}/A"], g[f"prior/nu{nu}/Ql"] = np.asarray(A), np.asarray(Ql)
        for dt in (0.1, 1e-3, 2.5):
            p, pinv = iwp.nordsieck_preconditioner_1d_raw(dt)
            g[f"prior/nu{nu}/p/{dt}"], g[f"prior/nu{nu}/pinv/{dt}"] = np.asarray(p), np.asarray(pinv)

    probs = small_problems(tornadox) if small else problems(tornadox)
    for name, (ivp, facts) in probs.items():
        y0 = np.asarray(ivp.y0, dtype=float)
        g[f"ivp/{name}/y0"] = y0
        # vector field + Jacobian diagonal at y0 and at a perturbed point (ivp.py)
        y1 = y0 + 0.1 * rng.standard_normal(y0.shape)
        g[f"ivp/{name}/y1"] = y1
        for tag, y in (("y0", y0), ("y1", y1)):
            g[f"ivp/{name}/f/{tag}"] = np.asarray(ivp.f(0.0, jnp.array(y)))
            g[f"ivp/{name}/df_diagonal/{tag}"] = np.asarray(ivp.df_diagonal(0.0, jnp.array(y)))

        for nu in (2, 3, 4):
            # Taylor-mode initial state (init.py:22-103)
            m0 = np.asarray(tornadox.init.TaylorMode.taylor_mode(ivp.f, ivp.y0, ivp.t0, nu))
            g[f"init/{name}/nu{nu}/taylor"] = m0
            # Runge-Kutta initialisation (init.py:109-315, use_df=False)
            rk = tornadox.init.RungeKutta(dt=0.01 if name not in ("fhn_2d", "brusselator") else 1e-3, method="RK45", use_df=False)
            m_rk, sc_rk = rk(f=ivp.f, df=ivp.df, y0=ivp.y0, t0=ivp.t0, num_derivatives=nu)
            g[f"init/{name}/nu{nu}/rk_mean"], g[f"init/{name}/nu{nu}/rk_cov_sqrtm"] = np.asarray(m_rk), np.asarray(sc_rk)
            # conditioning probe: the reference's own filter + smoother on RK points perturbed by ~1 ulp (fitting nu
            # derivatives to points dt apart amplifies rounding by ~dt^-nu; see tests/helpers.py: assert_close(noise=...))
            ts_rk, ys_rk = rk.rk_data(f=ivp.f, t0=ivp.t0, dt=rk.dt, num_steps=nu + 1, y0=ivp.y0, method="RK45")
            m_st, sc_st = rk.stack_initvals(f=ivp.f, df=ivp.df, y0=ivp.y0, t0=ivp.t0, num_derivatives=nu)
            ys_w = jnp.array(np.asarray(ys_rk) * (1.0 + 2.0 ** -52 * rng.uniform(-1, 1, size=np.shape(ys_rk))))
            m_w, sc_w = tornadox.init.RungeKutta.rk_init_improve(m=m_st, sc=sc_st, t0=ivp.t0, ts=ts_rk, ys=ys_w)
            g[f"init/{name}/nu{nu}/rk_noise_mean"] = np.max(np.abs(np.asarray(m_w) - np.asarray(m_rk)), axis=1)
            g[f"init/{name}/nu{nu}/rk_noise_cov_sqrtm"] = np.asarray(np.max(np.abs(np.asarray(sc_w) - np.asarray(sc_rk))))
            dts = {"vanderpol": [0.0123, 0.05, 0.02], "lorenz96": [0.01, 0.02, 0.005],
                   "brusselator": [0.02, 0.05, 0.01], "fhn_2d": [0.01, 0.03, 0.01],
                   "burgers_1d": [0.01, 0.02, 0.005], "wave_1d": [0.01, 0.03, 0.01],
                   "vanderpol_julia": [0.0123, 0.05, 0.02], "threebody": [1e-3, 2e-3, 5e-4],
                   "pleiades": [0.01, 0.02, 0.005], "lorenz96_loop": [0.01, 0.02, 0.005]}[name]
            g[f"steps/{name}/dts"] = np.asarray(dts)
            solvers = {
                "dek1": tornadox.ek1.DiagonalEK1,
                "ek0": tornadox.ek0.KroneckerEK0,
                "dek0": tornadox.ek0.DiagonalEK0,
            }
            for sname, cls in solvers.items():
                sol = cls(num_derivatives=nu, steprule=tornadox.step.AdaptiveSteps())
                state = sol.initialize(*ivp)
                for i, dt in enumerate(dts):
                    key = f"steps/{name}/nu{nu}/{sname}/{i}"
                    # conditioning probe: the same attempt from inputs perturbed by ~1 ulp, four draws, the largest
                    # change per output (see tests/helpers.py)
                    twins = [sol.attempt_step(_ulp_perturbed(tornadox, jnp, state, sname, rng), dt, *ivp)[0] for _ in range(4)]
                    state, info = sol.attempt_step(state, dt, *ivp)
                    for twin in twins:
                        for field, a, b in _outputs(state, twin, sname):
                            change = np.asarray(np.max(np.abs(np.asarray(a) - np.asarray(b))))
                            prev = g.get(key + "/noise/" + field)
                            g[key + "/noise/" + field] = change if prev is None else np.maximum(prev, change)
                    g[key + "/t"] = np.asarray(state.t)
                    g[key + "/mean"] = np.asarray(state.y.mean)
                    g[key + "/cov_sqrtm"] = np.asarray(state.y.cov_sqrtm if sname != "ek0" else state.y.cov_sqrtm_2)
                    g[key + "/error_estimate"] = np.asarray(state.error_estimate)
                    g[key + "/reference_state"] = np.asarray(state.reference_state)
                    norm = sol.steprule.scale_error_estimate(dt * state.error_estimate, state.reference_state)
                    g[key + "/scaled_error_norm"] = np.asarray(norm)

        # attempt_step from a DENSE random mid-solve state (not reachable from TaylorMode's zero factor)
        nu = 4
        n, d = nu + 1, y0.shape[0]
        mean = rng.standard_normal((n, d))
        mean[0] = y0
        blocks = 1e-2 * rng.standard_normal((d, n, n))
        single = 1e-2 * rng.standard_normal((n, n))
        dt = {"vanderpol": 0.0123, "lorenz96": 0.01, "brusselator": 1e-4, "fhn_2d": 1e-3, "burgers_1d": 1e-3,
              "wave_1d": 1e-3, "vanderpol_julia": 0.0123, "threebody": 1e-3, "pleiades": 0.01, "lorenz96_loop": 0.01}[name]
        g[f"dense/{name}/mean"], g[f"dense/{name}/blocks"], g[f"dense/{name}/single"] = mean, blocks, single
        g[f"dense/{name}/dt"] = np.asarray(dt)
        sol = tornadox.ek1.DiagonalEK1(num_derivatives=nu)
        sol.initialize(*ivp)
        st = tornadox.odefilter.ODEFilterState(
            t=0.3, y=tornadox.rv.BatchedMultivariateNormal(jnp.array(mean), jnp.array(blocks)),
            error_estimate=None, reference_state=None)
        new, _ = sol.attempt_step(st, dt, *ivp)
        g[f"dense/{name}/dek1/mean"] = np.asarray(new.y.mean)
        g[f"dense/{name}/dek1/cov_sqrtm"] = np.asarray(new.y.cov_sqrtm)
        g[f"dense/{name}/dek1/error_estimate"] = np.asarray(new.error_estimate)
        g[f"dense/{name}/dek1/reference_state"] = np.asarray(new.reference_state)
        sol0 = tornadox.ek0.KroneckerEK0(num_derivatives=nu)
        sol0.initialize(*ivp)
        st0 = tornadox.odefilter.ODEFilterState(
            t=0.3, y=tornadox.rv.LeftIsotropicMatrixNormal(jnp.array(mean), d, jnp.array(single)),
            error_estimate=None, reference_state=None)
        new0, _ = sol0.attempt_step(st0, dt, *ivp)
        g[f"dense/{name}/ek0/mean"] = np.asarray(new0.y.mean)
        g[f"dense/{name}/ek0/cov_sqrtm"] = np.asarray(new0.y.cov_sqrtm_2)
        g[f"dense/{name}/ek0/error_estimate"] = np.asarray(new0.error_estimate)
        g[f"dense/{name}/ek0/reference_state"] = np.asarray(new0.reference_state)

        # complete adaptive solves: accepted-step sequence, counters, final state (odefilter.py:80-223)
        for sname, cls in (("dek1", tornadox.ek1.DiagonalEK1), ("ek0", tornadox.ek0.KroneckerEK0)):
            sol = cls(num_derivatives=4, steprule=tornadox.step.AdaptiveSteps(abstol=1e-4, reltol=1e-2))
            ts, last, info = [], None, None
            for state, info in sol.solution_generator(ivp):
                ts.append(float(state.t))
                last = state
            key = f"solve/{name}/{sname}"
            g[key + "/ts"] = np.asarray(ts)
            g[key + "/final_mean"] = np.asarray(last.y.mean)
            g[key + "/final_cov_sqrtm"] = np.asarray(last.y.cov_sqrtm if sname == "dek1" else last.y.cov_sqrtm_2)
            g[key + "/info"] = np.asarray([info[k] for k in ("num_f_evaluations", "num_df_evaluations",
                                                              "num_df_diagonal_evaluations", "num_steps",
                                                              "num_attempted_steps")])
            if small:
                # conditioning probe: how far do the reference's own accepted times move when y0 changes by ~1 ulp?  (the
                # controller feeds every rounding difference of the error norm into all later steps; threebody starts next
                # to a singularity)  inf = the perturbed solve took a different number of steps
                worst = 0.0
                for _ in range(4):
                    yw = jnp.array(y0 * (1.0 + 2.0 ** -52 * rng.uniform(-1, 1, size=y0.shape)))
                    tw = [float(st.t) for st, _ in sol.solution_generator(ivp._replace(y0=yw))]
                    same = len(tw) == len(ts)
                    worst = max(worst, float(np.max(np.abs(np.asarray(tw)[1:] / np.asarray(ts)[1:] - 1.0))) if same else np.inf)
                g[key + "/ts_noise"] = np.asarray(worst)
        # ConstantSteps counters (tests/test_ek0.py:70-82, tests/test_ek1.py:65-78)
        sol = tornadox.ek1.DiagonalEK1(num_derivatives=3, steprule=tornadox.step.ConstantSteps(dt))
        ivp5 = ivp._replace(tmax=ivp.t0 + 5.0 * dt - 1e-12 * dt) if hasattr(ivp, "_replace") else ivp
        state, info = sol.simulate_final_state(ivp5)
        g[f"solve/{name}/dek1_const/final_mean"] = np.asarray(state.y.mean)
        g[f"solve/{name}/dek1_const/t"] = np.asarray(state.t)
        g[f"solve/{name}/dek1_const/num_steps"] = np.asarray(info["num_steps"])

    # step rule scalars (step.py:76-104)
    if not small:
        rule = tornadox.step.AdaptiveSteps(abstol=1e-3, reltol=1e-1)
        err, ref = rng.uniform(size=7), rng.uniform(size=7) + 0.5
        g["step/err"], g["step/ref"] = err, ref
        g["step/scaled"] = np.asarray(rule.scale_error_estimate(jnp.array(err), jnp.array(ref)))
        g["step/suggest"] = np.asarray([rule.suggest(0.1, e, local_convergence_rate=5) for e in (1e-8, 0.3, 1.0, 7.0, 1e9)])

    out = OUT_SMALL if small else OUT
    out.parent.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(out, **g)
    print(f"wrote {out} with {len(g)} arrays, {out.stat().st_size / 1024:.0f} KiB")


if __name__ == "__main__":
    main(small="--small" in sys.argv)
