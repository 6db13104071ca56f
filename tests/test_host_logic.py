"""CPU: host-side logic of the product (step rules, time stopper, problem factories, Taylor-mode series
algebra, state containers) and the reference's known-answer tests ported onto the oracle."""

import math

import numpy as np
import pytest
import torch

from helpers import assert_close, golden
from oracle.ref_np import driver as odriver
from oracle.ref_np import filters as ofilters
from oracle.ref_np import prior as oprior
from oracle.ref_np import problems as oproblems


# ---- reference known-answer tests, on the oracle -------------------------------------------------------
def test_ibm2_closed_form_at_dt_01():
    """tests/test_iwp.py:19-47 of the reference: closed-form IBM(2) transition and process noise at dt = 0.1."""
    dt, nu = 0.1, 2
    A, Ql = oprior.dense_transition_no_precon(nu, 1, dt)
    ah_22_ibm = np.array([[1.0, 0.1, 0.005], [0.0, 1.0, 0.1], [0.0, 0.0, 1.0]])
    np.testing.assert_allclose(A, ah_22_ibm, rtol=1e-12, atol=1e-15)
    Q = Ql @ Ql.T
    expected_Q = np.array([[dt**5 / 20, dt**4 / 8, dt**3 / 6], [dt**4 / 8, dt**3 / 3, dt**2 / 2],
                           [dt**3 / 6, dt**2 / 2, dt]])
    np.testing.assert_allclose(Q, expected_Q, rtol=1e-10)


def test_preconditioner_identities():
    """tests/test_iwp.py:50-93: P A_pre P^-1 equals the plain transition and P P^-1 = I."""
    for nu in (1, 2, 4):
        for dt in (0.1, 0.7):
            P, PI = oprior.nordsieck_diag(nu, dt)
            np.testing.assert_allclose(P @ PI, np.eye(nu + 1), rtol=1e-13, atol=1e-13)
            A = P @ oprior.transition_1d(nu) @ PI
            expected = np.array([[dt ** (j - i) / math.factorial(j - i) if j >= i else 0.0 for j in range(nu + 1)]
                                 for i in range(nu + 1)])
            np.testing.assert_allclose(A, expected, rtol=1e-12, atol=1e-14)


def test_propagate_cholesky_factor_identity():
    """tests/test_sqrt.py:39-48: L L^T = S1 S1^T + S2 S2^T and L is lower triangular (also batched)."""
    rng = np.random.default_rng(0)
    S1, S2 = rng.standard_normal((4, 4)), rng.standard_normal((4, 4))
    L = oprior.propagate_cholesky(S1, S2)
    np.testing.assert_allclose(L @ L.T, S1 @ S1.T + S2 @ S2.T, rtol=1e-12)
    np.testing.assert_allclose(L, np.tril(L))
    B1, B2 = rng.standard_normal((3, 4, 4)), rng.standard_normal((3, 4, 4))
    Lb = oprior.propagate_cholesky_batched(B1, B2)
    for i in range(3):
        np.testing.assert_allclose(Lb[i], oprior.propagate_cholesky(B1[i], B2[i]), rtol=1e-13, atol=1e-13)


def test_kronecker_ek0_equals_dense_reference_ek0():
    """tests/test_ek0.py:246-263: one step dt = 0.12345 on vanderpol(mu=1), nu = 2: Kronecker == dense."""
    nu, dt = 2, 0.12345
    prob = oproblems.vanderpol(t0=0.0, tmax=10.0, stiffness_constant=1.0)
    m0, c0 = odriver.taylor_init(odriver.series_vanderpol(1.0), prob.y0, 0.0, nu)
    kron, _ = ofilters.kronecker_ek0_attempt(ofilters.kronecker_ek0_initial_state(m0, c0, 0.0), dt, prob.f, nu)
    dense, _ = ofilters.reference_ek0_attempt(ofilters.dense_initial_state(m0, c0, 0.0), dt, prob.f, nu)
    np.testing.assert_allclose(kron.mean, dense.mean, rtol=1e-10, atol=1e-12)
    d = 2
    kron_cov = np.kron(np.eye(d), kron.cov_sqrtm @ kron.cov_sqrtm.T)
    np.testing.assert_allclose(kron_cov, dense.cov_sqrtm @ dense.cov_sqrtm.T, rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(kron.error_estimate, dense.error_estimate, rtol=1e-10)
    np.testing.assert_allclose(kron.reference_state, dense.reference_state, rtol=1e-10)


def test_diagonal_ek1_reference_state_equals_dense_reference_ek1():
    """tests/test_ek1.py:239-247: with a diagonalised Jacobian, reference_state (and the first-step mean) agree."""
    nu = 3
    prob = oproblems.vanderpol(t0=0.0, tmax=10.0, stiffness_constant=1.0)
    m0, c0 = odriver.taylor_init(odriver.series_vanderpol(1.0), prob.y0, 0.0, nu)
    for dt in (0.12121, 12.345):
        diag, _ = ofilters.diagonal_ek1_attempt(ofilters.batched_initial_state(m0, c0, 0.0), dt, prob.f,
                                                prob.df_diagonal, nu)
        df_diag_only = lambda t, y: np.diag(prob.df_diagonal(t, y))
        dense, _ = ofilters.reference_ek1_attempt(ofilters.dense_initial_state(m0, c0, 0.0), dt, prob.f, df_diag_only, nu)
        np.testing.assert_allclose(diag.reference_state, dense.reference_state, rtol=1e-9)
        np.testing.assert_allclose(diag.mean, dense.mean, rtol=1e-8, atol=1e-10)


def test_taylor_mode_threebody_literals():
    """tests/test_init.py:45-148 of the reference (first derivatives of the restricted three-body problem)."""
    y0 = np.array([0.994, 0, 0, -2.00158510637908252240537862224])
    m = odriver.taylor_mode(odriver.series_threebody(), y0, 0.0, 4)
    expected_col0 = [0.994, 0.0, -315.5430234888826712, 0.0, 6.390281114012432978e07]
    expected_col1 = [0.0, -2.001585106379082379, 0.0, 99972.09449511380974, 0.0]
    np.testing.assert_allclose(m[:, 0], expected_col0, rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(m[:, 1], expected_col1, rtol=1e-12, atol=1e-9)


def test_df_diagonal_equals_dense_jacobian_diagonal():
    """tests/test_ivp.py:60-64, extended to an fhn grid with edge AND interior nodes (5 x 5)."""
    rng = np.random.default_rng(3)
    probs = [oproblems.vanderpol(), oproblems.brusselator(N=6), oproblems.lorenz96(num_variables=9),
             oproblems.fhn_2d(dx=0.2, y0=rng.uniform(size=50)), oproblems.burgers_1d(dx=0.1), oproblems.wave_1d(dx=0.125)]
    for prob in probs:
        y = prob.y0 + 0.1 * rng.standard_normal(prob.y0.shape)
        np.testing.assert_allclose(prob.df_diagonal(0.0, y), np.diag(prob.df(0.0, y)), rtol=1e-11, atol=1e-11)
        eps = 1e-6
        J = np.stack([(prob.f(0.0, y + eps * e) - prob.f(0.0, y - eps * e)) / (2 * eps) for e in np.eye(y.size)], axis=1)
        np.testing.assert_allclose(prob.df(0.0, y), J, rtol=1e-5, atol=1e-4)


# ---- product host logic ---------------------------------------------------------------------------------
def test_step_rules_match_reference_values():
    from tornadox_b200 import step

    g = golden()
    rule = step.AdaptiveSteps(abstol=1e-3, reltol=1e-1)
    scaled = rule.scale_error_estimate(torch.as_tensor(g["step/err"]), torch.as_tensor(g["step/ref"]))
    assert_close(scaled, g["step/scaled"], 1e-14, "scale_error_estimate")
    sug = [rule.suggest(0.1, e, local_convergence_rate=5) for e in (1e-8, 0.3, 1.0, 7.0, 1e9)]
    assert_close(sug, g["step/suggest"], 1e-15, "suggest")
    assert rule.is_accepted(0.99) and not rule.is_accepted(1.0) and not rule.is_accepted(1.01)   # step.py:90-91
    with pytest.raises(ValueError):
        rule.suggest(0.1, 0.5)
    with pytest.raises(ValueError):
        rule.scale_error_estimate(torch.zeros(3), torch.zeros(4))
    const = step.ConstantSteps(0.25)
    assert const.suggest(1.0, 3.0) == 0.25 and const.is_accepted(123.0) is True
    assert const.scale_error_estimate(torch.zeros(2), torch.zeros(2)) is None
    assert const.first_dt(None, 0, 1, None, None, None) == 0.25
    assert step.norm_from_sq_sum(8.0, 2) == 2.0


def test_time_stopper():
    from tornadox_b200.odefilter import _TimeStopper

    ts = _TimeStopper([0.5, 1.0])
    assert ts.adjust_dt_to_time_stops(0.0, 0.3) == 0.3
    assert ts.adjust_dt_to_time_stops(0.3, 0.3) == pytest.approx(0.2)
    assert ts.adjust_dt_to_time_stops(0.5, 0.7) == pytest.approx(0.5)
    assert ts.adjust_dt_to_time_stops(1.0, 0.7) == 0.7


def test_problem_factories_and_specs():
    from tornadox_b200 import _lib, ivp

    p = ivp.fhn_2d(dx=0.25, y0=np.zeros(32))
    assert p.spec.kind == _lib.TDX_IVP_FHN2D and (p.spec.grid_rows, p.spec.grid_cols) == (4, 4)
    assert p.dimension == 32 and p.t_span == (0.0, 20.0) and p.df is None
    with pytest.raises(ValueError, match="quadratic spatial grids"):
        ivp.fhn_2d(bbox=[[0.0, 0.0], [1.0, 2.0]], dx=0.1)
    b = ivp.brusselator(N=5)
    np.testing.assert_allclose(b.y0, oproblems.brusselator(N=5).y0)
    lz = ivp.lorenz96(num_variables=7)
    np.testing.assert_allclose(lz.y0, oproblems.lorenz96(num_variables=7).y0)
    v = ivp.vanderpol()
    assert v.spec.params == (10.0,) and v.tmax == 30
    assert len(tuple(v)) == 6   # *ivp unpacking as in odefilter.py:109


@pytest.mark.parametrize("name,size", [("vanderpol", 1), ("lorenz96", 9), ("brusselator", 6), ("fhn_2d", 5),
                                       ("burgers_1d", 11), ("wave_1d", 9), ("vanderpol_julia", 1), ("threebody", 4),
                                       ("pleiades", 28)])
def test_taylor_mode_series_algebra_on_cpu_tensors(name, size):
    """The product's Taylor-mode initialisation is plain tensor algebra: check it on CPU tensors vs the oracle."""
    from helpers import make_pair
    from tornadox_b200 import init

    o, p, series_f = make_pair(name, size)
    for nu in (1, 3, 4):
        m0 = init.TaylorMode.taylor_mode(p.f, torch.as_tensor(o.y0), 0.0, nu)
        ref = odriver.taylor_mode(series_f, o.y0, 0.0, nu)
        assert_close(m0.numpy(), ref, 1e-13, "taylor mode")


def test_user_defined_vector_field_maps_to_the_external_kernel_tag():
    """A callable without a kernel tag is not rejected: it becomes the EXTERNAL vector field (the caller's torch code evaluates
    f at the predicted state, the fused kernels do the rest), and Taylor mode differentiates it with nested forward mode."""
    from tornadox_b200 import _lib, init

    y0 = torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64)
    spec = init._spec_of(lambda t, y: y, y0)
    assert spec.kind == _lib.TDX_IVP_EXTERNAL and (spec.grid_rows, spec.grid_cols, spec.n_fields) == (1, 3, 1)
    # dy/dt = t y: second derivative y + t dy/dt, third derivative 2 dy/dt + t d2y/dt2
    m, c = init.TaylorMode()(lambda t, y: t * y, None, y0, 0.5, 3)
    expect = torch.stack([y0, 0.5 * y0, 1.25 * y0, 1.625 * y0])
    assert torch.allclose(m, expect, rtol=1e-14, atol=0) and float(c.abs().max()) == 0.0
    with pytest.raises(ValueError):
        init._spec_of(None, y0)
    # Stack(use_df=True) with the caller's dense Jacobian (init.py:377-385)
    m, c = init.Stack(use_df=True)(lambda t, y: -y, lambda t, y: -torch.eye(3, dtype=torch.float64), y0, 0.0, 3)
    assert torch.equal(m[2], y0) and torch.equal(torch.diagonal(c), torch.tensor([0.0, 0.0, 0.0, 1e3], dtype=torch.float64))


def test_rv_containers():
    from tornadox_b200 import rv

    S = torch.tril(torch.rand(3, 2, 2, dtype=torch.float64))
    b = rv.BatchedMultivariateNormal(torch.zeros(2, 3), S)
    assert torch.allclose(b.cov, S @ S.transpose(1, 2))
    li = rv.LeftIsotropicMatrixNormal(torch.zeros(2, 3), 3, S[0])
    assert li.dense_cov().shape == (6, 6)
    assert torch.allclose(li.dense_cov(), torch.kron(torch.eye(3, dtype=torch.float64), S[0] @ S[0].T))
    assert li.cov.shape == (3, 2, 2)


def test_deferred_reference_state_resolves_on_access():
    """KroneckerEK0 stores reference_state = |mean[0]| lazily (odefilter.DeferredAbsRow); attribute access materialises it."""
    from tornadox_b200 import odefilter

    mean = torch.tensor([[1.0, -2.0, 3.0], [0.5, 0.5, 0.5]], dtype=torch.float64)
    st = odefilter.ODEFilterState(t=0.1, y=None, error_estimate=0.0, reference_state=odefilter.DeferredAbsRow(mean))
    assert torch.equal(st.reference_state, torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64))
    assert st.reference_state is st.reference_state            # computed once
    row = mean[0]
    plain = odefilter.ODEFilterState(t=0.1, y=None, error_estimate=0.0, reference_state=row)
    assert plain.reference_state is row
    assert plain._replace(t=0.2).t == 0.2 and len(tuple(plain)) == 4


def test_native_steprule_mirror_and_shard_ranges():
    """engine.native_steprule only maps the reference's own rule classes; burgers_1d / wave_1d refuse to shard."""
    from tornadox_b200 import engine, ivp, step

    class Mine(step.AdaptiveSteps):
        pass

    fake = engine.StepEngine.__new__(engine.StepEngine)
    r = engine.StepEngine.native_steprule(fake, step.AdaptiveSteps(abstol=1e-3, reltol=1e-2, max_changes=(0.1, 5.0), safety_scale=0.9))
    assert (r.adaptive, r.abstol, r.reltol, r.min_change, r.max_change, r.safety_scale) == (1, 1e-3, 1e-2, 0.1, 5.0, 0.9)
    c = engine.StepEngine.native_steprule(fake, step.ConstantSteps(0.25))
    assert (c.adaptive, c.constant_dt) == (0, 0.25)
    assert engine.StepEngine.native_steprule(fake, Mine()) is None
    with pytest.raises(NotImplementedError):
        engine.shard_range(ivp.burgers_1d(dx=0.1).spec, 0, 2)
    for small in (ivp.threebody(), ivp.pleiades(), ivp.vanderpol_julia()):
        assert engine.shard_range(small.spec, 0, 1) == (0, small.spec.grid_cols)
        with pytest.raises(ValueError):
            engine.shard_range(small.spec, 0, 2)
    assert engine.shard_range(ivp.lorenz96(num_variables=10).spec, 1, 2) == (5, 10)


def test_scripts_compile():
    """scripts/ holds development aids that no test runs on the CPU: at least keep them syntactically valid."""
    import pathlib
    import py_compile

    root = pathlib.Path(__file__).resolve().parent.parent / "scripts"
    names = sorted(root.glob("*.py"))
    assert len(names) >= 8
    for path in names:
        py_compile.compile(str(path), doraise=True)


def test_shard_boundaries_sit_on_the_summation_grid_where_the_size_allows():
    """engine.shard_range / reduction_is_sharding_independent (step.py:93-104 is ONE norm over all coordinates; the kernels add
    one partial per work unit in 16 global chunks, csrc/tdx_device.cuh RedPlan): BASELINE's large grids cut into 2 / 4 / 8 equal
    shards are aligned, sizes that cannot be cut that way are reported as such, and every split still covers the axis."""
    from tornadox_b200 import engine, ivp

    big = ivp.fhn_2d(dx=1.0 / 2048, y0=np.zeros(2)).spec
    for world in (1, 2, 4, 8, 16):
        assert engine.reduction_is_sharding_independent(big, world)
        bounds = [engine.shard_range(big, r, world) for r in range(world)]
        assert bounds[0][0] == 0 and bounds[-1][1] == 2048
        assert all(b[1] == c[0] for b, c in zip(bounds, bounds[1:]))
        assert all(e - b == 2048 // world for b, e in bounds)
    chain = ivp.brusselator(N=4096).spec
    assert all(engine.reduction_is_sharding_independent(chain, w) for w in (2, 4, 8))
    assert not engine.reduction_is_sharding_independent(ivp.brusselator(N=400).spec, 8)
    assert not engine.reduction_is_sharding_independent(ivp.fhn_2d(dx=1.0 / 50, y0=np.zeros(2)).spec, 2)
    # small / odd sizes: balanced to within one row, still a partition
    odd = ivp.fhn_2d(dx=1.0 / 50, y0=np.zeros(2)).spec
    bounds = [engine.shard_range(odd, r, 3) for r in range(3)]
    assert bounds[0][0] == 0 and bounds[-1][1] == 50 and all(b[1] == c[0] for b, c in zip(bounds, bounds[1:]))
    # the fused fhn KroneckerEK0 kernel sums per 8-row block on large grids only (a function of the GLOBAL grid)
    assert engine._ek0_row_block(2048, 8) == 8 and engine._ek0_row_block(256, 1) == 1 and engine._ek0_row_block(1024, 4) == 8


def test_chunk_ordered_sum_is_independent_of_the_number_of_ranks():
    """A NumPy model of the kernels' summation order (reduce_units in csrc/tdx_device.cuh): unit partials -> 16 chunk sums (32
    lane-strided sequential sums, xor-shuffle tree) -> total in (rank, chunk) order.  With shard boundaries on chunk boundaries
    the total has the same bits for 1, 2, 4, 8 ranks; the plain left-to-right sum of the same numbers does not."""
    rng = np.random.default_rng(0)
    units = 4096
    part = rng.uniform(size=units) * 10.0 ** rng.integers(-6, 6, size=units)

    def chunk_sum(v):
        lanes = np.zeros(32)
        for lane in range(32):
            acc = 0.0
            for x in v[lane::32]:
                acc += x
            lanes[lane] = acc
        o = 16
        while o:
            lanes = lanes + lanes[np.arange(32) ^ o]
            o >>= 1
        return float(lanes[0])

    def total(world):
        out = 0.0
        for r in range(world):                       # every rank sends the sums of the chunks it owns ...
            lo, hi = r * units // world, (r + 1) * units // world
            for c in range(16):
                b0, b1 = c * units // 16, (c + 1) * units // 16
                b0, b1 = max(b0, lo), min(b1, hi)
                if b0 < b1:
                    out += chunk_sum(part[b0:b1])     # ... and everybody adds them in (rank, chunk) order
        return out

    ref = total(1)
    assert all(total(w) == ref for w in (2, 4, 8, 16))
    # 3 ranks do not sit on the chunk grid: deterministic, but not the single-GPU bits in general
    assert abs(total(3) - ref) <= 1e-12 * abs(ref)
