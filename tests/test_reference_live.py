"""CPU, build container only: run the UNMODIFIED reference (on the NumPy-backed jax shim) next to the NumPy
restatement on fresh random inputs -- beyond the committed golden vectors.  Skipped where /root/reference is absent."""

import numpy as np
import pytest

from oracle import run_reference
from oracle.ref_np import filters as ofilters
from oracle.ref_np import problems as oproblems

pytestmark = pytest.mark.skipif(not run_reference.available(), reason="/root/reference only exists in the build container")


@pytest.fixture(scope="module")
def tornadox():
    return run_reference.import_reference()


@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("nu", [1, 2, 5])
def test_random_states_fhn(tornadox, seed, nu):
    import jax.numpy as jnp

    rng = np.random.default_rng(seed)
    side = 4
    y0 = rng.uniform(size=2 * side * side)
    ref_ivp = tornadox.ivp.fhn_2d(y0=jnp.array(y0), dx=1.0 / side)
    prob = oproblems.fhn_2d(y0=y0, dx=1.0 / side)
    n, d = nu + 1, y0.size
    mean = rng.standard_normal((n, d))
    blocks = 1e-1 * rng.standard_normal((d, n, n))
    single = 1e-1 * rng.standard_normal((n, n))
    dt = 10.0 ** rng.uniform(-5, -2)

    sol = tornadox.ek1.DiagonalEK1(num_derivatives=nu)
    sol.initialize(*ref_ivp)
    st = tornadox.odefilter.ODEFilterState(0.1, tornadox.rv.BatchedMultivariateNormal(jnp.array(mean), jnp.array(blocks)), None, None)
    ref, _ = sol.attempt_step(st, dt, *ref_ivp)
    mine, _ = ofilters.diagonal_ek1_attempt(ofilters.FilterState(0.1, mean, blocks, None, None), dt, prob.f, prob.df_diagonal, nu)
    np.testing.assert_allclose(mine.mean, np.asarray(ref.y.mean), rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(mine.cov_sqrtm, np.asarray(ref.y.cov_sqrtm), rtol=1e-11, atol=1e-14)
    np.testing.assert_allclose(mine.error_estimate, np.asarray(ref.error_estimate), rtol=1e-11, atol=1e-14)

    sol0 = tornadox.ek0.KroneckerEK0(num_derivatives=nu)
    sol0.initialize(*ref_ivp)
    st0 = tornadox.odefilter.ODEFilterState(0.1, tornadox.rv.LeftIsotropicMatrixNormal(jnp.array(mean), d, jnp.array(single)), None, None)
    ref0, _ = sol0.attempt_step(st0, dt, *ref_ivp)
    mine0, _ = ofilters.kronecker_ek0_attempt(ofilters.FilterState(0.1, mean, single, None, None), dt, prob.f, nu)
    np.testing.assert_allclose(mine0.mean, np.asarray(ref0.y.mean), rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(mine0.cov_sqrtm, np.asarray(ref0.y.cov_sqrtm_2), rtol=1e-11, atol=1e-14)


def test_golden_file_is_reproducible(tornadox, tmp_path, monkeypatch):
    """oracle/make_golden.py regenerates exactly the committed fixtures."""
    import oracle.make_golden as mg

    monkeypatch.setattr(mg, "OUT", tmp_path / "regen.npz")
    monkeypatch.setattr(mg, "OUT_SMALL", tmp_path / "regen_small.npz")
    mg.main()
    mg.main(small=True)
    import oracle.make_golden_ek0_const as mgc

    monkeypatch.setattr(mgc, "OUT", tmp_path / "regen_ek0_const.npz")
    mgc.main()
    from helpers import golden

    old = golden()
    seen = set()
    for path in (tmp_path / "regen.npz", tmp_path / "regen_small.npz", tmp_path / "regen_ek0_const.npz"):
        with np.load(path) as new:
            assert not (seen & set(new.files)), "the fixture files must not share keys"
            seen |= set(new.files)
            for k in new.files:
                np.testing.assert_array_equal(new[k], old[k], err_msg=k)
    assert seen == set(old)
