"""GPU: user-defined vector fields (reference: every f / df_diagonal is a Python callable, odefilter.py:346-352).  The fused
kernels read f and the Jacobian diagonal from arrays the caller's own torch code produced at the predicted state."""

import numpy as np
import pytest
import torch

from helpers import assert_close, make_pair
from oracle.ref_np import driver as odriver
from oracle.ref_np import filters as ofilters

pytestmark = pytest.mark.gpu


def _lorenz96_torch(forcing=8.0):
    def f(_, y):
        return (torch.roll(y, -1) - torch.roll(y, 2)) * torch.roll(y, 1) - y + forcing      # ivp.py:275-277

    def df_diagonal(_, y):
        return -torch.ones_like(y)                                                          # ivp.py:290-291

    return f, df_diagonal


def _vanderpol_torch(mu=1.0):
    def f(_, y):
        return torch.stack((y[1], mu * (1.0 - y[0] ** 2) * y[1] - y[0]))                     # ivp.py:34-35

    def df_diagonal(_, y):
        return torch.stack((torch.zeros_like(y[0]), mu * (1.0 - y[0] ** 2)))                 # ivp.py:43-44

    return f, df_diagonal


@pytest.mark.parametrize("solver_name", ["ek1", "ek0", "dek0"])
@pytest.mark.parametrize("kind,size,tmax", [("lorenz96", 300, 0.5), ("vanderpol", 1, 1.0)])
def test_user_defined_f_reproduces_the_in_kernel_field(cuda_device, solver_name, kind, size, tmax):
    """The same ODE once as an in-kernel vector field and once as torch callables: same accepted steps, same states."""
    from tornadox_b200 import ek0, ek1, ivp, step

    _, p, _ = make_pair(kind, size, tmax=tmax) if kind == "vanderpol" else make_pair(kind, size)
    p = p._replace(tmax=tmax)
    f, dfd = _lorenz96_torch() if kind == "lorenz96" else _vanderpol_torch()
    user = ivp.InitialValueProblem(f=f, t0=p.t0, tmax=tmax, y0=p.y0, df=None, df_diagonal=dfd)
    cls = {"ek0": ek0.KroneckerEK0, "ek1": ek1.DiagonalEK1, "dek0": ek0.DiagonalEK0}[solver_name]
    ref_sol = cls(num_derivatives=3, steprule=step.AdaptiveSteps())
    ref_sol.device_loop = False
    ref = ref_sol.solve(p)
    got = cls(num_derivatives=3, steprule=step.AdaptiveSteps()).solve(user)
    assert got.info == ref.info
    np.testing.assert_allclose(got.t.numpy(), ref.t.numpy(), rtol=1e-10, atol=0)
    assert_close(got.mean.cpu().numpy(), ref.mean.cpu().numpy(), 1e-9, "means of all accepted states")


def test_taylor_mode_of_a_torch_callable(cuda_device):
    """init.TaylorMode on a user-defined f (nested forward-mode derivatives) against the truncated power series of the same field."""
    from tornadox_b200 import init

    for kind, size in (("lorenz96", 50), ("vanderpol", 1)):
        o, p, series_f = make_pair(kind, size)
        f, _ = _lorenz96_torch() if kind == "lorenz96" else _vanderpol_torch()
        for nu in (2, 4, 5):
            m_ref, _ = odriver.taylor_init(series_f, o.y0, o.t0, nu)
            m, c = init.TaylorMode()(f=f, df=None, y0=p.y0, t0=p.t0, num_derivatives=nu)
            assert c.shape == (nu + 1, nu + 1) and float(c.abs().max()) == 0.0
            assert_close(m.cpu().numpy(), m_ref, 1e-12, f"{kind} nu={nu}")


def test_user_defined_f_against_the_oracle(cuda_device):
    """A field the library has never heard of (logistic growth with a time-dependent rate), DiagonalEK1 and KroneckerEK0
    against the NumPy oracle driven by the same function: step by step."""
    from tornadox_b200 import ek0, ek1, ivp, odefilter, rv, step

    d, nu = 777, 3
    rng = np.random.default_rng(5)
    y0 = 0.1 + 0.8 * rng.uniform(size=d)
    rate = 1.0 + rng.uniform(size=d)
    rate_t = torch.as_tensor(rate, device="cuda")

    def f(t, y):
        return (1.0 + 0.5 * t) * rate_t * y * (1.0 - y)

    def dfd(t, y):
        return (1.0 + 0.5 * t) * rate_t * (1.0 - 2.0 * y)

    def f_np(t, y):
        return (1.0 + 0.5 * t) * rate * y * (1.0 - y)

    def dfd_np(t, y):
        return (1.0 + 0.5 * t) * rate * (1.0 - 2.0 * y)

    prob = ivp.InitialValueProblem(f=f, t0=0.0, tmax=1.0, y0=y0, df=None, df_diagonal=dfd)
    n = nu + 1
    mean = rng.standard_normal((n, d))
    mean[0] = y0
    blocks = 1e-2 * rng.standard_normal((d, n, n))
    single = 1e-2 * rng.standard_normal((n, n))
    dt, t = 0.05, 0.3
    dev = torch.device("cuda")
    for name in ("dek1", "ek0"):
        cls = ek1.DiagonalEK1 if name == "dek1" else ek0.KroneckerEK0
        sol = cls(num_derivatives=nu, steprule=step.AdaptiveSteps())
        sol.initialize(*prob)
        if name == "dek1":
            y = rv.BatchedMultivariateNormal(torch.as_tensor(mean, device=dev), torch.as_tensor(blocks, device=dev))
            ost, _ = ofilters.diagonal_ek1_attempt(ofilters.FilterState(t, mean, blocks, None, None), dt, f_np, dfd_np, nu)
        else:
            y = rv.LeftIsotropicMatrixNormal(torch.as_tensor(mean, device=dev), d, torch.as_tensor(single, device=dev))
            ost, _ = ofilters.kronecker_ek0_attempt(ofilters.FilterState(t, mean, single, None, None), dt, f_np, nu)
        st, _ = sol.attempt_step(odefilter.ODEFilterState(t, y, None, None), dt, *prob)
        assert_close(st.y.mean.cpu().numpy(), ost.mean, 1e-10, f"{name} mean")
        fac = st.y.cov_sqrtm if name == "dek1" else st.y.cov_sqrtm_2
        assert_close(fac.cpu().numpy(), ost.cov_sqrtm, 1e-10, f"{name} factor")
        norm = step.norm_from_sq_sum(st.scaled_error_norm_sq_sum, d)
        ref_norm = odriver.AdaptiveSteps().scale_error_estimate(dt * ost.error_estimate, ost.reference_state)
        assert abs(norm - ref_norm) <= 1e-10 * abs(ref_norm)
    # and a whole adaptive solve: close to the closed form where the rate is constant in time
    const = ivp.InitialValueProblem(f=lambda t, y: rate_t * y * (1.0 - y), t0=0.0, tmax=1.0, y0=y0, df=None,
                                    df_diagonal=lambda t, y: rate_t * (1.0 - 2.0 * y))
    fin, info = ek1.DiagonalEK1(num_derivatives=4, steprule=step.AdaptiveSteps(abstol=1e-8, reltol=1e-6)).simulate_final_state(const)
    exact = 1.0 / (1.0 + (1.0 / y0 - 1.0) * np.exp(-rate * 1.0))
    assert info["num_steps"] > 3 and np.max(np.abs(fin.y.mean[0].cpu().numpy() - exact)) < 1e-4
