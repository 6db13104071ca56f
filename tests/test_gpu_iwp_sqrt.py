"""GPU: the standalone prior / square-root helpers (tornadox_b200.iwp, tornadox_b200.sqrt) against the oracle and the golden
vectors; the reference's own checks of these modules (tests/test_sqrt.py:39-48, 88-116, tests/test_iwp.py:68-93) restated."""

import numpy as np
import pytest
import torch

from helpers import FP32_RTOL, FP64_RTOL, assert_close, golden
from oracle.ref_np import prior as oprior

pytestmark = pytest.mark.gpu


def _t(x, dtype=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=dtype, device="cuda:0")


@pytest.mark.parametrize("n", [2, 3, 4, 5, 6])
@pytest.mark.parametrize("dtype,rtol", [(torch.float64, FP64_RTOL), (torch.float32, FP32_RTOL)])
def test_batched_propagate_cholesky_factor(cuda_device, n, dtype, rtol):
    from tornadox_b200 import sqrt

    rng = np.random.default_rng(n)
    S1, S2 = rng.standard_normal((1000, n, n)), rng.standard_normal((1000, n, n))
    S1[3] = 0.0                                   # a block with zero columns: LAPACK leaves them untouched
    S2[5] = np.tril(S2[5])
    ref = oprior.propagate_cholesky_batched(S1, S2)
    out = sqrt.batched_propagate_cholesky_factor(_t(S1, dtype), _t(S2, dtype)).cpu().numpy().astype(np.float64)
    assert_close(out, ref, rtol, "factor (LAPACK sign convention)")
    assert np.all(np.triu(out, 1) == 0.0)
    # tests/test_sqrt.py:39-48: the factor reproduces the covariance
    assert_close(out @ out.transpose(0, 2, 1), S1 @ S1.transpose(0, 2, 1) + S2 @ S2.transpose(0, 2, 1), 10 * rtol, "covariance")
    # one shared (n, n) process-noise factor for the whole batch (what DiagonalEK1 does)
    shared = sqrt.batched_propagate_cholesky_factor(_t(S1, dtype), _t(S2[0], dtype)).cpu().numpy().astype(np.float64)
    assert_close(shared, oprior.propagate_cholesky_batched(S1, np.broadcast_to(S2[0], S1.shape)), rtol, "shared S2")
    single = sqrt.propagate_cholesky_factor(_t(S1[7], dtype), _t(S2[7], dtype)).cpu().numpy().astype(np.float64)
    assert_close(single, oprior.propagate_cholesky(S1[7], S2[7]), rtol, "single pair")


@pytest.mark.parametrize("n", [2, 4, 5])
def test_sqrtm_to_cholesky(cuda_device, n):
    from tornadox_b200 import sqrt

    rng = np.random.default_rng(10 + n)
    for m in (n, 2 * n):
        St = rng.standard_normal((50, m, n))
        ref = np.stack([oprior.right_sqrt_to_cholesky(x) for x in St])
        out = sqrt.batched_sqrtm_to_cholesky(_t(St)).cpu().numpy()
        assert_close(out, ref, FP64_RTOL, f"sqrtm_to_cholesky ({m} x {n})")
        assert_close(sqrt.sqrtm_to_cholesky(_t(St[0])).cpu().numpy(), ref[0], FP64_RTOL, "single")
    with pytest.raises(ValueError):
        sqrt.batched_sqrtm_to_cholesky(_t(rng.standard_normal((3, n + 1, n))))


@pytest.mark.parametrize("nu", [1, 2, 3, 4, 5])
def test_integrated_wiener_transition_matches_reference(cuda_device, nu):
    from tornadox_b200 import iwp

    g = golden()
    prior = iwp.IntegratedWienerTransition(wiener_process_dimension=3, num_derivatives=nu)
    A, Ql = prior.preconditioned_discretize_1d
    assert np.array_equal(A.cpu().numpy(), g[f"prior/nu{nu}/A"]) and np.array_equal(Ql.cpu().numpy(), g[f"prior/nu{nu}/Ql"])
    for dt in (0.1, 1e-3, 2.5):
        p, pinv = prior.nordsieck_preconditioner_1d_raw(dt)
        assert_close(p.cpu().numpy(), g[f"prior/nu{nu}/p/{dt}"], 1e-14, "p")
        assert_close(pinv.cpu().numpy(), g[f"prior/nu{nu}/pinv/{dt}"], 1e-14, "p_inv")
    assert prior.state_dimension == 3 * (nu + 1)
    dA, dQl = prior.preconditioned_discretize
    rA, rQl = oprior.dense_transition(nu, 3)
    assert_close(dA.cpu().numpy(), rA, 1e-15, "kron(I, A)")
    assert_close(dQl.cpu().numpy(), rQl, 1e-15, "kron(I, Ql)")
    nA, nQl = prior.non_preconditioned_discretize(0.1)
    oA, oQl = oprior.dense_transition_no_precon(nu, 3, 0.1)
    assert_close(nA.cpu().numpy(), oA, 1e-12, "P A P^-1")
    assert_close(nQl.cpu().numpy(), oQl, 1e-12, "P Ql")
    assert_close(prior.projection_matrix(1).cpu().numpy(), oprior.dense_projection(nu, 3, 1), 0.0 + 1e-300, "projection")
    v = torch.arange(prior.state_dimension, dtype=torch.float64, device="cuda:0")
    back = prior.reorder_state_from_coordinate_to_derivative(prior.reorder_state_from_derivative_to_coordinate(v))
    assert torch.equal(back, v)
    assert_close(prior.reorder_state_from_derivative_to_coordinate(v).cpu().numpy(),
                 oprior.derivative_to_coordinate_order(nu, 3, np.arange(prior.state_dimension, dtype=float)), 1e-300, "reorder")
