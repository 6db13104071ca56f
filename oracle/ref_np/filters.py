"""ORACLE (test infrastructure): NumPy fp64 restatement of the filter steps of the reference.

* ``kronecker_ek0_attempt``  -- ``KroneckerEK0.attempt_step``              tornadox/ek0.py:152-193
* ``diagonal_ek1_attempt``   -- ``BatchedEK1.attempt_step`` + ``DiagonalEK1``  tornadox/ek1.py:228-369
* ``reference_ek0_attempt``  -- dense ``ReferenceEK0.attempt_step``         tornadox/ek0.py:47-82
* ``reference_ek1_attempt``  -- dense ``ReferenceEK1.attempt_step``         tornadox/ek1.py:46-133

States are plain tuples/namedtuples of NumPy arrays with the reference's layouts:
mean (n, d) C-order; DiagonalEK1 factors (d, n, n); KroneckerEK0 factor (n, n).
"""

from collections import namedtuple

import numpy as np
import scipy.linalg

from . import prior

FilterState = namedtuple("FilterState", "t mean cov_sqrtm error_estimate reference_state")


# ---------------------------------------------------------------------------------------------
# KroneckerEK0
# ---------------------------------------------------------------------------------------------


def kronecker_ek0_initial_state(mean0, cov_sqrtm0, t0):
    """tornadox/ek0.py:94-121 -- scalar NaN error, NaN reference state."""
    d = mean0.shape[1]
    return FilterState(t0, mean0, cov_sqrtm0, np.nan, np.nan * np.ones(d))


def kronecker_ek0_attempt(state, dt, f, nu, sum_over_shards=None, d_global=None):
    """One ``KroneckerEK0.attempt_step`` -- tornadox/ek0.py:152-193.

    ``sum_over_shards`` / ``d_global``: when the state is one shard of a larger problem (bench.py's CPU arm cuts the grid into
    row strips, one per core) the calibration needs z^T z and the dimension of the WHOLE problem (ek0.py:126-128); the callable
    turns the local sum into the global one.  Default: the state is the whole problem."""
    A, Ql = prior.transition_1d(nu), prior.diffusion_sqrt_1d(nu)
    e1 = prior.unit_row(nu, 1)
    _m, _Cl = state.mean, state.cov_sqrtm
    t_new = state.t + dt

    # [Preconditioners]  ek0.py:161-163
    P, PI = prior.nordsieck_diag(nu, dt)
    m = PI @ _m
    Cl = PI @ _Cl

    # [Predict]  ek0.py:166
    mp = A @ m

    # [Measure]  ek0.py:142-150
    _mp = P @ mp
    z = _mp[1] - f(t_new, _mp[0])
    H = e1 @ P

    # [Calibration]  ek0.py:123-130
    Q11 = Ql[1, :] @ Ql[1, :].T
    HQH = P[1, 1] ** 2 * Q11
    z_sq = z.T @ z if sum_over_shards is None else sum_over_shards(z.T @ z)
    sigma_squared = z_sq / HQH / (z.shape[0] if d_global is None else d_global)
    error_estimate = np.sqrt(sigma_squared * HQH)

    # [Predict covariance]  ek0.py:175
    Clp = prior.propagate_cholesky(A @ Cl, np.sqrt(sigma_squared) * Ql)

    # [Update]  ek0.py:132-140
    C11 = Clp[1, :] @ Clp[1, :].T
    S = P[1, 1] ** 2 * C11
    K = Clp @ (Clp.T @ H.T) / S
    m_new = mp - K * z[None, :]
    Cl_new = Clp - K @ H @ Clp

    # [Undo preconditioning]  ek0.py:181-184
    _m_new = P @ m_new
    _Cl_new = P @ Cl_new
    y_new = np.abs(_m_new[0])

    new_state = FilterState(t_new, _m_new, _Cl_new, error_estimate, y_new)
    return new_state, dict(num_f_evaluations=1)


# ---------------------------------------------------------------------------------------------
# DiagonalEK1 (BatchedEK1 base + DiagonalEK1 unit step)
# ---------------------------------------------------------------------------------------------


def batched_initial_state(mean0, cov_sqrtm0, t0):
    """tornadox/ek1.py:216-226 -- stack the (n, n) factor d times, NaN error/reference."""
    d = mean0.shape[1]
    return FilterState(
        t0, mean0, np.stack([cov_sqrtm0] * d), np.nan * np.ones(d), np.nan * np.ones(d)
    )


def _dek1_h(p, J, blocks):
    """h_i = p[1] * blocks[i,1,:] - J_i p[0] * blocks[i,0,:] -- ek1.py:318-323 / 337-342 / 355-360."""
    no_precon = p[None, :, None] * blocks
    return no_precon[:, 1, :] - J[:, None] * no_precon[:, 0, :]


def diagonal_ek1_attempt(state, dt, f, df_diagonal, nu, return_intermediates=False):
    """One ``DiagonalEK1`` attempt -- tornadox/ek1.py:228-260 wrapping ek1.py:274-369."""
    phi_1d, sq_1d = prior.transition_1d(nu), prior.diffusion_sqrt_1d(nu)
    d = state.mean.shape[1]
    batched_sq = np.broadcast_to(sq_1d, (d,) + sq_1d.shape)  # ek1.py:213 (stack -> broadcast)

    # ek1.py:231-235
    p, p_inv = prior.nordsieck_raw(nu, dt)
    m = p_inv[:, None] * state.mean
    sc = p_inv[None, :, None] * state.cov_sqrtm
    t = state.t + dt

    # predict_mean  ek1.py:262-265
    m_pred = phi_1d @ m

    # evaluate_ode  ek1.py:302-312
    m_pred_no_precon = p[:, None] * m_pred
    m_at = m_pred_no_precon[0]
    fx = f(t, m_at)
    z = m_pred_no_precon[1] - fx
    J = df_diagonal(t, m_at)

    # estimate_error  ek1.py:314-331
    h_sq = _dek1_h(p, J, batched_sq)
    s = np.einsum("dn,dn->d", h_sq, h_sq)
    xi = z / np.sqrt(s)
    sigma = np.abs(xi)
    error = sigma * np.sqrt(s)

    # predict_cov_sqrtm  ek1.py:267-270
    sc_pred = prior.propagate_cholesky_batched(phi_1d @ sc, sigma[:, None, None] * batched_sq)

    # observe_cov_sqrtm  ek1.py:333-349
    h_sc = _dek1_h(p, J, sc_pred)
    s2 = np.einsum("dn,dn->d", h_sc, h_sc)
    s2 = s2 + 1e-16
    cross = sc_pred @ h_sc[..., None]
    kgain = cross / s2[..., None, None]

    # correct_cov_sqrtm  ek1.py:351-362
    new_sc = sc_pred - kgain @ h_sc[:, None, :]

    # correct_mean  ek1.py:364-369
    correction = kgain @ z[:, None, None]
    new_mean = m_pred - correction[:, :, 0].T

    # ek1.py:246-251
    new_mean = p[:, None] * new_mean
    cov_sqrtm = p[None, :, None] * new_sc
    reference_state = np.maximum(np.abs(state.mean[0]), np.abs(new_mean[0]))

    new_state = FilterState(t, new_mean, cov_sqrtm, error, reference_state)
    info = dict(num_f_evaluations=1, num_df_diagonal_evaluations=1)
    if return_intermediates:
        return new_state, info, dict(m_at=m_at, fx=fx, z=z, J=J, sigma=sigma, sc_pred=sc_pred, kgain=kgain)
    return new_state, info


def diagonal_ek0_attempt(state, dt, f, nu):
    """``DiagonalEK0`` (tornadox/ek0.py:196-280) == DiagonalEK1 with J = 0 and no df_diagonal call."""
    d = state.mean.shape[1]
    new_state, _ = diagonal_ek1_attempt(state, dt, f, lambda t, y: np.zeros(d), nu)
    return new_state, dict(num_f_evaluations=1)


# ---------------------------------------------------------------------------------------------
# Dense reference solvers (small d; cross-checks as in tests/test_ek0.py:246-263, test_ek1.py)
# ---------------------------------------------------------------------------------------------


def dense_initial_state(mean0, cov_sqrtm0, t0):
    """tornadox/ek0.py:36-45 / ek1.py:35-44 -- kron(I_d, cov_sqrtm0)."""
    d = mean0.shape[1]
    return FilterState(
        t0, mean0, np.kron(np.eye(d), cov_sqrtm0), np.nan * np.ones(d), np.nan * np.ones(d)
    )


def reference_ek0_attempt(state, dt, f, nu):
    """Dense ``ReferenceEK0.attempt_step`` -- tornadox/ek0.py:47-82."""
    n, d = state.mean.shape
    m, Cl = state.mean.reshape((-1,), order="F"), state.cov_sqrtm
    A, Ql = prior.dense_transition_no_precon(nu, d, dt)
    E0, E1 = prior.dense_projection(nu, d, 0), prior.dense_projection(nu, d, 1)

    mp = A @ m
    z = E1 @ mp - f(state.t + dt, E0 @ mp)
    H = E1
    S = H @ Ql @ Ql.T @ H.T
    sigma_squared = z @ np.linalg.solve(S, z) / z.shape[0]
    sigma = np.sqrt(sigma_squared)
    error = np.sqrt(np.diag(S)) * sigma
    Clp = prior.propagate_cholesky(A @ Cl, sigma * Ql)
    Cl_new, K, _Sl = prior.qr_update(H, Clp)
    m_new = (mp - K @ z).reshape((n, d), order="F")
    y_new = np.abs(m_new[0])
    return FilterState(state.t + dt, m_new, Cl_new, error, y_new), dict(num_f_evaluations=1)


def reference_ek1_attempt(state, dt, f, df, nu):
    """Dense ``ReferenceEK1.attempt_step`` -- tornadox/ek1.py:46-133."""
    n, d = state.mean.shape
    P, Pinv = prior.dense_nordsieck(nu, d, dt)
    A, SQ = prior.dense_transition(nu, d)
    E0, E1 = prior.dense_projection(nu, d, 0), prior.dense_projection(nu, d, 1)
    t = state.t + dt
    m, SC = Pinv @ state.mean.reshape((-1,), order="F"), Pinv @ state.cov_sqrtm

    # attempt_unit_step  ek1.py:82-100
    m_pred = A @ m
    P0, P1 = E0 @ P, E1 @ P  # evaluate_ode  ek1.py:114-124
    m_at = P0 @ m_pred
    fx = f(t, m_at)
    Jx = df(t, m_at)
    H = P1 - Jx @ P0
    b = Jx @ m_at - fx
    z = H @ m_pred + b
    # estimate_error  ek1.py:126-133
    s_chol = prior.right_sqrt_to_cholesky((H @ SQ).T)
    whitened = scipy.linalg.solve_triangular(s_chol.T, z, lower=False)
    sigma = np.sqrt(whitened.T @ whitened / whitened.shape[0])
    error_estimate = sigma * np.sqrt(np.diag(s_chol @ s_chol.T))
    SC_pred = prior.propagate_cholesky(A @ SC, sigma * SQ)
    cov_cholesky, Kgain, _ = prior.qr_update(H, SC_pred)
    new_mean = m_pred - Kgain @ z

    cov_cholesky = P @ cov_cholesky
    new_mean = (P @ new_mean).reshape((n, d), order="F")
    reference_state = np.maximum(np.abs(state.mean[0]), np.abs(new_mean[0]))
    new_state = FilterState(t, new_mean, cov_cholesky, error_estimate, reference_state)
    return new_state, dict(num_f_evaluations=1, num_df_evaluations=1)
