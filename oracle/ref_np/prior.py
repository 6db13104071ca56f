"""ORACLE (test infrastructure, never imported by the product).

NumPy fp64 restatement of the integrated-Wiener-process prior and of the QR-based
square-root arithmetic of the reference.  File:line citations are into /root/reference.

``numpy.linalg.qr(mode="r")`` and ``numpy.linalg.cholesky`` are LAPACK ``dgeqrf`` / ``dpotrf``,
the routines ``jnp.linalg.qr`` / ``jnp.linalg.cholesky`` dispatch to on CPU, so factor signs
follow the reference's convention.
"""

import math

import numpy as np
import scipy.linalg


# ---------------------------------------------------------------------------------------------
# IWP(nu) prior, one coordinate            (tornadox/iwp.py)
# ---------------------------------------------------------------------------------------------


def transition_1d(nu):
    """Preconditioned 1-d transition ``A`` -- tornadox/iwp.py:31-35.

    flip(lower Pascal) in both axes, i.e. the unit upper-triangular matrix A[i,j] = C(nu-i, nu-j).
    """
    n = nu + 1
    return np.flip(scipy.linalg.pascal(n, kind="lower", exact=False)).astype(np.float64)


def diffusion_sqrt_1d(nu):
    """Preconditioned 1-d process-noise Cholesky factor ``Ql`` -- tornadox/iwp.py:36-37.

    Q = flip(Hilbert(n)), i.e. Q[i,j] = 1/(2 nu + 1 - i - j);  Ql = chol(Q) (lower).
    """
    n = nu + 1
    return np.linalg.cholesky(np.flip(scipy.linalg.hilbert(n)))


def nordsieck_raw(nu, dt):
    """(p, p_inv) -- tornadox/iwp.py:65-73.

    p[k] = |dt|**(nu-k+1/2) / (nu-k)!   and   p_inv[k] = |dt|**-(nu-k+1/2) * (nu-k)!.
    """
    orders = np.arange(nu, -1, -1)
    fact = np.array([float(math.factorial(int(q))) for q in orders])
    expo = orders + 0.5
    return (np.abs(dt) ** expo) / fact, (np.abs(dt) ** (-expo)) * fact


def nordsieck_diag(nu, dt):
    """Diagonal-matrix form -- tornadox/iwp.py:75-89."""
    p, p_inv = nordsieck_raw(nu, dt)
    return np.diag(p), np.diag(p_inv)


def unit_row(nu, q):
    """e_q as a (1, n) row -- tornadox/iwp.py:145-147."""
    return np.eye(1, nu + 1, q)


def dense_transition(nu, d):
    """kron(I_d, A), kron(I_d, Ql) -- tornadox/iwp.py:41-63 (dense oracles only)."""
    eye = np.eye(d)
    return np.kron(eye, transition_1d(nu)), np.kron(eye, diffusion_sqrt_1d(nu))


def dense_nordsieck(nu, d, dt):
    """kron(I_d, P), kron(I_d, P^-1) -- tornadox/iwp.py:91-110."""
    P, PI = nordsieck_diag(nu, dt)
    eye = np.eye(d)
    return np.kron(eye, P), np.kron(eye, PI)


def dense_transition_no_precon(nu, d, dt):
    """P A P^-1 and P Ql -- tornadox/iwp.py:112-136."""
    P, PI = dense_nordsieck(nu, d, dt)
    A, Ql = dense_transition(nu, d)
    return P @ A @ PI, P @ Ql


def dense_projection(nu, d, q):
    """kron(I_d, e_q) -- tornadox/iwp.py:139-142."""
    return np.kron(np.eye(d), unit_row(nu, q))


def derivative_to_coordinate_order(nu, d, vec):
    """tornadox/iwp.py:155-160."""
    total = d * (nu + 1)
    idx = np.concatenate([np.arange(i, total, d) for i in range(d)])
    return vec[idx]


def coordinate_to_derivative_order(nu, d, vec):
    """tornadox/iwp.py:162-167."""
    total = d * (nu + 1)
    idx = np.concatenate([np.arange(i, total, nu + 1) for i in range(nu + 1)])
    return vec[idx]


# ---------------------------------------------------------------------------------------------
# square-root propagation                 (tornadox/sqrt.py)
# ---------------------------------------------------------------------------------------------


def right_sqrt_to_cholesky(St):
    """R-factor of a 'right' square root, transposed -- tornadox/sqrt.py:15-23."""
    return np.linalg.qr(St, mode="r").T


def propagate_cholesky(S1, S2):
    """Lower factor L with L L^T = S1 S1^T + S2 S2^T -- tornadox/sqrt.py:8-12."""
    return right_sqrt_to_cholesky(np.vstack((S1.T, S2.T)))


def propagate_cholesky_batched(S1, S2):
    """vmap of the above over the leading axis -- tornadox/sqrt.py:27-30."""
    stacked = np.concatenate((np.swapaxes(S1, 1, 2), np.swapaxes(S2, 1, 2)), axis=1)
    return np.swapaxes(np.linalg.qr(stacked, mode="r"), 1, 2)


def qr_update(H, C_chol):
    """Noise-free square-root measurement update -- tornadox/sqrt.py:33-70.

    Returns (posterior Cholesky, Kalman gain, innovation Cholesky).  Only the dense
    Reference* oracles use it; EK0/DiagonalEK1 use closed-form rank-one updates.
    """
    k, m = H.shape
    block = np.block([[C_chol.T @ H.T, C_chol.T], [np.zeros((k, m)).T, np.zeros((m, m)).T]])
    (R,) = scipy.linalg.qr(block, mode="r", pivoting=False)
    R1, R2, R3 = R[:k, :k], R[:k, k:], R[k : k + m, k : k + m]
    gain = scipy.linalg.solve_triangular(R1, R2, lower=False).T
    return R3.T, gain, R1.T
