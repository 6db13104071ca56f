"""Integrated Wiener process prior.  Mirrors ``tornadox.iwp.IntegratedWienerTransition`` (iwp.py:13-167).

The solvers never materialise these matrices: A = flip(Pascal) and Ql = chol(flip(Hilbert)) are compile-time / per-launch
constants of the step kernels and the Nordsieck vectors travel as kernel arguments.  This class hands out the same objects
(from the library's own ``tdx_prior`` / ``tdx_nordsieck``) as device tensors for code written against the reference; the
``kron(I_d, .)`` variants are dense O(d^2) objects and only meant for small problems, as upstream.
"""

from collections import namedtuple

import numpy as np
import torch

from . import _lib


class IntegratedWienerTransition(namedtuple("_IWP", "wiener_process_dimension num_derivatives")):
    def _t(self, a, device=None, dtype=torch.float64):
        return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device=device or "cuda")

    @property
    def preconditioned_discretize_1d(self):
        """(A, Ql) of one coordinate (iwp.py:19-37); Ql is LAPACK's dpotrf factor, as the reference's."""
        A, _ = _lib.prior(self.num_derivatives)
        return self._t(A), self._t(_lib.lapack_diffusion_sqrt(self.num_derivatives))

    @property
    def preconditioned_discretize(self):
        """kron(I_d, A), kron(I_d, Ql) (iwp.py:41-63)."""
        A, Ql = self.preconditioned_discretize_1d
        eye = torch.eye(self.wiener_process_dimension, dtype=A.dtype, device=A.device)
        return torch.kron(eye, A), torch.kron(eye, Ql)

    def nordsieck_preconditioner_1d_raw(self, dt):
        """(p, p_inv) (iwp.py:65-73)."""
        p, pinv = _lib.nordsieck(self.num_derivatives, float(dt))
        return self._t(p), self._t(pinv)

    def nordsieck_preconditioner_1d(self, dt):
        """diag(p), diag(p_inv) (iwp.py:75-89)."""
        p, pinv = self.nordsieck_preconditioner_1d_raw(dt)
        return torch.diag(p), torch.diag(pinv)

    def nordsieck_preconditioner(self, dt):
        """kron(I_d, diag(p)), kron(I_d, diag(p_inv)) (iwp.py:91-110)."""
        P, PI = self.nordsieck_preconditioner_1d(dt)
        eye = torch.eye(self.wiener_process_dimension, dtype=P.dtype, device=P.device)
        return torch.kron(eye, P), torch.kron(eye, PI)

    def non_preconditioned_discretize(self, dt):
        """P A P^-1 and P Ql (iwp.py:112-136)."""
        P, PI = self.nordsieck_preconditioner(dt)
        A, Ql = self.preconditioned_discretize
        return P @ A @ PI, P @ Ql

    def projection_matrix(self, derivative_to_project_onto):
        """kron(I_d, e_p) (iwp.py:139-142)."""
        e = self.projection_matrix_1d(derivative_to_project_onto)
        return torch.kron(torch.eye(self.wiener_process_dimension, dtype=e.dtype, device=e.device), e)

    def projection_matrix_1d(self, derivative_to_project_onto):
        """e_p as a (1, n) row (iwp.py:145-147)."""
        e = torch.zeros((1, self.num_derivatives + 1), dtype=torch.float64, device="cuda")
        e[0, derivative_to_project_onto] = 1.0
        return e

    @property
    def state_dimension(self):
        return self.wiener_process_dimension * (self.num_derivatives + 1)

    def reorder_state_from_derivative_to_coordinate(self, state):
        """iwp.py:155-160."""
        d, dim = self.wiener_process_dimension, self.state_dimension
        idx = torch.cat([torch.arange(i, dim, d, device=state.device) for i in range(d)])
        return state[idx]

    def reorder_state_from_coordinate_to_derivative(self, state):
        """iwp.py:162-167."""
        n, dim = self.num_derivatives, self.state_dimension
        idx = torch.cat([torch.arange(i, dim, n + 1, device=state.device) for i in range(n + 1)])
        return state[idx]
