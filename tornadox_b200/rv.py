"""State containers (device tensors).  Mirrors tornadox/rv.py:9-73."""

from collections import namedtuple

import torch


class MultivariateNormal(namedtuple("_MultivariateNormal", "mean cov_sqrtm")):
    """Dense Gaussian (rv.py:9-13)."""

    @property
    def cov(self):
        return self.cov_sqrtm @ self.cov_sqrtm.T


class BatchedMultivariateNormal(namedtuple("_BatchedMultivariateNormal", "mean cov_sqrtm")):
    """mean (n, d), one (n, n) factor per coordinate: cov_sqrtm (d, n, n)  (rv.py:16-23)."""

    @property
    def cov(self):
        return self.cov_sqrtm @ self.cov_sqrtm.transpose(1, 2)


class MatrixNormal(namedtuple("_MatrixNormal", "mean cov_sqrtm_1 cov_sqrtm_2")):
    """Matrix-variate normal with Kronecker covariance (rv.py:26-41)."""

    @property
    def cov_1(self):
        return self.cov_sqrtm_1 @ self.cov_sqrtm_1.T

    @property
    def cov_2(self):
        return self.cov_sqrtm_2 @ self.cov_sqrtm_2.T

    def dense_cov(self):
        return torch.kron(self.cov_1, self.cov_2)

    def dense_cov_sqrtm(self):
        return torch.kron(self.cov_sqrtm_1, self.cov_sqrtm_2)


class LeftIsotropicMatrixNormal(namedtuple("_LeftIsotropicMatrixNormal", "mean d cov_sqrtm_2")):
    """EK0 state: identity left factor stored as its dimension only (rv.py:44-73)."""

    def _eye(self):
        return torch.eye(self.d, dtype=self.cov_sqrtm_2.dtype, device=self.cov_sqrtm_2.device)

    @property
    def cov_sqrtm_1(self):
        return self._eye()

    @property
    def cov_1(self):
        return self._eye()

    @property
    def cov_2(self):
        return self.cov_sqrtm_2 @ self.cov_sqrtm_2.T

    @property
    def cov(self):
        return self.cov_2.unsqueeze(0).expand(self.d, -1, -1)

    def dense_cov(self):
        return torch.kron(self.cov_1, self.cov_2)

    def dense_cov_sqrtm(self):
        return torch.kron(self.cov_sqrtm_1, self.cov_sqrtm_2)
