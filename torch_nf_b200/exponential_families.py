"""Exponential families for EFN training: the consumers of ``(z, log_q_z)`` next to the hot path.

API mirror of the reference's ``torch_nf/exponential_families.py`` (``ExponentialFamily`` :10-101, ``MVN`` :106-222,
``Dirichlet`` :225-308): same class and method names, argument meaning, shapes and error behaviour.  ``eta`` is the
(augmented) natural parameter, ``T(z)`` the sufficient statistic with ``log h(z)`` appended where the base measure is
not constant.  Only ``T`` sits behind a sampled batch ``z (M, N, D)``; it is evaluated with torch ops on the device
``z`` lives on (a handful of elementwise / gather passes over ``M N D_eta`` values - far off the metric) and
is differentiable, so the EFN loss ``mean(log_q - eta . T(z))`` back-propagates into the flow's CUDA backward kernels.
The prior samplers and parameter maps are host-side numpy / scipy, as in the reference.
"""
import numpy as np
import scipy.stats
import torch

from .bijectors import Bijector, ToSimplex
from .error_formatters import format_type_err_msg


def _triu(D, k=0):
    r, c = np.triu_indices(D, k)
    return r, c


class ExponentialFamily(object):
    """Base class (reference exponential_families.py:10-101)."""

    def __init__(self, D, support_layer=None):
        self.D = D
        self.support_layer = support_layer
        self.D_eta = self._get_D_eta()

    @property
    def D(self):
        return self._D

    @D.setter
    def D(self, val):
        if type(val) is not int:
            raise TypeError(format_type_err_msg(self, "D", val, int))
        if val < 1:
            raise ValueError("Exponential family dimensionality must be greater than 1.")
        self._D = val

    @property
    def support_layer(self):
        return self._support_layer

    @support_layer.setter
    def support_layer(self, val):
        # the reference takes a Bijector CLASS here (``issubclass``, :45), e.g. ``ToSimplex`` for the Dirichlet
        if not (val is None or (isinstance(val, type) and issubclass(val, Bijector))):
            raise TypeError(format_type_err_msg(self, "support_layer", val, Bijector))
        self._support_layer = val

    def _get_D_eta(self):
        return self.D

    def sample_eta(self, N):
        raise NotImplementedError()

    def mu_to_eta(self, mu):
        raise NotImplementedError()

    def eta_to_mu(self, eta):
        raise NotImplementedError()

    def T(self, z):
        raise NotImplementedError()


class MVN(ExponentialFamily):
    """Multivariate normal: T(z) = (z, triu(z z^T)), eta = (Sigma^-1 mu, -Sigma^-1 / 2 in minimal form with doubled
    off-diagonals) (reference :106-222)."""

    def __init__(self, D):
        super().__init__(D, None)

    def _get_D_eta(self):
        return int(self.D + self.D * (self.D + 1) // 2)

    def sample_eta(self, N=50, sigma_mu=1., iw_df_fac=5):
        """mu_i ~ N(0, sigma_mu), Sigma ~ IW(df = iw_df_fac D, scale = df I) (:113-137)."""
        mu = np.random.normal(0.0, sigma_mu, (N, self.D))
        df = iw_df_fac * self.D
        Sigma = scipy.stats.invwishart(df=df, scale=df * np.eye(self.D)).rvs(N)
        return self.mu_to_eta(mu, np.reshape(Sigma, (N, self.D, self.D)))

    def T(self, z):
        """(M, N, D) -> (M, N, D + D(D+1)/2): z followed by the upper triangle of z z^T, row-major (:139-156)."""
        r, c = _triu(self.D)
        r = torch.as_tensor(r, device=z.device)
        c = torch.as_tensor(c, device=z.device)
        return torch.cat((z, z.index_select(2, r) * z.index_select(2, c)), dim=2)

    def mu_to_eta(self, mu, Sigma):
        """(N, D), (N, D, D) -> (N, D_eta) (:158-184)."""
        P = np.linalg.inv(Sigma)
        eta1 = np.einsum("nij,nj->ni", P, mu).astype(np.float64)
        r, c = _triu(self.D)
        scale = np.where(r == c, -0.5, -1.0)            # -P/2 on the diagonal, 2 * (-P/2) off it
        return np.concatenate((eta1, P[:, r, c] * scale), axis=1)

    def eta_to_mu(self, eta):
        """(N, D_eta) -> mu (N, D), Sigma (N, D, D) (:186-206)."""
        N, D = eta.shape[0], self.D
        r, c = _triu(D)
        half = np.zeros((N, D, D))
        half[:, r, c] = eta[:, D:]
        eta2 = 0.5 * (half + half.transpose(0, 2, 1))
        Sigma = -0.5 * np.linalg.inv(eta2)
        mu = np.einsum("nij,nj->ni", Sigma, eta[:, :D])
        return mu, Sigma

    def KL(self, z, log_prob, eta):
        """Monte-Carlo KL(q || p_eta) per eta from samples z (M, N, D) with their log q (:208-216)."""
        mu, Sigma = self.eta_to_mu(eta)
        return np.array([np.mean(log_prob[i] - scipy.stats.multivariate_normal(mean=mu[i], cov=Sigma[i]).logpdf(z[i]))
                         for i in range(z.shape[0])])


class Dirichlet(ExponentialFamily):
    """Dirichlet on the simplex: T(z) = (log z, sum log z), eta = (alpha, 1) (reference :225-308)."""

    def __init__(self, D):
        super().__init__(D, ToSimplex)

    def _get_D_eta(self):
        return self.D + 1          # + the log base measure

    def sample_eta(self, N=50, lb=0.5, ub=2.):
        """alpha_i ~ U[lb, ub], a one appended for the base measure (:236-257)."""
        return self.mu_to_eta(np.random.uniform(lb, ub, (N, self.D)))

    def T(self, z):
        """(M, N, D) -> (M, N, D + 1): log(z + 1e-10) and its sum (:259-276)."""
        log_z = torch.log(z + 1e-10)
        return torch.cat((log_z, log_z.sum(dim=2, keepdim=True)), dim=2)

    def mu_to_eta(self, alpha):
        return np.concatenate((alpha, np.ones((alpha.shape[0], 1))), axis=1)

    def eta_to_mu(self, eta):
        return eta[:, :self.D]

    def KL(self, z, log_prob, eta):
        alpha = self.eta_to_mu(eta)
        out = np.zeros((z.shape[0],))
        for i in range(z.shape[0]):
            zi = np.float64(z[i]) + 1e-32
            zi = zi / np.sum(zi, axis=1, keepdims=True)
            out[i] = np.mean(log_prob[i] - scipy.stats.dirichlet(alpha=np.float64(alpha[i])).logpdf(zi.T))
        return out
