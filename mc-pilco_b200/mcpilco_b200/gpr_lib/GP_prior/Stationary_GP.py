"""Stationary kernels — the reference's `gpr_lib/GP_prior/Stationary_GP.py` surface (Stationary_GP :14-109, RBF :112-181).

k(x, x') = lambda * exp(-sum_j ((x_j - x'_j) / l_j)^2)   — NO factor 1/2 in the exponent (reference :162-170),
with l = exp(log_lengthscales_par), lambda = exp(log_lambda_par), constant prior mean mean_par.
"""
import numpy as np
import torch

from ... import _pack as P
from . import GP_prior


class Stationary_GP(GP_prior.GP_prior):
    def __init__(self, active_dims, lengthscales_init=None, flg_train_lengthscales=True, sigma_n_init=None, flg_train_sigma_n=True,
                 name="", dtype=torch.float64, sigma_n_num=None, device=None):
        super().__init__(active_dims, sigma_n_init=sigma_n_init, flg_train_sigma_n=flg_train_sigma_n, name=name, dtype=dtype,
                         sigma_n_num=sigma_n_num, device=device)
        if active_dims is None:
            raise RuntimeError("Stationary_GP obj require active_dims")
        self.num_features = np.asarray(active_dims).size
        if lengthscales_init is None:
            lengthscales_init = np.ones(self.num_features)
        lengthscales_init = np.asarray(lengthscales_init, dtype=np.float64)
        self.flg_ARD = lengthscales_init.size != 1
        self.log_lengthscales_par = torch.nn.Parameter(torch.tensor(np.log(lengthscales_init), dtype=self.dtype, device=self.device),
                                                       requires_grad=flg_train_lengthscales)

    def get_weigted_distances(self, X1, X2):
        """sum_j ((x_j - x'_j)/l_j)^2, recovered from the native SE covariance of a unit-lambda copy of this kernel."""
        from ... import _ops as ops
        spec = P.new_gp_spec(X1.shape[1])
        ls = self.log_lengthscales_par if self.flg_ARD else self.log_lengthscales_par.expand(self.num_features)
        P.add_se(spec, P._np(self.active_dims), ls, 0.0, 0.0)
        return -torch.log(ops.gp_covariance(spec, X1, X2))


class RBF(Stationary_GP):
    """Squared-exponential GP with constant mean (reference :112-181)."""

    def __init__(self, active_dims, lengthscales_init=None, flg_train_lengthscales=True, sigma_n_init=None, flg_train_sigma_n=True,
                 lambda_init=None, flg_train_lambda=True, mean_init=None, flg_train_mean=False, name="", dtype=torch.float64,
                 sigma_n_num=None, device=None):
        super().__init__(active_dims, lengthscales_init=lengthscales_init, flg_train_lengthscales=flg_train_lengthscales,
                         sigma_n_init=sigma_n_init, flg_train_sigma_n=flg_train_sigma_n, name=name, dtype=dtype, sigma_n_num=sigma_n_num,
                         device=device)
        lambda_init = np.ones(1) if lambda_init is None else np.asarray(lambda_init, dtype=np.float64)
        if lambda_init.size != 1:
            raise RuntimeError("Lambda must be a np array qith dimension 1")
        self.log_lambda_par = torch.nn.Parameter(torch.tensor(np.log(lambda_init), dtype=self.dtype, device=self.device),
                                                 requires_grad=flg_train_lambda)
        mean_init = np.zeros(1) if mean_init is None else np.asarray(mean_init, dtype=np.float64)
        self.mean_par = torch.nn.Parameter(torch.tensor(mean_init, dtype=self.dtype, device=self.device), requires_grad=flg_train_mean)

    def _kernel_terms(self, D):
        inv = torch.exp(-self.log_lengthscales_par).reshape(-1)
        if not self.flg_ARD:
            inv = inv.expand(self.num_features)
        return {"inv_ls": self._scatter(inv, D), "lam": torch.exp(self.log_lambda_par).reshape(()), "mean": self.mean_par.reshape(-1)[0],
                "polys": []}
