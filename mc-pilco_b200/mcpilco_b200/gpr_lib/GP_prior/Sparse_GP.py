"""Linear / multiplicative-polynomial kernels — the reference's `gpr_lib/GP_prior/Sparse_GP.py` surface used by the
rollout: Linear_GP (:295-490), MPK_GP (:559-668), get_Volterra_MPK_GP (:671-737).

  Linear_GP:  k(x, x') = phi(x)^T Sigma phi(x'),  phi = [x[active], 1?]
  MPK_GP:     k(x, x') = prod_{d < P} phi(x)^T Sigma_d phi(x'),   Sigma_d = diag(((P - d) * exp(p_d))**2)
              — the reference's get_Sigma loop (:613-623) adds the SAME slice p_d for every remaining degree, hence the
              factor (P - d); reproduced, not "fixed".
  Volterra:   sum over degrees 1..P of MPK terms; only the first has the offset column and the noise.

The native kernels take diagonal Sigma matrices, which is what every MC-PILCO configuration uses
(`diagonal_covariance`); a Sigma_function returning a non-diagonal matrix is rejected loudly.
Out of scope: get_SOR_GP (broken in the reference itself), Poly_GP.
"""
import numpy as np
import torch

from ... import _pack as P
from ..Utils import Parameters_covariance_functions as _PCF
from . import GP_prior


def get_pos_par_sqrt(par):
    return par ** 2


def f_init_pos_par_sqrt(par):
    return np.sqrt(par)


def get_pos_par_log(par):
    return torch.exp(par)


def f_init_pos_par_log(par):
    return np.log(par)


class Linear_GP(GP_prior.GP_prior):
    """Dot-product kernel with a parametrised weight covariance (reference :295-490)."""

    def __init__(self, active_dims, mean_init=None, flg_mean_trainable=False, flg_no_mean=False, sigma_n_init=None,
                 flg_train_sigma_n=True, Sigma_function=None, Sigma_f_additional_par_list=None, Sigma_pos_par_init=None,
                 flg_train_Sigma_pos_par=True, Sigma_free_par_init=None, flg_train_Sigma_free_par=True, flg_offset=False,
                 f_transofrm_pos_par=get_pos_par_log, f_init_pos_par=f_init_pos_par_log, name="", dtype=torch.float64,
                 sigma_n_num=None, device=None):
        super().__init__(active_dims, sigma_n_init=sigma_n_init, flg_train_sigma_n=flg_train_sigma_n, name=name, dtype=dtype,
                         sigma_n_num=sigma_n_num, device=device)
        if active_dims is None:
            raise RuntimeError("Active_dims are needed")
        self.num_features = np.asarray(active_dims).size
        self.flg_offset = flg_offset
        self.f_transofrm_pos_par = f_transofrm_pos_par
        self.f_init_pos_par = f_init_pos_par
        if mean_init is None:
            mean_init, flg_no_mean = np.zeros(1), True
        self.flg_no_mean = flg_no_mean
        self.mean_par = torch.nn.Parameter(torch.tensor(mean_init, dtype=self.dtype, device=self.device), requires_grad=flg_mean_trainable)
        if Sigma_function is None:
            raise RuntimeError("Specify a Sigma function")
        self.Sigma_function = Sigma_function
        self.Sigma_f_additional_par_list = Sigma_f_additional_par_list
        self.Sigma_pos_par = None if Sigma_pos_par_init is None else torch.nn.Parameter(
            torch.tensor(self.f_init_pos_par(Sigma_pos_par_init), dtype=self.dtype, device=self.device), requires_grad=flg_train_Sigma_pos_par)
        self.Sigma_free_par = None if Sigma_free_par_init is None else torch.nn.Parameter(
            torch.tensor(Sigma_free_par_init, dtype=self.dtype, device=self.device), requires_grad=flg_train_Sigma_free_par)

    def get_phi(self, X):
        if self.flg_offset:
            return torch.cat([X[:, self.active_dims], torch.ones(X.shape[0], 1, dtype=self.dtype, device=X.device)], 1)
        return X[:, self.active_dims]

    def get_Sigma(self):
        pos = None if self.Sigma_pos_par is None else self.f_transofrm_pos_par(self.Sigma_pos_par)
        return self.Sigma_function(pos, self.Sigma_free_par, *self.Sigma_f_additional_par_list)

    def get_Sigma_list(self):
        return [self.get_Sigma()]

    def _factor_weights(self, Sigma, D):
        """One factor's weight row [D + 1] (per input dimension, then the offset) from its Sigma matrix; diagonal Sigma only."""
        off = Sigma - torch.diag(torch.diagonal(Sigma))
        if float(off.detach().abs().max()) != 0.0:
            raise NotImplementedError("the CUDA kernels support diagonal Sigma matrices only (diagonal_covariance)")
        d = torch.diagonal(Sigma)
        row = self._scatter(d[:self.num_features], D, extra=1)
        if self.flg_offset:
            row = row + torch.nn.functional.one_hot(torch.tensor(D, device=d.device), D + 1).to(d.dtype) * d[self.num_features]
        return row

    def _kernel_terms(self, D):
        if not self.flg_no_mean:
            raise NotImplementedError("a linear prior mean phi(X) w is not supported on the CUDA path (all configurations use flg_no_mean)")
        return {"inv_ls": None, "lam": None, "mean": None, "polys": [self._factor_weights(self.get_Sigma(), D).reshape(1, -1)]}


class MPK_GP(Linear_GP):
    """Multiplicative polynomial kernel of degree poly_deg (reference :559-668)."""

    def __init__(self, active_dims, poly_deg, sigma_n_init=None, flg_train_sigma_n=True, Sigma_pos_par_init=None,
                 flg_train_Sigma_pos_par=True, flg_offset=True, name="", dtype=torch.float64, sigma_n_num=None, device=None):
        n_par = np.asarray(active_dims).size + (1 if flg_offset else 0)
        super().__init__(active_dims=active_dims, mean_init=None, flg_mean_trainable=False, flg_no_mean=True, sigma_n_init=sigma_n_init,
                         flg_train_sigma_n=flg_train_sigma_n, Sigma_function=_PCF.diagonal_covariance,
                         Sigma_f_additional_par_list=[n_par, True], Sigma_pos_par_init=None, flg_train_Sigma_pos_par=False,
                         Sigma_free_par_init=None, flg_train_Sigma_free_par=False, flg_offset=flg_offset, name=name, dtype=dtype,
                         sigma_n_num=sigma_n_num, device=device)
        self.poly_deg = poly_deg
        Sigma_pos_par_init = np.asarray(Sigma_pos_par_init, dtype=np.float64)
        self.Sigma_pos_par = torch.nn.Parameter(torch.tensor(np.log(Sigma_pos_par_init), dtype=self.dtype, device=self.device),
                                                requires_grad=flg_train_Sigma_pos_par)
        self.num_Sigma_pos_par = int(Sigma_pos_par_init.size / poly_deg)
        self.current_deg = 0

    def get_Sigma_deg(self, current_deg):
        n = self.num_Sigma_pos_par
        pos = (self.poly_deg - current_deg) * torch.exp(self.Sigma_pos_par[current_deg * n:(current_deg + 1) * n])
        return self.Sigma_function(pos, None, *self.Sigma_f_additional_par_list)

    def get_Sigma(self):
        return self.get_Sigma_deg(self.current_deg)

    def _kernel_terms(self, D):
        rows = [self._factor_weights(self.get_Sigma_deg(d), D) for d in range(self.poly_deg)]
        return {"inv_ls": None, "lam": None, "mean": None, "polys": [torch.stack(rows)]}


def get_Volterra_MPK_GP(active_dims, poly_deg, sigma_n_init=None, flg_train_sigma_n=True, Sigma_pos_par_init_list=[],
                        flg_train_Sigma_pos_par_list=[], name="", dtype=torch.float64, sigma_n_num=None, device=None):
    """Sum of MPK terms of degree 1..poly_deg; the degree-1 term carries the offset and the noise (reference :671-737)."""
    terms = [MPK_GP(active_dims, poly_deg=1, sigma_n_init=sigma_n_init, flg_train_sigma_n=flg_train_sigma_n,
                    Sigma_pos_par_init=Sigma_pos_par_init_list[0], flg_train_Sigma_pos_par=flg_train_Sigma_pos_par_list[0],
                    flg_offset=True, name="MPK_1", dtype=dtype, sigma_n_num=sigma_n_num, device=device)]
    for deg in range(1, poly_deg):
        terms.append(MPK_GP(active_dims, poly_deg=deg + 1, sigma_n_init=None, flg_train_sigma_n=False,
                            Sigma_pos_par_init=Sigma_pos_par_init_list[deg], flg_train_Sigma_pos_par=flg_train_Sigma_pos_par_list[deg],
                            flg_offset=False, name="MPK_" + str(deg + 1), dtype=dtype, sigma_n_num=None, device=device))
    return GP_prior.Sum_Independent_GP(*terms)


def get_SOR_GP(exact_GP_object):
    raise NotImplementedError("subset-of-regressors GPs are not used by any MC-PILCO configuration (and SOR_forward is broken in the "
                              "reference, Sparse_GP.py:226); outside the hot path")
