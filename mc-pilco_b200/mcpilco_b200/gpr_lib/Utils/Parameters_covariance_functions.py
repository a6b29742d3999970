"""Sigma parametrisations of the linear kernels — the part of the reference's
`gpr_lib/Utils/Parameters_covariance_functions.py` the hot path uses (diagonal_covariance :18-27, _ARD :30-32).
These build the (tiny) weight matrices on the host side; the native kernels take their diagonals."""
import torch


def diagonal_covariance(pos_par=None, free_par=None, num_par=None, flg_ARD=False):
    """diag(pos_par**2) with ARD, pos_par**2 * I otherwise."""
    if flg_ARD:
        if num_par != pos_par.size()[0]:
            raise RuntimeError("The number of positive parameters and num_par must be equal when flg_ARD=True")
        return torch.diag(pos_par ** 2)
    return pos_par ** 2 * torch.eye(num_par, dtype=pos_par.dtype, device=pos_par.device)


def diagonal_covariance_ARD(pos_par=None, free_par=None):
    return torch.diag(pos_par ** 2)
