"""Training criteria — the reference's `gpr_lib/Likelihood/Gaussian_likelihood.py` surface.

Marginal_log_likelihood (:12-24): 0.5 ((Y - m_X)^T K_X^-1 (Y - m_X) + log det K_X) — the N log 2 pi constant is dropped, as
in the reference.  Called on the handle `GP_prior.forward` returns in training mode, the value and its hyper-parameter
gradients come from ONE native call (blocked Cholesky + analytic gradient); no autograd graph through a factorisation."""
import torch

from ..GP_prior.GP_prior import PriorOutput


class Marginal_log_likelihood(torch.nn.modules.loss._Loss):
    def forward(self, output_GP_prior, Y):
        if isinstance(output_GP_prior, PriorOutput):
            return output_GP_prior.gp.nlml(output_GP_prior.X, Y)
        m_X, _, K_X_inv, log_det = output_GP_prior  # plain values (no training graph): torch ops on the device
        r = Y - m_X
        return 0.5 * (torch.matmul(r.transpose(0, 1), torch.matmul(K_X_inv, r)) + log_det)


class Posterior_log_likelihood(torch.nn.modules.loss._Loss):
    """sum_i (Y_i - Yhat_i)^2 / (2 var_i) + 0.5 log var_i (:27-37)."""

    def forward(self, Y, Y_hat, var):
        d = Y - Y_hat
        return torch.sum(d ** 2 / (2 * var) + 0.5 * torch.log(var))
