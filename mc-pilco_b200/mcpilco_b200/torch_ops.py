"""`torch.ops.mcpilco.*` — the GP operators of the hot path registered as torch custom ops (torch.library), a thin layer over the C ABI.

Custom ops take tensors and scalars only, so a kernel specification travels as a CPU uint8 tensor holding the bytes of the POD
`McpGpSpec` (`spec_tensor(spec)` / `GP_prior.gp_spec(D)`).  CUDA float64 tensors only; there is no CPU implementation to dispatch to.

    torch.ops.mcpilco.gp_covariance(spec, X1, X2, add_noise) -> K
    torch.ops.mcpilco.gp_diag_covariance(spec, X) -> diag
    torch.ops.mcpilco.gp_precompute(spec, X, y) -> (alpha [N,1], Kinv [N,N])
    torch.ops.mcpilco.gp_predict(spec, Xtr, alpha, Kinv, Xs, var_scale) -> (mean [M,1], var [M,1])
    torch.ops.mcpilco.gp_predict_jac(spec, Xtr, alpha, Kinv, Xs, var_scale) -> (mean, var, dmean/dx [M,D], dvar/dx [M,D])
    torch.ops.mcpilco.gp_nlml(spec, X, y) -> packed value + gradient (layout: include/mcpilco_b200.h)

The particle rollout itself is exposed as one autograd node (policy_learning.MC_PILCO._ParticleRollout), not as a flat op: its
descriptor structs do not map onto tensor arguments.
"""
import ctypes as C
from typing import Optional, Tuple

import torch

from . import _native as N
from . import _ops as ops


def spec_tensor(spec):
    """CPU uint8 tensor with the bytes of a McpGpSpec."""
    return torch.frombuffer(bytearray(C.string_at(C.addressof(spec), C.sizeof(spec))), dtype=torch.uint8).clone()


def _spec(t):
    if t.device.type != "cpu" or t.dtype != torch.uint8 or t.numel() != C.sizeof(N.GpSpec):
        raise RuntimeError("mcpilco: spec must be a CPU uint8 tensor of %d bytes (torch_ops.spec_tensor)" % C.sizeof(N.GpSpec))
    return N.GpSpec.from_buffer_copy(bytes(t.contiguous().numpy().tobytes()))


@torch.library.custom_op("mcpilco::gp_covariance", mutates_args=(), device_types="cuda")
def gp_covariance(spec: torch.Tensor, X1: torch.Tensor, X2: Optional[torch.Tensor], add_noise: bool) -> torch.Tensor:
    return ops.gp_covariance(_spec(spec), X1, X2, add_noise)


@torch.library.custom_op("mcpilco::gp_diag_covariance", mutates_args=(), device_types="cuda")
def gp_diag_covariance(spec: torch.Tensor, X: torch.Tensor) -> torch.Tensor:
    return ops.gp_diag_covariance(_spec(spec), X)


@torch.library.custom_op("mcpilco::gp_precompute", mutates_args=(), device_types="cuda")
def gp_precompute(spec: torch.Tensor, X: torch.Tensor, y: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    alpha, Kinv = ops.gp_precompute(_spec(spec), X, y)
    return alpha, Kinv.contiguous()


@torch.library.custom_op("mcpilco::gp_predict", mutates_args=(), device_types="cuda")
def gp_predict(spec: torch.Tensor, Xtr: torch.Tensor, alpha: torch.Tensor, Kinv: torch.Tensor, Xs: torch.Tensor,
               var_scale: float) -> Tuple[torch.Tensor, torch.Tensor]:
    mean, var = ops.gp_predict([ops.FittedGp(_spec(spec), Xtr, alpha, Kinv, var_scale=var_scale)], Xs)
    return mean, var


@torch.library.custom_op("mcpilco::gp_predict_jac", mutates_args=(), device_types="cuda")
def gp_predict_jac(spec: torch.Tensor, Xtr: torch.Tensor, alpha: torch.Tensor, Kinv: torch.Tensor, Xs: torch.Tensor,
                   var_scale: float) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    mean, var, jm, jv = ops.gp_predict([ops.FittedGp(_spec(spec), Xtr, alpha, Kinv, var_scale=var_scale)], Xs, jac=True)
    return mean, var, jm[:, 0, :].contiguous(), jv[:, 0, :].contiguous()


@torch.library.custom_op("mcpilco::gp_nlml", mutates_args=(), device_types="cuda")
def gp_nlml(spec: torch.Tensor, X: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    return ops.gp_nlml(_spec(spec), X, y)


# shape functions so the ops compose with tracing / fake tensors
@gp_covariance.register_fake
def _(spec, X1, X2, add_noise):
    return X1.new_empty(X1.shape[0], X1.shape[0] if X2 is None else X2.shape[0])


@gp_diag_covariance.register_fake
def _(spec, X):
    return X.new_empty(X.shape[0])


@gp_precompute.register_fake
def _(spec, X, y):
    return X.new_empty(X.shape[0], 1), X.new_empty(X.shape[0], X.shape[0])


@gp_predict.register_fake
def _(spec, Xtr, alpha, Kinv, Xs, var_scale):
    return Xs.new_empty(Xs.shape[0], 1), Xs.new_empty(Xs.shape[0], 1)


@gp_predict_jac.register_fake
def _(spec, Xtr, alpha, Kinv, Xs, var_scale):
    return Xs.new_empty(Xs.shape[0], 1), Xs.new_empty(Xs.shape[0], 1), Xs.new_empty(Xs.shape), Xs.new_empty(Xs.shape)


@gp_nlml.register_fake
def _(spec, X, y):
    return X.new_empty(4 + N.MAX_D + N.MAX_POLY * N.MAX_DEG * (N.MAX_D + 1))
