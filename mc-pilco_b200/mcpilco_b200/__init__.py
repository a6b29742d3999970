"""mcpilco_b200 — B200 (sm_100a) implementation of MC-PILCO's Monte-Carlo GP particle-rollout hot path.

Layout: `_native` (ctypes binding of the C ABI in include/mcpilco_b200.h), `_pack` (host-side flattening of
kernels / models / policies / costs into the ABI structs), `_ops` (torch-facing operators), `torch_ops` (the GP operators as `torch.ops.mcpilco.*` custom ops), and the mirror of the
reference's class API for this path: `gpr_lib.GP_prior`, `model_learning.Model_learning`,
`policy_learning.{Policy, Cost_function, MC_PILCO}`.  There is no CPU fallback anywhere on the path.
"""
__version__ = "0.1.0"
