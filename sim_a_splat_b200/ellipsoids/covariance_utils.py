"""Covariance helpers with the reference's names and signatures
(/root/reference/sim_a_splat/ellipsoids/covariance_utils.py:141-157).

The reference converts quaternion -> angle-axis -> rotation matrix (ceres-style, with `theta + 1e-6` in a
divisor, :91); here the rotation comes straight from the normalised quaternion, which agrees with the
reference to ~1e-6 (tests/golden/compute_cov.npz, tolerance 2e-5 of the covariance scale).
"""
import torch


def quaternion_to_rotation_matrix(quat: torch.Tensor) -> torch.Tensor:
    q = quat / torch.linalg.norm(quat, dim=-1, keepdim=True)
    w, x, y, z = q.unbind(-1)
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                     2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                     2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], dim=-1)
    return R.reshape(*quat.shape[:-1], 3, 3)


def scaling_to_mat(scaling: torch.Tensor, exp: bool = False) -> torch.Tensor:
    return torch.diag_embed(torch.exp(scaling) if exp else scaling, 0)


def compute_cov(quat: torch.Tensor, scaling: torch.Tensor, exp: bool = False) -> torch.Tensor:
    R = quaternion_to_rotation_matrix(quat)
    M = R @ scaling_to_mat(scaling, exp=exp)
    return M @ M.transpose(-2, -1)
