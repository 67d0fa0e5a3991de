"""Torch helpers of the render path that are not kernels.

``ndc_depth`` / ``inverse_ndc_depth`` follow taichi_splatting/torch_lib/projection.py:120-129
(plain eager torch here; the reference wraps them in ``torch.compile``), the point
(un)projection helpers follow :47-57 and are used by the synthetic scene generators.
"""
import torch


def ndc_depth(depth: torch.Tensor, near: float, far: float) -> torch.Tensor:
  # 0 at the near plane, 1 at the far plane
  return 1 - (1. / depth - 1. / far) / (1. / near - 1. / far)


def inverse_ndc_depth(ndc: torch.Tensor, near: float, far: float) -> torch.Tensor:
  return 1.0 / ((1.0 - ndc) * (1 / near - 1 / far) + 1 / far)


def inverse_sigmoid(x: torch.Tensor):
  return torch.log(x / (1 - x))


def quat_to_mat(quat: torch.Tensor) -> torch.Tensor:
  """Rotation matrix of a unit quaternion in (x, y, z, w) order."""
  x, y, z, w = quat.unbind(-1)
  x2, y2, z2 = x * x, y * y, z * z
  rows = [1 - 2 * y2 - 2 * z2, 2 * x * y - 2 * w * z, 2 * x * z + 2 * w * y,
          2 * x * y + 2 * w * z, 1 - 2 * x2 - 2 * z2, 2 * y * z - 2 * w * x,
          2 * x * z - 2 * w * y, 2 * y * z + 2 * w * x, 1 - 2 * x2 - 2 * y2]
  return torch.stack(rows, dim=-1).reshape(quat.shape[:-1] + (3, 3))


def join_rt(r, t):
  T = torch.eye(4, device=r.device, dtype=r.dtype)
  T[0:3, 0:3] = r
  T[0:3, 3] = t
  return T


def make_homog(points):
  ones = torch.ones(points.shape[:-1] + (1,), dtype=points.dtype, device=points.device)
  return torch.cat([points, ones], dim=-1)


def transform44(transform, points):
  return (transform.reshape(1, 4, 4) @ points.reshape(-1, 4, 1))[..., 0]


def project_points(transform, xyz):
  homog = transform44(transform, make_homog(xyz))
  depth = homog[..., 2:3]
  return homog[..., 0:2] / depth, depth


def unproject_points(uv, depth, transform):
  points = torch.cat([uv * depth, depth, torch.ones_like(depth)], dim=-1)
  transformed = transform44(torch.inverse(transform), points)
  return transformed[..., 0:3] / transformed[..., 3:4]
