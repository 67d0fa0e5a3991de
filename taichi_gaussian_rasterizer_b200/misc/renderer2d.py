"""2D gaussian helpers: the torch "projection" of Gaussians2D to packed records and the 2D renderer.

Follows taichi_splatting/misc/renderer2d.py:17-33 (project_gaussians2d), :36-58 (basis helpers) and
:135-149 (render_gaussians).  The split helpers of that file are training policy and out of scope.
"""
from numbers import Integral

import torch
from beartype import beartype
from beartype.typing import Tuple

from ..data_types import Gaussians2D, RasterConfig
from ..rasterizer import rasterize


@beartype
def project_gaussians2d(points: Gaussians2D) -> torch.Tensor:
  """Packed (N, 7) records [mean(2), axis(2), sigma(2), alpha] of a Gaussians2D (differentiable torch)."""
  alpha = torch.sigmoid(points.alpha_logit)
  sigma = points.scaling
  v1 = points.rotation / torch.norm(points.rotation, dim=1, keepdim=True)
  return torch.cat([points.position, v1, sigma, alpha.reshape(-1, 1)], dim=-1)


def point_basis(points: Gaussians2D, eps: float = 1e-4):
  scale = torch.clamp_min(points.scaling, eps)
  v1 = points.rotation / torch.norm(points.rotation, dim=1, keepdim=True)
  v2 = torch.stack([-v1[..., 1], v1[..., 0]], dim=-1)
  return torch.stack([v1, v2], dim=2) * scale.unsqueeze(-2)


def point_covariance(gaussians):
  basis = point_basis(gaussians)
  return torch.bmm(basis, basis.transpose(1, 2))


def render_gaussians(gaussians: Gaussians2D, image_size: Tuple[Integral, Integral],
                     raster_config: RasterConfig = RasterConfig()):
  gaussians2d = project_gaussians2d(gaussians)
  return rasterize(gaussians2d=gaussians2d,
                   depth=torch.clamp(gaussians.z_depth, 0, 1),
                   features=gaussians.feature,
                   image_size=image_size,
                   config=raster_config)
