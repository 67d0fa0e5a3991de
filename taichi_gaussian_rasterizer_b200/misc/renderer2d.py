"""2D gaussian helpers: the torch "projection" of Gaussians2D to packed records and the 2D renderer.

Follows taichi_splatting/misc/renderer2d.py:17-33 (project_gaussians2d), :36-58 (basis helpers), :60-132 (the split
operations densification uses: ``split_gaussians2d`` samples the children from the parent, ``uniform_split_gaussians2d``
places them evenly along one principal axis) and :135-149 (render_gaussians).
"""
import math
from dataclasses import replace
from numbers import Integral
from typing import Optional

import torch
import torch.nn.functional as F
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


def point_rotation(points: Gaussians2D):
  v1 = points.rotation / torch.norm(points.rotation, dim=1, keepdim=True)
  v2 = torch.stack([-v1[..., 1], v1[..., 0]], dim=-1)
  return torch.stack([v1, v2], dim=1)


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


# ---------------------------------------------------------------------------------------------- splitting
def repeat_sample_gaussians(samples: torch.Tensor, points: Gaussians2D, n: int = 2) -> torch.Tensor:
  """Offsets (N, n, 2) in image space of ``samples`` (N, n, 2) given in each parent's own basis (unit = 1 sigma)."""
  basis = point_basis(points)                                        # (N, 2, 2), columns = sigma-scaled axes
  return torch.einsum('pij,pkj->pki', basis, samples.reshape(-1, n, 2).to(basis.dtype))


def sample_gaussians(points: Gaussians2D) -> torch.Tensor:
  """One offset per gaussian drawn from the gaussian itself."""
  return repeat_sample_gaussians(torch.randn_like(points.position).unsqueeze(1), points, n=1).squeeze(1)


def split_with_offsets(points: Gaussians2D, offsets: torch.Tensor, depth_noise: float = 1e-2) -> Gaussians2D:
  """``n`` children per parent: the parent's fields repeated, positions moved by ``offsets`` (N, n, 2), sort depth
  jittered by ``depth_noise`` so that siblings do not tie (renderer2d.py:60-70)."""
  num_points, n, _ = offsets.shape
  children = points.apply(lambda t: t.repeat_interleave(n, dim=0), batch_size=[num_points * n])
  depth = children.z_depth + torch.randn_like(children.z_depth) * depth_noise
  return replace(children, position=children.position + offsets.reshape(-1, 2),
                 z_depth=depth.clamp_min(1e-6), batch_size=(num_points * n,))


def split_gaussians2d(points: Gaussians2D, n: int = 2, scaling: Optional[float] = None,
                      depth_noise: float = 1e-2) -> Gaussians2D:
  """The splitting operation of gaussian-splatting densification: ``n`` children sampled from the parent (0.5 sigma),
  each scaled by ``scaling`` (default 1 / sqrt(n)) (renderer2d.py:73-100)."""
  samples = 0.5 * torch.randn((points.batch_size[0], n, 2), device=points.position.device)
  offsets = repeat_sample_gaussians(samples, points, n)
  factor = math.log(1.0 / math.sqrt(n) if scaling is None else scaling)
  shrunk = replace(points, log_scaling=points.log_scaling + factor, batch_size=points.batch_size)
  return split_with_offsets(shrunk, offsets, depth_noise)


def uniform_split_gaussians2d(points: Gaussians2D, n: int = 2, scaling: Optional[float] = None,
                              depth_noise: float = 1e-2, sep: float = 0.7, random_axis: bool = False,
                              eps: float = 1e-6) -> Gaussians2D:
  """``n`` children evenly spaced in [-sep, sep] sigma along ONE principal axis — the longer one, or one drawn with
  probability proportional to its scale when ``random_axis`` — and shrunk along that axis only by ``scaling``
  (default sqrt(n) / n) (renderer2d.py:113-132)."""
  if random_axis:
    probs = F.normalize(points.scaling + eps, p=1, dim=1)
    axis = torch.multinomial(probs, num_samples=1).squeeze(1)
  else:
    axis = torch.argmax(points.log_scaling, dim=1)
  along = F.one_hot(axis, num_classes=2)                              # (N, 2) 1 on the chosen axis
  steps = torch.linspace(-sep, sep, n, device=points.position.device)
  samples = steps.view(1, n, 1) * along.view(-1, 1, 2)                # (N, n, 2) in the parent's basis
  offsets = repeat_sample_gaussians(samples, points, n)
  shrink = math.sqrt(n) / n if scaling is None else scaling
  shrunk = points.set_scaling(points.scaling * (along * shrink + (1 - along)))
  return split_with_offsets(shrunk, offsets, depth_noise)
