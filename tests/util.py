"""Shared helpers for the parity tests."""
import torch

from taichi_gaussian_rasterizer_b200 import RasterConfig
from taichi_gaussian_rasterizer_b200.misc.renderer2d import project_gaussians2d
from taichi_gaussian_rasterizer_b200.synthetic import random_2d_gaussians, random_3d_gaussians, random_camera

# tolerances stated by BASELINE.json north_star
IMAGE_REL_L2 = 1e-5   # images, depths, features (fp32)
GRAD_REL_L2 = 1e-4    # gradients (fp32)


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
  a, b = a.detach().double().cpu(), b.detach().double().cpu()
  denom = b.norm().item()
  if denom == 0:
    return (a - b).norm().item()
  return (a - b).norm().item() / denom


def scene2d(seed, n, image_size, channels=3, scale_factor=1.0, alpha_range=(0.1, 0.9), dtype=torch.float32):
  torch.manual_seed(seed)
  g = random_2d_gaussians(n, image_size, num_channels=channels, scale_factor=scale_factor, alpha_range=alpha_range)
  packed = project_gaussians2d(g).to(dtype)
  depth = g.z_depth.clamp(0, 1).to(torch.float32)
  return packed.contiguous(), depth.contiguous(), g.feature.to(dtype).contiguous()


def scene3d(seed, n, image_size=None, scale_factor=1.0, margin=0.3, sh_degree=None, channels=3):
  torch.manual_seed(seed)
  camera = random_camera(image_size=image_size)
  gaussians = random_3d_gaussians(n, camera, scale_factor=scale_factor, margin=margin, sh_degree=sh_degree,
                                  num_channels=channels)
  return gaussians, camera
