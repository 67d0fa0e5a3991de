"""Seeded synthetic scenes for tests and benchmarks.

Same recipes as taichi_splatting/tests/random_data.py (random_camera :15-44, random_3d_gaussians
:50-77, random_2d_gaussians :80-105), generated on the CPU from the torch global RNG (call
``torch.manual_seed`` first) and moved to the device by the caller.
"""
import math
from typing import Optional, Tuple

import torch
import torch.nn.functional as F

from .data_types import Gaussians2D, Gaussians3D
from .perspective import CameraParams
from .torch_lib import projection as torch_proj


def random_camera(pos_scale: float = 1., image_size: Optional[Tuple[int, int]] = None,
                  image_size_range=(256, 1024), near_plane=0.1) -> CameraParams:
  assert near_plane > 0
  q = F.normalize(torch.randn((1, 4)))
  t = torch.randn((3)) * pos_scale

  T_world_camera = torch_proj.join_rt(torch_proj.quat_to_mat(q)[0], t)
  T_camera_world = torch.inverse(T_world_camera)

  if image_size is None:
    lo, hi = image_size_range
    image_size = [int(x) for x in torch.randint(size=(2,), low=lo, high=hi)]

  w, h = image_size
  cx, cy = torch.tensor([w / 2, h / 2]) + torch.randn(2) * (w / 20)

  fov = torch.deg2rad(torch.rand(1) * 70 + 30)
  fx = w / (2 * torch.tan(fov / 2))
  fy = h / (2 * torch.tan(fov / 2))
  projection = torch.tensor([fx.item(), fy.item(), cx.item(), cy.item()], dtype=torch.float32)

  return CameraParams(T_camera_world=T_camera_world, projection=projection, image_size=(int(w), int(h)),
                      near_plane=near_plane, far_plane=near_plane * 1000.)


def random_3d_gaussians(n, camera_params: CameraParams, scale_factor: float = 1.0, alpha_range=(0.1, 0.9),
                        margin=0.0, sh_degree: Optional[int] = None, num_channels: int = 3) -> Gaussians3D:
  w, h = camera_params.image_size
  uv_pos = (torch.rand(n, 2) * (1 + margin) - margin * 0.5) * torch.tensor([w, h], dtype=torch.float32).unsqueeze(0)
  depth = torch_proj.inverse_ndc_depth(torch.rand(n), camera_params.near_plane, camera_params.far_plane)
  position = torch_proj.unproject_points(uv_pos, depth.unsqueeze(1), camera_params.T_image_world)
  fx = camera_params.T_image_camera[0, 0]

  scale = (w / math.sqrt(n)) * (depth / fx) * scale_factor
  scaling = (torch.rand(n, 3) + 0.2) * scale.unsqueeze(1)

  rotation = F.normalize(torch.randn(n, 4), dim=1)
  low, high = alpha_range
  alpha = torch.rand(n) * (high - low) + low

  if sh_degree is None:
    feature = torch.rand(n, num_channels)
  else:  # DC term carries the colour, higher bands a small view dependent perturbation
    d = (sh_degree + 1) ** 2
    feature = torch.randn(n, num_channels, d) * 0.1
    feature[:, :, 0] = (torch.rand(n, num_channels) - 0.5) / 0.282094791773878

  return Gaussians3D(position=position, log_scaling=torch.log(scaling), rotation=rotation,
                     alpha_logit=torch_proj.inverse_sigmoid(alpha).unsqueeze(1), feature=feature,
                     batch_size=(n,))


def random_2d_gaussians(n, image_size: Tuple[int, int], num_channels=3, scale_factor=1.0,
                        alpha_range=(0.1, 0.9), depth_range=(0.0, 1.0)) -> Gaussians2D:
  w, h = image_size
  position = torch.rand(n, 2) * torch.tensor([w, h], dtype=torch.float32).unsqueeze(0)
  depth = torch.rand((n, 1)) * (depth_range[1] - depth_range[0]) + depth_range[0]

  density_scale = scale_factor * w / (1 + math.sqrt(n))
  scaling = (torch.rand(n, 2) + 0.2) * density_scale

  rotation = torch.randn(n, 2)
  rotation = rotation / torch.norm(rotation, dim=1, keepdim=True)

  low, high = alpha_range
  alpha = torch.rand(n) * (high - low) + low

  return Gaussians2D(position=position, z_depth=depth, log_scaling=torch.log(scaling), rotation=rotation,
                     alpha_logit=torch_proj.inverse_sigmoid(alpha), feature=torch.rand(n, num_channels),
                     batch_size=(n,))


# ----------------------------------------------------------------------------------------------- BASELINE.json scenes
# The five configurations of BASELINE.json plus the workload its metric is quoted on ("bench"), as seeded synthetic
# scenes: one definition shared by bench.py, benchmarks/configs.py and the full-size parity tests.
BASELINE_SCENES = {
  # metric workload: 3 M gaussians, SH degree 3, 2048x1365
  "bench": dict(n=3_000_000, image_size=(2048, 1365), sh_degree=3, scale_factor=1.5),
  # config 2: render_gaussians 3D, 1 M gaussians, SH degree 3, 1920x1080
  "c2": dict(n=1_000_000, image_size=(1920, 1080), sh_degree=3),
  # config 3: bicycle-scale, 6 M gaussians, Mip-NeRF360-like scale / opacity distribution, visibility + split/prune stats
  "c3": dict(n=6_000_000, image_size=(2048, 1365), sh_degree=3, bimodal=True, stats=True),
  # config 4: feature lifting, 2 M gaussians, 32 feature channels + depth / depth variance at 3840x2160
  "c4": dict(n=2_000_000, image_size=(3840, 2160), channels=32, render_depth=True),
  # config 5: one view of the batched multi-view step, 3 M gaussians at 1600x1064
  "c5": dict(n=3_000_000, image_size=(1600, 1064), sh_degree=3, scale_factor=1.5),
}


def baseline_scene(name: str, seed: int = 0, n: Optional[int] = None, image_size: Optional[Tuple[int, int]] = None):
  """(gaussians, camera, spec) of a BASELINE.json configuration on the CPU; ``n`` / ``image_size`` override the size
  (the per gaussian scale follows n as in the recipe, so the screen coverage stays comparable).
  ``bimodal`` = log-normal scales (sigma_ln = 1) and opacity mass near 0.05 and 0.95."""
  spec = dict(BASELINE_SCENES[name])
  if n is not None:
    spec["n"] = n
  if image_size is not None:
    spec["image_size"] = tuple(image_size)
  torch.manual_seed(seed)
  cam = random_camera(image_size=spec["image_size"])
  g = random_3d_gaussians(spec["n"], cam, scale_factor=spec.get("scale_factor", 1.0), sh_degree=spec.get("sh_degree"),
                          num_channels=spec.get("channels", 3))
  if spec.get("bimodal"):
    g.log_scaling = g.log_scaling + torch.randn(spec["n"], 1)
    u = torch.rand(spec["n"])
    alpha = torch.where(u < 0.5, 0.02 + 0.08 * torch.rand(spec["n"]), 0.9 + 0.09 * torch.rand(spec["n"]))
    g.alpha_logit = torch.logit(alpha).unsqueeze(1)
  return g, cam, spec
