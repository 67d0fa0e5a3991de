"""B200-native drop-in for the differentiable render path of taichi_splatting
(uc-vision/taichi_gaussian_rasterizer): same operator names, arguments and autograd behaviour,
hand-written sm_100a CUDA underneath (libgsplat_b200.so, C ABI in include/gsplat_b200.h).

    import taichi_gaussian_rasterizer_b200 as taichi_splatting

The names below are the ones taichi_splatting/__init__.py:1-33 exports, plus the helpers this package adds
(render_projected, viewspace_gradient, CameraParams, cuda_lib, set_raster_options, taichi_queue).
"""
__version__ = "0.1.0"

# containers and configuration
from .data_types import Gaussians2D, Gaussians3D, RasterConfig
from .perspective import CameraParams
from . import perspective
# operators, in pipeline order
from .spherical_harmonics import evaluate_sh_at, evaluate_sh_views
from .mapper.tile_mapper import map_to_tiles, pad_to_tile
from . import cuda_lib
from .rasterizer import rasterize, rasterize_with_tiles, set_raster_options
# composition
from .rendering import Rendering
from .renderer import render_gaussians, render_projected, viewspace_gradient
from .graphs import CapturedStep, overlap_capacity_for
# launch-queue compatibility shim (the Taichi runtime it guarded does not exist here)
from .taichi_queue import TaichiQueue, taichi_queue

__all__ = [
  "Gaussians2D", "Gaussians3D", "RasterConfig", "CameraParams", "perspective",
  "evaluate_sh_at", "evaluate_sh_views", "map_to_tiles", "pad_to_tile", "cuda_lib", "rasterize", "rasterize_with_tiles",
  "set_raster_options", "Rendering", "render_gaussians", "render_projected", "viewspace_gradient",
  "CapturedStep", "overlap_capacity_for",
  "TaichiQueue", "taichi_queue",
]
