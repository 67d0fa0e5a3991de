"""B200-native drop-in for the differentiable render path of taichi_splatting
(uc-vision/taichi_gaussian_rasterizer): same operator names, arguments and autograd behaviour,
hand-written sm_100a CUDA underneath (libgsplat_b200.so, C ABI in include/gsplat_b200.h).

    import taichi_gaussian_rasterizer_b200 as taichi_splatting

Exports mirror taichi_splatting/__init__.py:1-33.
"""
from .renderer import render_gaussians, render_projected, Rendering, viewspace_gradient
from .data_types import Gaussians2D, Gaussians3D, RasterConfig
from .mapper.tile_mapper import map_to_tiles, pad_to_tile
from .rasterizer import rasterize, rasterize_with_tiles, set_raster_options

from .spherical_harmonics import evaluate_sh_at

from . import perspective
from . import cuda_lib
from .perspective import CameraParams
from .taichi_queue import TaichiQueue, taichi_queue

__version__ = "0.1.0"

__all__ = [
  'render_gaussians',
  'render_projected',
  'Rendering',
  'viewspace_gradient',

  'map_to_tiles',
  'pad_to_tile',

  'Gaussians2D',
  'Gaussians3D',

  'RasterConfig',
  'evaluate_sh_at',

  'rasterize',
  'rasterize_with_tiles',
  'set_raster_options',

  'perspective',
  'cuda_lib',
  'CameraParams',
  'TaichiQueue',
  'taichi_queue',
]
