"""CPU oracle of the render path — TEST INFRASTRUCTURE, never imported by the product package.

ctypes front end of ``oracle/liboracle.so`` (``oracle.cpp``, a C++/OpenMP restatement of the
reference algorithm, built by ``oracle/Makefile``) plus the pure-torch restatements in
``oracle/torch_ref.py``.  Function names and argument meanings follow the reference operators
so the parity tests read like the reference's own tests:

  project_to_image / projection_apply   perspective/projection.py:190-248
  evaluate_sh_at                         spherical_harmonics.py:167-178
  map_to_tiles (+ stage functions)       mapper/tile_mapper.py:202-223
  rasterize_with_tiles / rasterize       rasterizer/function.py:96-161   (CPU autograd.Function)

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import
this package.  All tensors are CPU tensors.
"""
import ctypes
import os
import subprocess
from pathlib import Path
from typing import NamedTuple, Optional, Tuple

import numpy as np
import torch

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "liboracle.so"


def build(force: bool = False) -> Path:
  src = _HERE / "oracle.cpp"
  hdr = _HERE.parent / "include" / "gs_numeric.h"
  stale = (not _LIB_PATH.exists()
           or _LIB_PATH.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime))
  if force or stale:
    subprocess.run(["make", "-C", str(_HERE), "-B", "liboracle.so"], check=True,
                   stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
  return _LIB_PATH


_lib = None


def lib():
  global _lib
  if _lib is None:
    if not _LIB_PATH.exists():
      build()
    _lib = ctypes.CDLL(str(_LIB_PATH))
    _lib.orc_expf.restype = ctypes.c_float
    _lib.orc_expf.argtypes = [ctypes.c_float]
    _lib.orc_logf.restype = ctypes.c_float
    _lib.orc_logf.argtypes = [ctypes.c_float]
    _lib.orc_num_threads.restype = ctypes.c_int
  return _lib


class _Cfg(ctypes.Structure):
  _fields_ = [("tile_size", ctypes.c_int32), ("antialias", ctypes.c_int32),
              ("use_alpha_blending", ctypes.c_int32), ("compute_visibility", ctypes.c_int32),
              ("compute_point_heuristic", ctypes.c_int32), ("stride_x", ctypes.c_int32),
              ("stride_y", ctypes.c_int32), ("pad_", ctypes.c_int32),
              ("clamp_max_alpha", ctypes.c_double), ("alpha_threshold", ctypes.c_double),
              ("saturate_threshold", ctypes.c_double)]


def _cfg(config) -> _Cfg:
  return _Cfg(config.tile_size, int(config.antialias), int(config.use_alpha_blending),
              int(config.compute_visibility), int(config.compute_point_heuristic),
              config.pixel_stride[0], config.pixel_stride[1], 0,
              config.clamp_max_alpha, config.alpha_threshold, config.saturate_threshold)


def _p(t: Optional[torch.Tensor]):
  if t is None:
    return ctypes.c_void_p(0)
  assert t.device.type == "cpu" and t.is_contiguous(), "oracle takes contiguous CPU tensors"
  return ctypes.c_void_p(t.data_ptr())


def _suffix(dtype):
  return {torch.float32: "f32", torch.float64: "f64"}[dtype]


def num_threads() -> int:
  return int(lib().orc_num_threads())


def set_num_threads(n: int):
  lib().orc_set_num_threads(ctypes.c_int(n))


def expf(x: float) -> float:
  return float(lib().orc_expf(x))


def logf(x: float) -> float:
  return float(lib().orc_logf(x))


# ----------------------------------------------------------------------------- projection
def projection_forward(position, log_scaling, rotation, alpha_logit, T_camera_world, projection,
                       image_size, depth_range, blur_cov=0.0, clamp_margin=0.15,
                       alpha_threshold=1. / 255.):
  """project_kernel + nonzero + gather (perspective/projection.py:31-80, :146-149).
  Returns (points (V,7), depth (V,1), indexes (V,) int64)."""
  dtype = position.dtype
  n = position.shape[0]
  args = [t.detach().contiguous() for t in (position, log_scaling, rotation, alpha_logit.reshape(-1),
                                            T_camera_world.to(dtype), projection.to(dtype))]
  points = torch.empty((n, 7), dtype=dtype)
  depth = torch.empty((n,), dtype=dtype)
  fn = getattr(lib(), f"orc_project_fwd_{_suffix(dtype)}")
  fn(ctypes.c_int64(n), *[_p(a) for a in args], ctypes.c_int(int(image_size[0])), ctypes.c_int(int(image_size[1])),
     ctypes.c_double(depth_range[0]), ctypes.c_double(depth_range[1]), ctypes.c_double(blur_cov),
     ctypes.c_double(clamp_margin), ctypes.c_double(alpha_threshold), _p(points), _p(depth))
  indexes = torch.nonzero(depth).squeeze(1)
  return points[indexes], depth[indexes].unsqueeze(1), indexes


def projection_backward(position, log_scaling, rotation, alpha_logit, T_camera_world, projection, image_size,
                        indexes, grad_points, grad_depth=None, blur_cov=0.0, clamp_margin=0.15):
  """Gradients of (points, depth) = project_to_image(...) w.r.t. the six inputs, in the inputs' dtype: the reverse
  sweep of project_one (oracle.cpp project_bwd; the reference differentiates indexed_project_kernel with Taichi
  autodiff, perspective/projection.py:83-118, :164-185).  Returns (grads dict, cond (V, 2)) where
  cond[:, 0] = sqrt(gap) / trace and cond[:, 1] = |n| / trace of the projected covariance: the two quantities the
  eigen decomposition divides by (small = ill conditioned in float32 for any implementation)."""
  dtype = position.dtype
  n, nv = position.shape[0], indexes.shape[0]
  args = [t.detach().contiguous() for t in (position, log_scaling, rotation, alpha_logit.reshape(-1),
                                            T_camera_world.to(dtype), projection.to(dtype))]
  out = dict(position=torch.zeros((n, 3), dtype=dtype), log_scaling=torch.zeros((n, 3), dtype=dtype),
             rotation=torch.zeros((n, 4), dtype=dtype), alpha_logit=torch.zeros((n, 1), dtype=dtype),
             T_camera_world=torch.zeros((4, 4), dtype=dtype), projection=torch.zeros((4,), dtype=dtype))
  cond = torch.zeros((nv, 2), dtype=dtype)
  gd = None if grad_depth is None else grad_depth.detach().to(dtype).reshape(-1).contiguous()
  fn = getattr(lib(), f"orc_project_bwd_{_suffix(dtype)}")
  fn(ctypes.c_int64(nv), _p(indexes.contiguous()), *[_p(a) for a in args], ctypes.c_int(int(image_size[0])),
     ctypes.c_int(int(image_size[1])), ctypes.c_double(blur_cov), ctypes.c_double(clamp_margin),
     _p(grad_points.detach().to(dtype).contiguous()), _p(gd), *[_p(out[k]) for k in
     ("position", "log_scaling", "rotation", "alpha_logit", "T_camera_world", "projection")], _p(cond))
  return out, cond


def project_to_image(gaussians, camera_params, config):
  return projection_forward(*gaussians.shape_tensors(), camera_params.T_camera_world, camera_params.projection,
                            camera_params.image_size, camera_params.depth_range, config.blur_cov,
                            config.clamp_margin, config.alpha_threshold)


# ----------------------------------------------------------------------------- spherical harmonics
def evaluate_sh_at(sh_params, positions, indexes, camera_pos):
  dtype = sh_params.dtype
  M, K, D = sh_params.shape
  nv = indexes.shape[0]
  out = torch.empty((nv, K), dtype=dtype)
  fn = getattr(lib(), f"orc_sh_fwd_{_suffix(dtype)}")
  fn(ctypes.c_int64(nv), ctypes.c_int(K), ctypes.c_int(D), _p(sh_params.detach().contiguous()),
     _p(positions.detach().contiguous()), _p(indexes.contiguous()), _p(camera_pos.detach().contiguous()), _p(out))
  return out


# ----------------------------------------------------------------------------- tile mapper
def pad_to_tile(image_size, tile_size):
  return tuple(int(-(-int(x) // tile_size) * tile_size) for x in image_size)


def tile_counts(gaussians, image_size, config):
  padded = pad_to_tile(image_size, config.tile_size)
  n = gaussians.shape[0]
  counts = torch.zeros((n,), dtype=torch.int32)
  lib().orc_tile_counts(ctypes.c_int64(n), _p(gaussians.contiguous()), ctypes.c_int(padded[0]),
                        ctypes.c_int(padded[1]), ctypes.c_int(config.tile_size),
                        ctypes.c_float(config.alpha_threshold), _p(counts))
  return counts


def full_cumsum(x: torch.Tensor) -> Tuple[torch.Tensor, int]:
  """cuda_lib/full_cumsum.cu:16-47: exclusive scan with the total appended."""
  out = torch.zeros((x.shape[0] + 1,), dtype=x.dtype)
  out[1:] = torch.cumsum(x, 0)
  return out, int(out[-1])


def tile_emit_keys(gaussians, depth, cum, total, image_size, config, use_depth16=False):
  padded = pad_to_tile(image_size, config.tile_size)
  n = gaussians.shape[0]
  keys = torch.zeros((total,), dtype=torch.int64)
  values = torch.zeros((total,), dtype=torch.int32)
  lib().orc_tile_emit_keys(ctypes.c_int64(n), _p(gaussians.contiguous()), _p(depth.reshape(-1).contiguous()),
                           _p(cum.contiguous()), ctypes.c_int(padded[0]), ctypes.c_int(padded[1]),
                           ctypes.c_int(config.tile_size), ctypes.c_float(config.alpha_threshold),
                           ctypes.c_int(int(use_depth16)), _p(keys), _p(values))
  return keys, values


def radix_sort_pairs(keys: torch.Tensor, values: torch.Tensor, start_bit=0, end_bit=None):
  """keys: int64 tensor holding the uint64 bit pattern."""
  if end_bit is None or end_bit < 0:
    end_bit = 64
  n = keys.shape[0]
  ko, vo = torch.empty_like(keys), torch.empty_like(values)
  lib().orc_radix_sort_pairs(ctypes.c_int64(n), _p(keys.contiguous()), _p(values.contiguous()),
                             ctypes.c_int(start_bit), ctypes.c_int(end_bit), _p(ko), _p(vo))
  return ko, vo


def find_ranges(sorted_keys, num_tiles, use_depth16=False):
  ranges = torch.zeros((num_tiles, 2), dtype=torch.int32)
  lib().orc_find_ranges(ctypes.c_int64(sorted_keys.shape[0]), _p(sorted_keys.contiguous()),
                        ctypes.c_int(int(use_depth16)), _p(ranges))
  return ranges


def map_to_tiles(gaussians, depth, image_size, config, use_depth16=False):
  """mapper/tile_mapper.py:168-196.  Returns (overlap_to_point (K,) int32, tile_ranges (TH,TW,2) int32)."""
  assert gaussians.dtype == torch.float32 and depth.dtype == torch.float32, "tile mapper is f32 only"
  padded = pad_to_tile(image_size, config.tile_size)
  tile_shape = (padded[1] // config.tile_size, padded[0] // config.tile_size)
  assert tile_shape[0] * tile_shape[1] < 65535
  counts = tile_counts(gaussians, image_size, config)
  cum, total = full_cumsum(counts)
  if total == 0:
    return torch.empty((0,), dtype=torch.int32), torch.zeros((*tile_shape, 2), dtype=torch.int32)
  keys, values = tile_emit_keys(gaussians, depth, cum[:-1], total, image_size, config, use_depth16)
  keys, values = radix_sort_pairs(keys, values, 0, 32 if use_depth16 else 48)
  ranges = find_ranges(keys, tile_shape[0] * tile_shape[1], use_depth16)
  return values, ranges.view(*tile_shape, 2)


# ----------------------------------------------------------------------------- rasterizer
class RasterOut(NamedTuple):
  image: torch.Tensor
  image_weight: torch.Tensor
  point_heuristic: torch.Tensor
  visibility: torch.Tensor


def raster_forward(gaussians2d, features, overlap_to_point, tile_overlap_ranges, image_size, config,
                   emulate_stale_tail=True):
  dtype = gaussians2d.dtype
  w, h = int(image_size[0]), int(image_size[1])
  V, F = features.shape
  image = torch.zeros((h, w, F), dtype=dtype)
  alpha = torch.zeros((h, w), dtype=dtype)
  vis = torch.zeros((V,), dtype=dtype) if config.compute_visibility else torch.empty((0,), dtype=dtype)
  cfg = _cfg(config)
  fn = getattr(lib(), f"orc_raster_fwd_{_suffix(dtype)}")
  fn(ctypes.byref(cfg), ctypes.c_int64(V), ctypes.c_int(F), _p(gaussians2d.detach().contiguous()),
     _p(features.detach().contiguous()), _p(tile_overlap_ranges.contiguous()), _p(overlap_to_point.contiguous()),
     ctypes.c_int(w), ctypes.c_int(h), ctypes.c_int(int(emulate_stale_tail)), _p(image), _p(alpha),
     _p(vis) if config.compute_visibility else ctypes.c_void_p(0))
  return image, alpha, vis


def raster_backward(gaussians2d, features, overlap_to_point, tile_overlap_ranges, image_size, config,
                    image, grad_image, points_requires_grad=True, features_requires_grad=True, heuristic=None):
  dtype = gaussians2d.dtype
  w, h = int(image_size[0]), int(image_size[1])
  V, F = features.shape
  gp = torch.zeros((V, 7), dtype=dtype)
  gf = torch.zeros((V, F), dtype=dtype)
  cfg = _cfg(config)
  fn = getattr(lib(), f"orc_raster_bwd_{_suffix(dtype)}")
  fn(ctypes.byref(cfg), ctypes.c_int64(V), ctypes.c_int(F), _p(gaussians2d.detach().contiguous()),
     _p(features.detach().contiguous()), _p(tile_overlap_ranges.contiguous()), _p(overlap_to_point.contiguous()),
     ctypes.c_int(w), ctypes.c_int(h), _p(image.contiguous()), _p(grad_image.contiguous()),
     ctypes.c_int(int(points_requires_grad)), ctypes.c_int(int(features_requires_grad)), _p(gp), _p(gf),
     _p(heuristic) if (config.compute_point_heuristic and heuristic is not None) else ctypes.c_void_p(0))
  return gp, gf


class _Rasterize(torch.autograd.Function):
  """CPU autograd wrapper with the conventions of rasterizer/function.py:41-91."""

  @staticmethod
  def forward(ctx, gaussians2d, features, overlap_to_point, tile_overlap_ranges, image_size, config, emulate):
    image, alpha, vis = raster_forward(gaussians2d, features, overlap_to_point, tile_overlap_ranges,
                                       image_size, config, emulate)
    V = gaussians2d.shape[0]
    heur = (torch.zeros((V, 2), dtype=gaussians2d.dtype) if config.compute_point_heuristic
            else torch.empty((0, 2), dtype=gaussians2d.dtype))
    ctx.args = (overlap_to_point, tile_overlap_ranges, image_size, config)
    ctx.heur = heur
    ctx.needs = (gaussians2d.requires_grad, features.requires_grad)
    ctx.mark_non_differentiable(alpha, heur, vis)
    ctx.save_for_backward(gaussians2d, features, image)
    return image, alpha, heur, vis

  @staticmethod
  def backward(ctx, grad_image, grad_alpha, grad_heur, grad_vis):
    gaussians2d, features, image = ctx.saved_tensors
    o2p, ranges, image_size, config = ctx.args
    gp, gf = raster_backward(gaussians2d, features, o2p, ranges, image_size, config, image,
                             grad_image.contiguous(), True, True, ctx.heur)
    return gp, gf, None, None, None, None, None


def rasterize_with_tiles(gaussians2d, features, overlap_to_point, tile_overlap_ranges, image_size, config,
                         emulate_stale_tail=True) -> RasterOut:
  return RasterOut(*_Rasterize.apply(gaussians2d, features, overlap_to_point, tile_overlap_ranges,
                                     image_size, config, emulate_stale_tail))


def rasterize(gaussians2d, depth, features, image_size, config, use_depth16=False, emulate_stale_tail=True):
  o2p, ranges = map_to_tiles(gaussians2d.detach().to(torch.float32), depth.detach().to(torch.float32),
                             image_size, config, use_depth16)
  return rasterize_with_tiles(gaussians2d, features, o2p, ranges.view(-1, 2), image_size, config,
                              emulate_stale_tail)
