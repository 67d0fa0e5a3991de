"""Morton (Z-order) sorting of 3D points: spatial reordering of the gaussians so that neighbours in space are
neighbours in memory (better gather locality for the projection, SH and rasterizer record loads).

Surface of taichi_splatting/misc/morton_sort.py:114-152: ``grid_at_resolution``, ``argsort``, ``sort``,
``argsort_dedup``, ``sort_dedup`` (+ ``morton_codes`` for the codes themselves).  The codes come from one CUDA kernel
(csrc/morton.cu, ``gs_morton_codes``) and are sorted by the package's onesweep radix sort on exactly the bits a grid
of that size can set; nothing is read back to the host (the reference builds its ``Grid`` from host scalars, which
synchronises on ``points.min``).

``argsort_dedup`` in the reference passes the float points as sort values and indexes the codes with a tuple
(morton_sort.py:138-141), which cannot run; here it does what its name and the use of ``unique_consecutive`` say: one
representative (the last in sorted order) per occupied grid cell, returned as indexes into ``points``.
"""
import ctypes
from typing import NamedTuple

import torch

from .. import _native as N
from .. import cuda_lib


class Grid(NamedTuple):
  """Uniform grid: ``lower`` (3,) and ``upper`` (3,) device tensors, ``size`` cells per axis (morton_sort.py:33-47)."""
  lower: torch.Tensor
  upper: torch.Tensor
  size: int

  @property
  def inc(self) -> torch.Tensor:
    return (self.upper - self.lower) / float(self.size)


def grid_at_resolution(points: torch.Tensor, resolution: float, size: int = 2 ** 20) -> Grid:
  lower = points.min(dim=0).values
  upper = lower + size * resolution
  return Grid(lower, upper, size)


def morton_codes(points: torch.Tensor, grid: Grid, bits: int = 64) -> torch.Tensor:
  """uint64 (``bits`` = 64, grid size <= 2^21) or uint32 (``bits`` = 32, grid size <= 2^10) Morton code per point."""
  N.require_cuda(points)
  assert points.ndim == 2 and points.shape[1] == 3 and points.dtype == torch.float32, \
    f"points must be (N, 3) float32, got {tuple(points.shape)} {points.dtype}"
  assert bits in (32, 64)
  points = points.contiguous()
  codes = torch.empty(points.shape[0], dtype=torch.uint64 if bits == 64 else torch.uint32, device=points.device)
  lower = grid.lower.to(device=points.device, dtype=torch.float32).contiguous()
  inc = grid.inc.to(device=points.device, dtype=torch.float32).contiguous()
  N.call("gs_morton_codes", ctypes.c_int64(points.shape[0]), N.ptr(points), N.ptr(lower), N.ptr(inc),
         ctypes.c_int64(grid.size), ctypes.c_int32(bits), N.ptr(codes), N.stream_ptr(points.device))
  return codes


def _code_bits(size: int) -> int:
  return 3 * max(1, (size - 1).bit_length())


def _sorted_codes(points: torch.Tensor, resolution: float, size: int = 2 ** 20):
  grid = grid_at_resolution(points, resolution, size=size)
  codes = morton_codes(points, grid, bits=64)
  idx = torch.arange(points.shape[0], dtype=torch.int32, device=points.device)
  return cuda_lib.radix_sort_pairs(codes, idx, 0, _code_bits(size))


def argsort(points: torch.Tensor, resolution: float) -> torch.Tensor:
  """Indexes (int32) that put ``points`` in Morton order on a 2^20 grid of cell size ``resolution`` anchored at the
  minimum corner (stable: equal codes keep their order)."""
  return _sorted_codes(points, resolution)[1]


def sort(points: torch.Tensor, resolution: float) -> torch.Tensor:
  return points[argsort(points, resolution).long()]


def argsort_dedup(points: torch.Tensor, resolution: float) -> torch.Tensor:
  """One index per occupied grid cell, in Morton order."""
  codes, idx = _sorted_codes(points, resolution)
  _, counts = torch.unique_consecutive(codes.view(torch.int64), return_counts=True)
  last = torch.cumsum(counts, dim=0) - 1
  return idx[last]


def sort_dedup(points: torch.Tensor, resolution: float) -> torch.Tensor:
  return points[argsort_dedup(points, resolution).long()]


__all__ = ["Grid", "grid_at_resolution", "morton_codes", "argsort", "sort", "argsort_dedup", "sort_dedup"]
