"""TEST INFRASTRUCTURE — CPU restatement (numpy, float32 / integer) of the reference's Morton coding,
/root/reference/taichi_splatting/misc/morton_sort.py: ``grid_at_resolution`` (:114-118), ``Grid.get_inc`` /
``grid_cell`` (:42-54), ``spreads_bits32/64`` (:13-30), ``cell_code32/64`` (:69-88), ``argsort`` (:121-126: codes
sorted by a stable radix sort).  Pinned by tests/golden/morton.npz, produced by the reference's own kernels under
tests/golden/ti_emu.py.  Only tests/ may import this module."""
import numpy as np


def grid_at_resolution(points: np.ndarray, resolution: float, size: int = 2 ** 20):
  lower = points.astype(np.float32).min(axis=0)
  upper = (lower + np.float32(size * resolution)).astype(np.float32)   # torch: f32 tensor + python scalar
  return lower, upper, size


def grid_cells(points: np.ndarray, lower, upper, size: int) -> np.ndarray:
  inc = ((upper - lower) / np.float32(size)).astype(np.float32)        # morton_sort.py:43
  v = ((points.astype(np.float32) - lower) / inc).astype(np.float32)   # :52
  v = np.minimum(np.maximum(v, np.float32(0)), np.float32(size - 1))   # :53 clamp, then truncation to u32
  return v.astype(np.uint32)


def spread_bits32(x: np.ndarray) -> np.ndarray:
  x = x.astype(np.uint32) & np.uint32(0x3ff)
  for shift, mask in ((16, 0x030000FF), (8, 0x0300F00F), (4, 0x030C30C3), (2, 0x09249249)):
    x = (x | (x << np.uint32(shift))) & np.uint32(mask)
  return x


def spread_bits64(x: np.ndarray) -> np.ndarray:
  x = x.astype(np.uint64) & np.uint64(0x1fffff)
  for shift, mask in ((32, 0x1f00000000ffff), (16, 0x1f0000ff0000ff), (8, 0x100f00f00f00f00f),
                      (4, 0x10c30c30c30c30c3), (2, 0x1249249249249249)):
    x = (x | (x << np.uint64(shift))) & np.uint64(mask)
  return x


def morton_codes(points: np.ndarray, lower, upper, size: int, bits: int = 64) -> np.ndarray:
  c = grid_cells(points, lower, upper, size)
  if bits == 64:
    return spread_bits64(c[:, 0]) | (spread_bits64(c[:, 1]) << np.uint64(1)) | (spread_bits64(c[:, 2]) << np.uint64(2))
  return spread_bits32(c[:, 0]) | (spread_bits32(c[:, 1]) << np.uint32(1)) | (spread_bits32(c[:, 2]) << np.uint32(2))


def argsort(points: np.ndarray, resolution: float) -> np.ndarray:
  lower, upper, size = grid_at_resolution(points, resolution)
  return np.argsort(morton_codes(points, lower, upper, size, 64), kind="stable").astype(np.int32)
