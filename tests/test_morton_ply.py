"""Morton reordering (misc/morton_sort.py) and gaussian-splatting PLY scene IO.
CPU: the numpy oracle against the fixture the reference's own kernels produced (tests/golden/morton.npz), PLY
round trips and a hand-packed file.  GPU: the CUDA kernel + radix sort against the fixture and the oracle."""
import struct
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import morton_ref
from taichi_gaussian_rasterizer_b200.data_types import Gaussians3D
from taichi_gaussian_rasterizer_b200.misc import ply_io

GOLDEN = np.load(Path(__file__).parent / "golden" / "morton.npz")
CASES = range(int(GOLDEN["num_cases"]))


# ----------------------------------------------------------------------------------------------- oracle vs reference
@pytest.mark.parametrize("i", CASES)
def test_oracle_morton_codes_equal_the_reference_kernels(i):
  pts, res = GOLDEN[f"c{i}_points"], float(GOLDEN[f"c{i}_resolution"])
  lower, upper, size = morton_ref.grid_at_resolution(pts, res)
  assert np.array_equal(morton_ref.morton_codes(pts, lower, upper, size, 64), GOLDEN[f"c{i}_codes64"])
  lower, upper, size = morton_ref.grid_at_resolution(pts, res * 1024, size=2 ** 10)
  assert np.array_equal(morton_ref.morton_codes(pts, lower, upper, size, 32), GOLDEN[f"c{i}_codes32"])
  assert np.array_equal(morton_ref.argsort(pts, res), GOLDEN[f"c{i}_argsort"])


def test_morton_order_is_spatially_coherent():
  rng = np.random.default_rng(0)
  pts = rng.random((4000, 3), dtype=np.float32)
  order = morton_ref.argsort(pts, 1e-5)
  step_sorted = np.linalg.norm(np.diff(pts[order], axis=0), axis=1).mean()
  step_random = np.linalg.norm(np.diff(pts, axis=0), axis=1).mean()
  assert step_sorted < 0.25 * step_random


# ----------------------------------------------------------------------------------------------- PLY
def random_scene(n, k, seed=0):
  g = torch.Generator().manual_seed(seed)
  return Gaussians3D(position=torch.randn(n, 3, generator=g), log_scaling=torch.randn(n, 3, generator=g),
                     rotation=torch.nn.functional.normalize(torch.randn(n, 4, generator=g), dim=1),
                     alpha_logit=torch.randn(n, 1, generator=g), feature=torch.randn(n, 3, k, generator=g),
                     batch_size=(n,))


@pytest.mark.parametrize("k", [1, 4, 16])
def test_ply_round_trip(tmp_path, k):
  scene = random_scene(37, k)
  ply_io.save_ply(scene, tmp_path / "scene.ply")
  back = ply_io.load_ply(tmp_path / "scene.ply")
  for name, t in scene.items():
    assert torch.equal(getattr(back, name), t), name
  assert ply_io.load_ply(tmp_path / "scene.ply", max_sh_degree=0).feature.shape == (37, 3, 1)


def test_ply_layout_is_the_gaussian_splatting_one(tmp_path):
  """A file packed by hand, property by property: quaternion w first, f_rest channel major, opacity a logit."""
  names = (["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"] + [f"f_rest_{i}" for i in range(9)] +
           ["opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"])
  rows = [[float(100 * r + c) for c in range(len(names))] for r in range(2)]
  header = "ply\nformat binary_little_endian 1.0\ncomment hand packed\nelement vertex 2\n" + \
           "".join(f"property float {n}\n" for n in names) + "end_header\n"
  path = tmp_path / "hand.ply"
  path.write_bytes(header.encode() + b"".join(struct.pack("<%df" % len(names), *r) for r in rows))
  g = ply_io.load_ply(path)
  col = {n: i for i, n in enumerate(names)}
  assert g.position.tolist() == [[0., 1., 2.], [100., 101., 102.]]
  assert g.rotation[0].tolist() == [float(col["rot_1"]), float(col["rot_2"]), float(col["rot_3"]), float(col["rot_0"])]
  assert g.alpha_logit[1].item() == 100. + col["opacity"] and g.log_scaling[0].tolist() == [19., 20., 21.]
  assert g.feature.shape == (2, 3, 4)
  assert g.feature[0, :, 0].tolist() == [6., 7., 8.]                     # DC per channel
  assert g.feature[0, 0, 1:].tolist() == [9., 10., 11.]                  # red's three degree-1 terms come first
  assert g.feature[0, 2, 1:].tolist() == [15., 16., 17.]
  with pytest.raises(ValueError):
    (tmp_path / "bad.ply").write_bytes(b"plx\n")
    ply_io.load_ply(tmp_path / "bad.ply")


# ----------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("i", CASES)
def test_cuda_morton_equals_the_reference_kernels(cuda_device, i):
  from taichi_gaussian_rasterizer_b200.misc import morton_sort
  pts = torch.from_numpy(GOLDEN[f"c{i}_points"]).to(cuda_device)
  res = float(GOLDEN[f"c{i}_resolution"])
  grid = morton_sort.grid_at_resolution(pts, res)
  assert np.array_equal(morton_sort.morton_codes(pts, grid, 64).cpu().numpy(), GOLDEN[f"c{i}_codes64"])
  grid32 = morton_sort.grid_at_resolution(pts, res * 1024, size=2 ** 10)
  assert np.array_equal(morton_sort.morton_codes(pts, grid32, 32).cpu().numpy(), GOLDEN[f"c{i}_codes32"])
  assert np.array_equal(morton_sort.argsort(pts, res).cpu().numpy(), GOLDEN[f"c{i}_argsort"])
  assert torch.equal(morton_sort.sort(pts, res), pts[torch.from_numpy(GOLDEN[f"c{i}_argsort"]).long().to(cuda_device)])


@pytest.mark.gpu
@pytest.mark.parametrize("n", [0, 1, 255, 256, 257, 1_000_003])
def test_cuda_morton_against_the_oracle_and_dedup(cuda_device, n):
  from taichi_gaussian_rasterizer_b200.misc import morton_sort
  torch.manual_seed(n)
  pts = (torch.rand(n, 3) * 40 - 7).float()
  if n > 300:
    pts[100:200] = pts[:100]      # exact duplicates
    pts[200] = float("nan")       # NaN lands in cell 0 on both sides
  res = 1e-3
  if n == 0:
    assert morton_sort.morton_codes(pts.to(cuda_device), morton_sort.Grid(torch.zeros(3), torch.ones(3), 8)).shape == (0,)
    return
  P = pts.numpy()
  lower = np.nanmin(P, axis=0).astype(np.float32)
  upper = (lower + np.float32(2 ** 20 * res)).astype(np.float32)
  grid = morton_sort.Grid(torch.from_numpy(lower).to(cuda_device), torch.from_numpy(upper).to(cuda_device), 2 ** 20)
  with np.errstate(invalid="ignore"):
    ref = morton_ref.morton_codes(np.nan_to_num(P, nan=-1e30), lower, upper, 2 ** 20, 64)
  codes = morton_sort.morton_codes(pts.to(cuda_device), grid, 64).cpu().numpy()
  assert np.array_equal(codes, ref)
  finite = pts.clone()
  finite[torch.isnan(finite)] = 0.
  d = finite.to(cuda_device)
  order = morton_sort.argsort(d, res).cpu().numpy()
  assert np.array_equal(order, morton_ref.argsort(finite.numpy(), res))
  keep = morton_sort.argsort_dedup(d, res).cpu().numpy()
  lo, up, size = morton_ref.grid_at_resolution(finite.numpy(), res)
  c = morton_ref.morton_codes(finite.numpy(), lo, up, size, 64)
  assert len(keep) == len(np.unique(c)) and np.all(np.diff(c[keep].astype(np.int64)) > 0)
  assert morton_sort.sort_dedup(d, res).shape == (len(keep), 3)
