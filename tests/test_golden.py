"""Golden-vector tests: the fixtures under tests/golden/*.npz were produced FROM THE REFERENCE ITSELF by
tests/golden/make_golden.py (the reference's torch implementation, and its Taichi kernel source executed by the
emulator tests/golden/ti_emu.py).  They pin

  * the oracle (CPU, `-m "not gpu"`): tile maps bit-exact, rasterizer images / gradients, projection visible
    sets / outputs / gradients, SH outputs / gradients;
  * the CUDA path (`-m gpu`): the same comparisons through the public operators / the C ABI.

Tolerances are BASELINE.json's: bit-exact integer outputs, images 1e-5 relative L2, gradients 1e-4 relative L2
(the f64 torch_lib comparisons use allclose like the reference's tests/test_projection.py:36-48).
"""
from pathlib import Path

import numpy as np
import pytest
import torch

import oracle
from oracle import torch_ref
from taichi_gaussian_rasterizer_b200 import RasterConfig
from util import GRAD_REL_L2, IMAGE_REL_L2, rel_l2

GOLDEN = Path(__file__).resolve().parent / "golden"


def load(name):
  path = GOLDEN / f"{name}.npz"
  assert path.exists(), f"{path} is missing: run tests/golden/make_golden.py where /root/reference is available"
  return np.load(path)


def T(a):
  return torch.from_numpy(np.ascontiguousarray(a))


def cases(name):
  return range(int(load(name)["num_cases"]))


def tile_case(d, i):
  size = tuple(int(x) for x in d[f"c{i}_image_size"])
  cfg = RasterConfig(tile_size=int(d[f"c{i}_tile_size"]))
  return T(d[f"c{i}_gaussians"]), T(d[f"c{i}_depth"]), size, cfg, bool(d[f"c{i}_use_depth16"])


def raster_case(d, i):
  size = tuple(int(x) for x in d[f"c{i}_image_size"])
  cfg = RasterConfig(tile_size=int(d[f"c{i}_tile_size"]), pixel_stride=tuple(int(x) for x in d[f"c{i}_pixel_stride"]),
                     antialias=bool(d[f"c{i}_antialias"]), compute_visibility=True, compute_point_heuristic=True)
  return size, cfg


def check_raster(d, i, out, grad_g, grad_f):
  assert rel_l2(out.image, T(d[f"c{i}_image"])) < IMAGE_REL_L2
  assert rel_l2(out.image_weight, T(d[f"c{i}_image_weight"])) < IMAGE_REL_L2
  assert rel_l2(out.visibility, T(d[f"c{i}_visibility"])) < GRAD_REL_L2
  assert rel_l2(out.point_heuristic, T(d[f"c{i}_point_heuristic"])) < GRAD_REL_L2
  assert rel_l2(grad_g, T(d[f"c{i}_grad_gaussians"])) < GRAD_REL_L2
  assert rel_l2(grad_f, T(d[f"c{i}_grad_features"])) < GRAD_REL_L2


# ----------------------------------------------------------------------------------------------- oracle (CPU)
@pytest.mark.parametrize("i", cases("tile_map"))
def test_oracle_tile_map_matches_reference_kernels(i):
  d = load("tile_map")
  g, depth, size, cfg, d16 = tile_case(d, i)
  o2p, ranges = oracle.map_to_tiles(g, depth, size, cfg, use_depth16=d16)
  assert torch.equal(o2p, T(d[f"c{i}_overlap_to_point"])), "overlap_to_point differs from the reference's kernels"
  assert torch.equal(ranges, T(d[f"c{i}_tile_ranges"])), "tile_ranges differ from the reference's kernels"


@pytest.mark.parametrize("i", cases("raster"))
def test_oracle_rasterizer_matches_reference_kernels(i):
  d = load("raster")
  size, cfg = raster_case(d, i)
  g = T(d[f"c{i}_gaussians"]).requires_grad_(True)
  f = T(d[f"c{i}_features"]).requires_grad_(True)
  out = oracle.rasterize_with_tiles(g, f, T(d[f"c{i}_overlap_to_point"]), T(d[f"c{i}_tile_ranges"]).view(-1, 2), size, cfg)
  (out.image * T(d[f"c{i}_grad_image"])).sum().backward()
  check_raster(d, i, out, g.grad, f.grad)


def cases2():
  d = load("raster2")
  return [i for i in range(int(d["num_cases"])) if f"c{i}_gaussians" in d.files]


@pytest.mark.parametrize("i", cases2())
def test_oracle_rasterizer_matches_reference_kernels_tile16(i):
  """raster2.npz (round 2): the reference's kernels at tile 16 / stride (2, 2) / statistics on with tile lists of
  330 and ~600 entries (two and three groups, C mod 256 != 0: stale slots at the measured configuration), 34 feature
  channels, and the antialiased pdf."""
  d = load("raster2")
  size, cfg = raster_case(d, i)
  g = T(d[f"c{i}_gaussians"]).requires_grad_(True)
  f = T(d[f"c{i}_features"]).requires_grad_(True)
  out = oracle.rasterize_with_tiles(g, f, T(d[f"c{i}_overlap_to_point"]), T(d[f"c{i}_tile_ranges"]).view(-1, 2), size, cfg)
  (out.image * T(d[f"c{i}_grad_image"])).sum().backward()
  check_raster(d, i, out, g.grad, f.grad)


def test_raster2_fixture_exercises_stale_tail_at_tile16():
  d = load("raster2")
  hit = 0
  for i in cases2():
    size, cfg = raster_case(d, i)
    ranges = T(d[f"c{i}_tile_ranges"]).view(-1, 2)
    counts = ranges[:, 1] - ranges[:, 0]
    if not ((counts > 256) & (counts % 256 != 0)).any() or cfg.antialias:
      continue
    hit += 1
    args = (T(d[f"c{i}_gaussians"]), T(d[f"c{i}_features"]), T(d[f"c{i}_overlap_to_point"]), ranges, size, cfg)
    img_off, _, _ = oracle.raster_forward(*args, emulate_stale_tail=False)
    assert rel_l2(img_off, T(d[f"c{i}_image"])) > 100 * IMAGE_REL_L2, "fixture does not exercise the stale slots"
  assert hit >= 1


def test_raster_fixture_exercises_stale_tail():
  """Case 1 has tile lists longer than a block (C > A, C mod A != 0): without the stale-slot emulation
   the oracle must NOT match the reference's kernels, with it it must."""
  d = load("raster")
  i = 1
  size, cfg = raster_case(d, i)
  ranges = T(d[f"c{i}_tile_ranges"]).view(-1, 2)
  area = cfg.tile_size ** 2
  counts = (ranges[:, 1] - ranges[:, 0])
  assert ((counts > area) & (counts % area != 0)).any()
  args = (T(d[f"c{i}_gaussians"]), T(d[f"c{i}_features"]), T(d[f"c{i}_overlap_to_point"]), ranges, size, cfg)
  img_on, _, _ = oracle.raster_forward(*args, emulate_stale_tail=True)
  img_off, _, _ = oracle.raster_forward(*args, emulate_stale_tail=False)
  ref = T(d[f"c{i}_image"])
  assert rel_l2(img_on, ref) < IMAGE_REL_L2
  assert rel_l2(img_off, ref) > 100 * IMAGE_REL_L2


def proj_inputs(d, i, dtype):
  names = ["position", "log_scaling", "rotation", "alpha_logit", "T_camera_world", "projection"]
  ts = [T(d[f"c{i}_{n}"]).to(dtype) for n in names]
  size = tuple(int(x) for x in d[f"c{i}_image_size"])
  drange = tuple(float(x) for x in d[f"c{i}_depth_range"])
  return names, ts, size, drange, float(d[f"c{i}_blur_cov"])


@pytest.mark.parametrize("i", cases("projection"))
def test_oracle_projection_matches_reference(i):
  d = load("projection")
  names, ts, size, drange, blur = proj_inputs(d, i, torch.float64)
  # f64 oracle kernel vs the reference's torch implementation and its Taichi kernel (f64)
  p, z, idx = oracle.projection_forward(*ts, size, drange, blur_cov=blur)
  for src in ("torch", "ti64"):
    assert torch.equal(idx, T(d[f"c{i}_{src}_indexes"])), f"visible set differs from {src}"
    assert torch.allclose(p, T(d[f"c{i}_{src}_points"]), rtol=1e-7, atol=1e-9)
    assert torch.allclose(z, T(d[f"c{i}_{src}_depth"]), rtol=1e-9, atol=1e-12)
  # f32 oracle kernel vs the reference's f32 Taichi kernel
  _, ts32, _, _, _ = proj_inputs(d, i, torch.float32)
  p32, z32, idx32 = oracle.projection_forward(*ts32, size, drange, blur_cov=blur)
  assert torch.equal(idx32, T(d[f"c{i}_ti32_indexes"]))
  assert rel_l2(p32, T(d[f"c{i}_ti32_points"])) < IMAGE_REL_L2
  assert rel_l2(z32, T(d[f"c{i}_ti32_depth"])) < IMAGE_REL_L2
  # gradients of the torch restatement (what the GPU backward is checked against) vs the reference's autograd
  tg = [t.clone().requires_grad_(True) for t in ts]
  pr, zr, _ = torch_ref.projection_apply(*tg, size, drange, blur_cov=blur)
  ((pr * T(d[f"c{i}_grad_out_points"])).sum() + (zr * T(d[f"c{i}_grad_out_depth"])).sum()).backward()
  for n, t in zip(names, tg):
    assert torch.allclose(t.grad, T(d[f"c{i}_grad_{n}"]), rtol=1e-6, atol=1e-9), f"grad {n}"


@pytest.mark.parametrize("i", cases("sh"))
def test_oracle_sh_matches_reference(i):
  d = load("sh")
  params, points, cam = (T(d[f"c{i}_{n}"]) for n in ("params", "points", "camera_pos"))
  idx = T(d[f"c{i}_indexes"])
  out32 = oracle.evaluate_sh_at(params, points, idx, cam)
  assert torch.allclose(out32.double(), T(d[f"c{i}_torch_out"]), atol=1e-5)       # tests/util.py:49-71 tolerance
  assert torch.allclose(out32, T(d[f"c{i}_ti32_out"]), atol=1e-6)
  ts = [t.double().requires_grad_(True) for t in (params, points, cam)]
  o = torch_ref.evaluate_sh_at(ts[0], ts[1], idx, ts[2])
  assert torch.allclose(o, T(d[f"c{i}_torch_out"]), rtol=1e-9, atol=1e-12)
  (o * T(d[f"c{i}_grad_out"])).sum().backward()
  for n, t in zip(("params", "points", "camera_pos"), ts):
    g = t.grad if t.grad is not None else torch.zeros_like(t)
    assert torch.allclose(g, T(d[f"c{i}_grad_{n}"]), rtol=1e-7, atol=1e-10), f"grad {n}"


@pytest.mark.parametrize("i", cases("projection"))
def test_oracle_projection_backward_matches_reference(i):
  """oracle.projection_backward (the hand restated reverse sweep, f64) against the gradients torch autograd gave
  through the reference's UNMODIFIED torch_lib/projection.py `apply` (make_golden.py), all six inputs; and its f32
  instantiation against the same gradients where the eigen decomposition is well conditioned."""
  d = load("projection")
  names, ts, size, drange, blur = proj_inputs(d, i, torch.float64)
  idx = T(d[f"c{i}_torch_indexes"])
  go_p, go_z = T(d[f"c{i}_grad_out_points"]), T(d[f"c{i}_grad_out_depth"])
  grads, cond = oracle.projection_backward(*ts, size, idx, go_p, go_z, blur_cov=blur)
  for n in names:
    ref = T(d[f"c{i}_grad_{n}"])
    assert torch.allclose(grads[n].reshape(ref.shape), ref, rtol=1e-7, atol=1e-10), f"{n}: {rel_l2(grads[n].reshape(ref.shape), ref)}"
  _, ts32, _, _, _ = proj_inputs(d, i, torch.float32)
  g32, cond32 = oracle.projection_backward(*ts32, size, idx, go_p.float(), go_z.float(), blur_cov=blur)
  good = torch.zeros(ts[0].shape[0], dtype=torch.bool)
  good[idx[(cond.min(dim=1).values > 0.05)]] = True
  assert int(good.sum()) > 0.5 * idx.shape[0]
  for n in ("position", "log_scaling", "rotation", "alpha_logit"):
    ref = T(d[f"c{i}_grad_{n}"])[good]
    assert rel_l2(g32[n][good], ref) < GRAD_REL_L2, f"f32 {n}: {rel_l2(g32[n][good], ref)}"


# ----------------------------------------------------------------------------------------------- CUDA path
@pytest.mark.gpu
@pytest.mark.parametrize("i", cases("tile_map"))
def test_gpu_tile_map_matches_reference_kernels(cuda_device, i):
  from taichi_gaussian_rasterizer_b200 import map_to_tiles
  d = load("tile_map")
  g, depth, size, cfg, d16 = tile_case(d, i)
  o2p, ranges = map_to_tiles(g.to(cuda_device), depth.to(cuda_device), size, cfg, use_depth16=d16)
  assert torch.equal(o2p.cpu(), T(d[f"c{i}_overlap_to_point"]))
  assert torch.equal(ranges.cpu(), T(d[f"c{i}_tile_ranges"]))


@pytest.mark.gpu
@pytest.mark.parametrize("i", cases("raster"))
def test_gpu_rasterizer_matches_reference_kernels(cuda_device, i):
  from taichi_gaussian_rasterizer_b200 import rasterize_with_tiles
  d = load("raster")
  size, cfg = raster_case(d, i)
  g = T(d[f"c{i}_gaussians"]).to(cuda_device).requires_grad_(True)
  f = T(d[f"c{i}_features"]).to(cuda_device).requires_grad_(True)
  out = rasterize_with_tiles(g, f, T(d[f"c{i}_overlap_to_point"]).to(cuda_device),
                             T(d[f"c{i}_tile_ranges"]).view(-1, 2).to(cuda_device), size, cfg)
  (out.image * T(d[f"c{i}_grad_image"]).to(cuda_device)).sum().backward()
  check_raster(d, i, out, g.grad, f.grad)


@pytest.mark.gpu
@pytest.mark.parametrize("i", cases2())
def test_gpu_fast_rasterizer_matches_reference_kernels_tile16(cuda_device, i):
  """The measured (fast) kernels DIRECTLY against the reference's kernels: tile 16, lists longer than a group with
  C mod 256 != 0 (the fast forward's re-walk of the stale slots), F = 34 (wide kernels), statistics on."""
  from taichi_gaussian_rasterizer_b200 import rasterize_with_tiles
  d = load("raster2")
  size, cfg = raster_case(d, i)
  g = T(d[f"c{i}_gaussians"]).to(cuda_device).requires_grad_(True)
  f = T(d[f"c{i}_features"]).to(cuda_device).requires_grad_(True)
  out = rasterize_with_tiles(g, f, T(d[f"c{i}_overlap_to_point"]).to(cuda_device),
                             T(d[f"c{i}_tile_ranges"]).view(-1, 2).to(cuda_device), size, cfg)
  (out.image * T(d[f"c{i}_grad_image"]).to(cuda_device)).sum().backward()
  check_raster(d, i, out, g.grad, f.grad)


@pytest.mark.gpu
@pytest.mark.parametrize("i", cases("projection"))
def test_gpu_projection_matches_reference(cuda_device, i):
  from taichi_gaussian_rasterizer_b200.perspective import projection as gpu_proj
  d = load("projection")
  names, ts, size, drange, blur = proj_inputs(d, i, torch.float64)
  tg = [t.to(cuda_device).requires_grad_(True) for t in ts]
  p, z, idx = gpu_proj.apply(*tg, size, drange, blur_cov=blur)
  assert torch.equal(idx.cpu(), T(d[f"c{i}_torch_indexes"]))
  assert torch.allclose(p.cpu(), T(d[f"c{i}_torch_points"]), rtol=1e-7, atol=1e-9)
  assert torch.allclose(z.cpu(), T(d[f"c{i}_torch_depth"]), rtol=1e-9, atol=1e-12)
  ((p * T(d[f"c{i}_grad_out_points"]).to(cuda_device)).sum() + (z * T(d[f"c{i}_grad_out_depth"]).to(cuda_device)).sum()).backward()
  for n, t in zip(names, tg):
    assert torch.allclose(t.grad.cpu(), T(d[f"c{i}_grad_{n}"]), rtol=1e-6, atol=1e-9), f"grad {n}"
  # f32 kernels against the reference's f32 Taichi kernel
  _, ts32, _, _, _ = proj_inputs(d, i, torch.float32)
  p32, z32, idx32 = gpu_proj.apply(*[t.to(cuda_device) for t in ts32], size, drange, blur_cov=blur)
  assert torch.equal(idx32.cpu(), T(d[f"c{i}_ti32_indexes"]))
  assert rel_l2(p32, T(d[f"c{i}_ti32_points"])) < IMAGE_REL_L2


@pytest.mark.gpu
@pytest.mark.parametrize("i", cases("sh"))
def test_gpu_sh_matches_reference(cuda_device, i):
  from taichi_gaussian_rasterizer_b200 import evaluate_sh_at
  d = load("sh")
  idx = T(d[f"c{i}_indexes"]).to(cuda_device)
  ts = [T(d[f"c{i}_{n}"]).double().to(cuda_device).requires_grad_(True) for n in ("params", "points", "camera_pos")]
  o = evaluate_sh_at(ts[0], ts[1], idx, ts[2])
  assert torch.allclose(o.cpu(), T(d[f"c{i}_torch_out"]), rtol=1e-9, atol=1e-12)
  (o * T(d[f"c{i}_grad_out"]).to(cuda_device)).sum().backward()
  for n, t in zip(("params", "points", "camera_pos"), ts):
    assert torch.allclose(t.grad.cpu(), T(d[f"c{i}_grad_{n}"]), rtol=1e-7, atol=1e-10), f"grad {n}"
  ts32 = [T(d[f"c{i}_{n}"]).to(cuda_device) for n in ("params", "points", "camera_pos")]
  o32 = evaluate_sh_at(ts32[0], ts32[1], idx, ts32[2])
  assert torch.allclose(o32.cpu(), T(d[f"c{i}_ti32_out"]), atol=1e-6)
