"""Oracle checks for projection + SH (CPU): the C++ restatement against the torch restatement, the
float64 gradcheck the reference runs (tests/test_projection.py:100-116,
tests/test_spherical_harmonics.py:52-62), and the host image of the CUDA math against the oracle
bit for bit (catches an operation-order divergence without a GPU)."""
import ctypes

import numpy as np
import pytest
import torch

import oracle
from oracle import torch_ref
from taichi_gaussian_rasterizer_b200 import _native
from util import scene3d


def _inputs(seed, n, dtype):
  gaussians, camera = scene3d(seed, n, margin=0.5, scale_factor=0.1)
  g = gaussians.to(dtype=dtype)
  cam = camera.to(dtype=dtype)
  return g, cam


@pytest.mark.parametrize("seed", range(8))
def test_oracle_matches_torch_restatement_f64(seed):
  torch.manual_seed(seed)
  n = int(torch.randint(1, 3000, (1,)))
  g, cam = _inputs(seed, n, torch.float64)
  args = (*g.shape_tensors(), cam.T_camera_world, cam.projection, cam.image_size, cam.depth_range)
  p1, d1, i1 = oracle.projection_forward(*args, blur_cov=0.3)
  p2, d2, i2 = torch_ref.projection_apply(*args, blur_cov=0.3)
  assert i1.shape == i2.shape and (i1 == i2).all(), "visible index set differs"
  assert torch.allclose(p1, p2, rtol=1e-9, atol=1e-9)
  assert torch.allclose(d1, d2, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("seed", range(4))
def test_oracle_f32_close_to_f64(seed):
  g, cam = _inputs(seed, 2000, torch.float64)
  args64 = (*g.shape_tensors(), cam.T_camera_world, cam.projection, cam.image_size, cam.depth_range)
  g32, cam32 = g.to(dtype=torch.float32), cam.to(dtype=torch.float32)
  args32 = (*g32.shape_tensors(), cam32.T_camera_world, cam32.projection, cam32.image_size, cam32.depth_range)
  p64, d64, i64 = oracle.projection_forward(*args64, blur_cov=0.3)
  p32, d32, i32 = oracle.projection_forward(*args32, blur_cov=0.3)
  common = np.intersect1d(i64.numpy(), i32.numpy())
  assert len(common) >= 0.98 * max(len(i64), 1)
  m64 = torch.isin(i64, torch.from_numpy(common))
  m32 = torch.isin(i32, torch.from_numpy(common))
  assert torch.allclose(p32[m32].double(), p64[m64], rtol=2e-3, atol=2e-3)


def test_torch_restatement_gradcheck():
  for seed in range(3):
    g, cam = _inputs(seed, 12, torch.float64)
    tensors = [t.detach().clone().requires_grad_(True) for t in
               (*g.shape_tensors(), cam.T_camera_world, cam.projection)]

    def f(*ts):
      pts, depth, _ = torch_ref.projection_apply(*ts, cam.image_size, cam.depth_range, blur_cov=0.3)
      return pts, depth
    assert torch.autograd.gradcheck(f, tensors, eps=1e-6, atol=1e-5)


@pytest.mark.parametrize("degree", [0, 1, 2, 3])
def test_sh_oracle_matches_torch(degree):
  torch.manual_seed(degree)
  n, k = 57, 3
  params = torch.rand(n, k, (degree + 1) ** 2, dtype=torch.float64)
  points = torch.randn(n, 3, dtype=torch.float64)
  cam = torch.randn(3, dtype=torch.float64)
  idx = torch.randint(0, n, (n // 2,))
  a = oracle.evaluate_sh_at(params, points, idx, cam)
  b = torch_ref.evaluate_sh_at(params, points, idx, cam)
  assert torch.allclose(a, b, rtol=1e-12, atol=1e-12)
  a32 = oracle.evaluate_sh_at(params.float(), points.float(), idx, cam.float())
  assert torch.allclose(a32.double(), b, atol=1e-5)


def test_sh_torch_gradcheck():
  torch.manual_seed(0)
  params = (torch.rand(9, 2, 16, dtype=torch.float64) * 0.2).requires_grad_(True)
  points = torch.randn(9, 3, dtype=torch.float64).requires_grad_(True)
  cam = torch.randn(3, dtype=torch.float64).requires_grad_(True)
  idx = torch.randint(0, 9, (5,))
  assert torch.autograd.gradcheck(lambda p, x, c: torch_ref.evaluate_sh_at(p, x, idx, c), (params, points, cam))


def test_cuda_host_image_is_bit_identical_to_oracle():
  """geom_math.cuh compiled for the host (gs_selftest_project_one_f32) vs oracle.cpp, same inputs."""
  lib = _native.lib()
  fn = lib.gs_selftest_project_one_f32
  fn.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                         ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p]
  mismatches = 0
  total = 0
  for seed in range(4):
    g, cam = scene3d(seed, 1500, margin=0.5, scale_factor=0.3)
    Tcw = cam.T_camera_world.float().contiguous()
    proj = cam.projection.float().contiguous()
    pos, ls, rot, logit = [t.float().contiguous() for t in g.shape_tensors()]
    n = pos.shape[0]
    points = torch.empty((n, 7)); depth = torch.empty((n,))
    oracle.lib().orc_project_fwd_f32(
      ctypes.c_int64(n), *[ctypes.c_void_p(t.data_ptr()) for t in (pos, ls, rot, logit, Tcw, proj)],
      ctypes.c_int(cam.image_size[0]), ctypes.c_int(cam.image_size[1]), ctypes.c_double(cam.near_plane),
      ctypes.c_double(cam.far_plane), ctypes.c_double(0.3), ctypes.c_double(0.15), ctypes.c_double(1 / 255.),
      ctypes.c_void_p(points.data_ptr()), ctypes.c_void_p(depth.data_ptr()))
    out8 = torch.empty(8)
    inview = ctypes.c_int(0)
    for i in range(n):
      fn(pos[i].data_ptr(), ls[i].data_ptr(), rot[i].data_ptr(), float(logit[i]), Tcw.data_ptr(), proj.data_ptr(),
         cam.image_size[0], cam.image_size[1], cam.near_plane, cam.far_plane, 0.3, 0.15, 1 / 255.,
         out8.data_ptr(), ctypes.byref(inview))
      total += 1
      visible = depth[i].item() != 0.0
      if bool(inview.value) != visible:
        mismatches += 1
      elif visible:
        a = np.concatenate([points[i].numpy(), depth[i:i + 1].numpy()]).view(np.uint32)
        b = out8.numpy().view(np.uint32)
        if not (a == b).all():
          mismatches += 1
  assert mismatches == 0, f"{mismatches}/{total} gaussians differ between geom_math.cuh (host) and oracle.cpp"
