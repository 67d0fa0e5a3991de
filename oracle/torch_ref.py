"""Differentiable pure-torch restatements (TEST INFRASTRUCTURE).

Used for autograd cross-checks of the hand-derived CUDA backward passes, the same way the
reference checks its Taichi autodiff against torch (tests/test_projection.py:76-96,
tests/test_spherical_harmonics.py:33-45).  They restate the KERNEL semantics
(perspective/projection.py:50-118, spherical_harmonics.py:39-134), not torch_lib; the committed
golden fixtures pin them against the reference's torch_lib.
"""
import math

import torch


def quat_to_mat(q):
  x, y, z, w = q.unbind(-1)  # x y z w order, taichi_lib/generic.py:418-427
  x2, y2, z2 = x * x, y * y, z * z
  m = [1 - 2 * y2 - 2 * z2, 2 * x * y - 2 * w * z, 2 * x * z + 2 * w * y,
       2 * x * y + 2 * w * z, 1 - 2 * x2 - 2 * z2, 2 * y * z - 2 * w * x,
       2 * x * z - 2 * w * y, 2 * y * z + 2 * w * x, 1 - 2 * x2 - 2 * y2]
  return torch.stack(m, -1).reshape(q.shape[:-1] + (3, 3))


def project_all(position, log_scaling, rotation, alpha_logit, T_camera_world, projection, image_size,
                blur_cov=0.0, clamp_margin=0.15):
  """Projection math for every gaussian (no culling): returns points (N,7), depth (N,1)."""
  dtype = position.dtype
  w, h = image_size
  size = torch.tensor([w, h], dtype=dtype, device=position.device)
  q = rotation / rotation.norm(dim=-1, keepdim=True)
  s = torch.exp(log_scaling)
  W = T_camera_world[:3, :3]
  t = T_camera_world[:3, 3]
  cam = position @ W.T + t
  f, c = projection[0:2], projection[2:4]
  z = cam[:, 2]
  uv = f * cam[:, 0:2] / z.unsqueeze(1) + c
  lo = -size * clamp_margin
  hi = (size - 1) * (1 + clamp_margin)
  tc = torch.maximum(torch.minimum(uv, hi), lo)  # clamp, gradient 1 strictly inside
  zero = torch.zeros_like(z)
  J = torch.stack([f[0] / z, zero, -(tc[:, 0] - c[0]) / z,
                   zero, f[1] / z, -(tc[:, 1] - c[1]) / z], -1).reshape(-1, 2, 3)
  RS = quat_to_mat(q) * s.unsqueeze(1)
  m = J @ (W.unsqueeze(0) @ RS)
  cov = m @ m.transpose(1, 2)
  a = cov[:, 0, 0] + blur_cov
  b = cov[:, 0, 1]
  cc = cov[:, 1, 1] + blur_cov
  tr = a + cc
  det = a * cc - b * b
  gap = tr * tr - 4 * det
  sg = torch.sqrt(torch.clamp_min(gap, 0))
  l1 = (tr + sg) * 0.5
  l2 = (tr - sg) * 0.5
  v = torch.stack([a - l2, b], -1)
  v1 = v / v.norm(dim=-1, keepdim=True)
  sigma = torch.sqrt(torch.stack([l1, l2], -1))
  alpha = torch.sigmoid(alpha_logit.reshape(-1, 1))
  points = torch.cat([uv, v1, sigma, alpha], dim=-1)
  return points, z.unsqueeze(1)


def in_view_mask(points, depth, image_size, depth_range, alpha_threshold=1. / 255.):
  w, h = image_size
  mean, v1, sigma, alpha = points[:, 0:2], points[:, 2:4], points[:, 4:6], points[:, 6]
  g = torch.sqrt(2 * torch.log(alpha / alpha_threshold))
  sx, sy = sigma[:, 0] * g, sigma[:, 1] * g
  v2 = torch.stack([-v1[:, 1], v1[:, 0]], -1)
  extent = torch.sqrt((v1 * sx.unsqueeze(1)) ** 2 + (v2 * sy.unsqueeze(1)) ** 2)
  lower, upper = mean - extent, mean + extent
  z = depth[:, 0]
  size = torch.tensor([w, h], dtype=points.dtype, device=points.device)
  return ((z > depth_range[0]) & (z < depth_range[1]) & (upper > 0).all(1) & (lower < size).all(1))


def projection_apply(position, log_scaling, rotation, alpha_logit, T_camera_world, projection, image_size,
                     depth_range, blur_cov=0.0, clamp_margin=0.15, alpha_threshold=1. / 255.):
  """Same signature and outputs as perspective/projection.py:190-215 ``apply``."""
  points, depth = project_all(position, log_scaling, rotation, alpha_logit, T_camera_world, projection,
                              image_size, blur_cov, clamp_margin)
  with torch.no_grad():
    idx = in_view_mask(points, depth, image_size, depth_range, alpha_threshold).nonzero(as_tuple=True)[0]
  return points[idx], depth[idx], idx


def rsh_cart(degree, d):
  x, y, z = d.unbind(-1)
  out = [torch.full_like(x, 0.282094791773878)]
  if degree >= 1:
    out += [-0.48860251190292 * y, 0.48860251190292 * z, -0.48860251190292 * x]
  if degree >= 2:
    x2, y2, z2, xy, xz, yz = x * x, y * y, z * z, x * y, x * z, y * z
    out += [1.09254843059208 * xy, -1.09254843059208 * yz, 0.94617469575756 * z2 - 0.31539156525252,
            -1.09254843059208 * xz, 0.54627421529604 * x2 - 0.54627421529604 * y2]
  if degree >= 3:
    out += [-0.590043589926644 * y * (3.0 * x2 - y2), 2.89061144264055 * xy * z,
            0.304697199642977 * y * (1.5 - 7.5 * z2),
            1.24392110863372 * z * (1.5 * z2 - 0.5) - 0.497568443453487 * z,
            0.304697199642977 * x * (1.5 - 7.5 * z2), 1.44530572132028 * z * (x2 - y2),
            -0.590043589926644 * x * (x2 - 3.0 * y2)]
  return torch.stack(out, -1)


def evaluate_sh_at(sh_params, positions, indexes, camera_pos):
  """spherical_harmonics.py:118-134 in torch: (M,K,D),(M,3),(V,),(3,) -> (V,K)."""
  degree = int(math.isqrt(sh_params.shape[2])) - 1
  d = positions[indexes] - camera_pos.unsqueeze(0)
  d = d / d.norm(dim=1, keepdim=True)
  coeffs = rsh_cart(degree, d)
  out = torch.einsum('nd,nkd->nk', coeffs, sh_params[indexes])
  return torch.clamp(out + 0.5, 0., 1.)


def ndc_depth(depth, near, far):
  return 1 - (1. / depth - 1. / far) / (1. / near - 1. / far)
