"""Pinhole camera description used by the projection operator.

Same fields, properties and helper methods as taichi_splatting/perspective/params.py:9-101
(``projection`` = [fx, fy, cx, cy]; ``T_camera_world`` is the 4x4 view matrix;
``image_size`` is (width, height)).
"""
from dataclasses import dataclass, replace

import torch
from beartype import beartype
from beartype.typing import Tuple


@beartype
@dataclass
class CameraParams:
  projection: torch.Tensor      # (4,)  fx, fy, cx, cy
  T_camera_world: torch.Tensor  # (4, 4) world -> camera

  near_plane: float
  far_plane: float
  image_size: Tuple[int, int]   # (width, height)

  def __post_init__(self):
    assert self.projection.shape == (4,), f"Expected shape (4,), got {self.projection.shape}"
    assert self.T_camera_world.shape == (4, 4), f"Expected shape (4, 4), got {self.T_camera_world.shape}"
    assert len(self.image_size) == 2
    assert self.near_plane > 0
    assert self.far_plane > self.near_plane

  @property
  def depth_range(self):
    return (self.near_plane, self.far_plane)

  @property
  def device(self):
    return self.projection.device

  @property
  def dtype(self):
    return self.projection.dtype

  @property
  def focal_length(self):
    return self.projection[0:2]

  @property
  def principal_point(self):
    return self.projection[2:4]

  @property
  def T_image_camera(self):
    fx, fy, cx, cy = self.projection
    return torch.tensor([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], device=self.device, dtype=self.dtype)

  @property
  def T_image_world(self):
    K = torch.eye(4, device=self.device, dtype=self.dtype)
    K[0:3, 0:3] = self.T_image_camera
    return K @ self.T_camera_world

  @property
  def camera_position(self):
    """World position of the camera centre: ``inverse(T_camera_world)[:3, 3]`` (params.py:75-78 of the reference).
    T_camera_world is an affine view matrix [A | t; 0 0 0 1] (the projection only ever reads its first three rows,
    perspective/projection.py:212), so the position is -A^-1 t; A^-1 is formed from cross products.  Unlike
    ``torch.inverse`` on a CUDA tensor this launches no LU factorisation and does not synchronise the host
    (the info check of linalg.inv costs a device round trip per frame), and it stays differentiable."""
    T = self.T_camera_world
    if T.is_cuda and T.dtype in (torch.float32, torch.float64) and not (T.requires_grad and torch.is_grad_enabled()):
      # one single-thread kernel (gs_camera_position) instead of ten ATen launches of the same closed form
      import ctypes
      from .. import _native as N
      Tc = T.detach().contiguous()
      out = torch.empty((3,), dtype=T.dtype, device=T.device)
      N.call("gs_camera_position", ctypes.c_int32(N.dtype_code(T.dtype)), N.ptr(Tc), N.ptr(out), N.stream_ptr(T.device))
      return out
    A, t = self.T_camera_world[0:3, 0:3], self.T_camera_world[0:3, 3]
    c0 = torch.linalg.cross(A[1], A[2])
    c1 = torch.linalg.cross(A[2], A[0])
    c2 = torch.linalg.cross(A[0], A[1])
    det = torch.dot(A[0], c0)
    return -(torch.stack([c0, c1, c2], dim=1) @ t) / det

  def transformed(self, t: torch.Tensor) -> 'CameraParams':
    return replace(self, T_camera_world=t @ self.T_camera_world)

  def scale_image(self, scale: float):
    w, h = self.image_size
    return replace(self, image_size=(int(w * scale), int(h * scale)), projection=self.projection * scale)

  def requires_grad_(self, requires_grad: bool):
    self.projection.requires_grad_(requires_grad)
    self.T_camera_world.requires_grad_(requires_grad)
    return self

  def detach(self):
    return replace(self, projection=self.projection.detach(), T_camera_world=self.T_camera_world.detach())

  def to(self, device=None, dtype=None):
    return CameraParams(
      projection=self.projection.to(device=device, dtype=dtype),
      T_camera_world=self.T_camera_world.to(device=device, dtype=dtype),
      near_plane=self.near_plane, far_plane=self.far_plane, image_size=self.image_size)

  def __repr__(self):
    w, h = self.image_size
    fx, fy, cx, cy = [float(x) for x in self.projection.detach().cpu()]
    pos = ", ".join(f"{float(x):.3f}" for x in self.camera_position.detach().cpu())
    return (f"CameraParams({w}x{h}, fx={fx:.4f}, fy={fy:.4f}, cx={cx:.4f}, cy={cy:.4f}, "
            f"clipping={self.near_plane:.4f}-{self.far_plane:.4f}, position=({pos}))")
