"""Perspective projection of 3D gaussians to packed 2D gaussians (EWA splatting).

Operator surface of taichi_splatting/perspective/projection.py: ``apply`` (:190-215) and
``project_to_image`` (:218-248) with the same arguments, outputs and autograd conventions.
Underneath, one sm_100a kernel projects, culls and compacts in a single pass
(csrc/geom_kernels.cu, replacing project_kernel + torch.nonzero + two gathers, :31-80 and
:146-149) and a hand-derived backward kernel (csrc/point_kernels.cu) replaces the Taichi
autodiff of indexed_project_kernel (:83-118, :164-185).  The camera is read by the kernels
directly; it is not expanded per point as in :212-213.
"""
import ctypes
from numbers import Integral

import torch
from beartype import beartype
from beartype.typing import Tuple

from .. import _native as N
from ..grad_sinks import grad_sink
from ..data_types import Gaussians3D, RasterConfig
from .params import CameraParams


def _pinned_count(device) -> torch.Tensor:
  """A pinned int32 word for the asynchronous read-back of the visible count: the next slot of a per-device ring
  (_native.PinnedWords), so overlapping read-backs on one stream never share a word."""
  return N.pinned_words.take(device)


class _ProjectFunction(torch.autograd.Function):

  @staticmethod
  def forward(ctx, position, log_scaling, rotation, alpha_logit, T_camera_world, projection,
              image_size, depth_range, blur_cov, clamp_margin, alpha_threshold, after_launch=None):
    dtype, device = position.dtype, position.device
    n = position.shape[0]
    params = N.GsProjectParams(N.dtype_code(dtype), int(image_size[0]), int(image_size[1]), 0, n,
                               float(depth_range[0]), float(depth_range[1]), float(blur_cov),
                               float(clamp_margin), float(alpha_threshold))
    points = torch.empty((n, 7), dtype=dtype, device=device)
    depth = torch.empty((n, 1), dtype=dtype, device=device)
    indexes = torch.empty((n,), dtype=torch.int64, device=device)
    count = torch.zeros((1,), dtype=torch.int32, device=device)
    lib = N.lib()
    ws = N.workspace(lib.gs_project_fwd_workspace_bytes(ctypes.byref(params)), device)
    N.call("gs_project_fwd", ctypes.byref(params), N.ptr(position), N.ptr(log_scaling), N.ptr(rotation),
                               N.ptr(alpha_logit), N.ptr(T_camera_world), N.ptr(projection), N.ptr(points),
                               N.ptr(depth), N.ptr(indexes), N.ptr(count), N.ptr(ws),
                               ctypes.c_size_t(ws.numel()), N.stream_ptr(device))
    if after_launch is None:
      v = int(count.item())  # the one host read-back: the number of gaussians in view
    else:
      # The copy of the count is enqueued FIRST (pinned buffer + event), then the caller's work that only needs the
      # device-side visible set (render_gaussians: the SH evaluation); the host waits on the event, not on the
      # stream, so that work runs on the GPU while the host is blocked and while it prepares the next launches.
      host_count = _pinned_count(device)
      host_count.copy_(count, non_blocking=True)
      ready = torch.cuda.Event()
      ready.record(torch.cuda.current_stream(device))
      after_launch(indexes, count, depth, points)
      ready.synchronize()
      v = int(host_count.item())
    points, depth, indexes = points[:v], depth[:v], indexes[:v]

    ctx.params = params
    ctx.mark_non_differentiable(indexes)
    ctx.set_materialize_grads(False)   # no zero fills for outputs nobody differentiated (depth, indexes)
    ctx.save_for_backward(position, log_scaling, rotation, alpha_logit, T_camera_world, projection, indexes)
    return points, depth, indexes

  @staticmethod
  def backward(ctx, dpoints, ddepth, dindexes):
    position, log_scaling, rotation, alpha_logit, T_camera_world, projection, indexes = ctx.saved_tensors
    need = ctx.needs_input_grad
    v = indexes.shape[0]
    if dpoints is None:   # only the depth was differentiated
      dpoints = torch.zeros((v, 7), dtype=position.dtype, device=position.device)
    inputs = (position, log_scaling, rotation, alpha_logit, T_camera_world, projection)
    # fused accumulation (grad_sinks.py): when every per gaussian input that needs a gradient has a sink, the kernel
    # adds the visible rows into the sinks and autograd gets no gradient for them
    sinks = [grad_sink(t) if need[i] else None for i, t in enumerate(inputs[:4])]
    fused = any(need[:4]) and all(s is not None for s, nd in zip(sinks, need[:4]) if nd)
    if fused:
      targets = sinks + [torch.empty_like(t) if need[4 + i] else None for i, t in enumerate(inputs[4:])]
    else:
      targets = [torch.empty_like(t) if need[i] else None for i, t in enumerate(inputs)]
    params = ctx.params
    params.accumulate_grads = int(fused)
    N.call("gs_project_bwd", ctypes.byref(params), ctypes.c_int64(v), N.ptr(position), N.ptr(log_scaling), N.ptr(rotation),
      N.ptr(alpha_logit), N.ptr(T_camera_world), N.ptr(projection), N.ptr(indexes),
      N.ptr(dpoints.contiguous()), N.ptr(None if ddepth is None else ddepth.contiguous()), *[N.ptr(g) for g in targets],
      N.stream_ptr(position.device))
    params.accumulate_grads = 0
    grads = ([None] * 4 + targets[4:]) if fused else targets
    return (*grads, None, None, None, None, None, None)


class _ProjectStaticFunction(torch.autograd.Function):
  """The projection with NOTHING read back (extension; render_gaussians(..., overlap_capacity=)): the outputs keep
  their capacity of N rows, the first ``count[0]`` of them valid, and the visible count stays on the device as a
  fourth output.  The backward hands the capacity-sized gradients to gs_project_bwd_counted.  Rows past the count are
  uninitialised in the outputs and ignored in the gradients."""

  @staticmethod
  def forward(ctx, position, log_scaling, rotation, alpha_logit, T_camera_world, projection,
              image_size, depth_range, blur_cov, clamp_margin, alpha_threshold):
    dtype, device = position.dtype, position.device
    n = position.shape[0]
    params = N.GsProjectParams(N.dtype_code(dtype), int(image_size[0]), int(image_size[1]), 0, n,
                               float(depth_range[0]), float(depth_range[1]), float(blur_cov),
                               float(clamp_margin), float(alpha_threshold))
    points = torch.empty((n, 7), dtype=dtype, device=device)
    depth = torch.empty((n, 1), dtype=dtype, device=device)
    indexes = torch.empty((n,), dtype=torch.int64, device=device)
    count = torch.zeros((1,), dtype=torch.int32, device=device)
    lib = N.lib()
    ws = N.workspace(lib.gs_project_fwd_workspace_bytes(ctypes.byref(params)), device)
    N.call("gs_project_fwd", ctypes.byref(params), N.ptr(position), N.ptr(log_scaling), N.ptr(rotation),
           N.ptr(alpha_logit), N.ptr(T_camera_world), N.ptr(projection), N.ptr(points), N.ptr(depth), N.ptr(indexes),
           N.ptr(count), N.ptr(ws), ctypes.c_size_t(ws.numel()), N.stream_ptr(device))
    ctx.params = params
    ctx.mark_non_differentiable(indexes, count)
    ctx.set_materialize_grads(False)
    ctx.save_for_backward(position, log_scaling, rotation, alpha_logit, T_camera_world, projection, indexes, count)
    return points, depth, indexes, count

  @staticmethod
  def backward(ctx, dpoints, ddepth, dindexes, dcount):
    position, log_scaling, rotation, alpha_logit, T_camera_world, projection, indexes, count = ctx.saved_tensors
    need = ctx.needs_input_grad
    cap = indexes.shape[0]
    if dpoints is None:   # only the depth was differentiated
      dpoints = torch.zeros((cap, 7), dtype=position.dtype, device=position.device)
    inputs = (position, log_scaling, rotation, alpha_logit, T_camera_world, projection)
    sinks = [grad_sink(t) if need[i] else None for i, t in enumerate(inputs[:4])]
    fused = any(need[:4]) and all(s is not None for s, nd in zip(sinks, need[:4]) if nd)
    if fused:
      targets = sinks + [torch.empty_like(t) if need[4 + i] else None for i, t in enumerate(inputs[4:])]
    else:
      targets = [torch.empty_like(t) if need[i] else None for i, t in enumerate(inputs)]
    params = ctx.params
    params.accumulate_grads = int(fused)
    N.call("gs_project_bwd_counted", ctypes.byref(params), ctypes.c_int64(cap), N.ptr(count), N.ptr(position),
           N.ptr(log_scaling), N.ptr(rotation), N.ptr(alpha_logit), N.ptr(T_camera_world), N.ptr(projection),
           N.ptr(indexes), N.ptr(dpoints.contiguous()), N.ptr(None if ddepth is None else ddepth.contiguous()),
           *[N.ptr(g) for g in targets], N.stream_ptr(position.device))
    params.accumulate_grads = 0
    grads = ([None] * 4 + targets[4:]) if fused else targets
    return (*grads, None, None, None, None, None)


def project_to_image_static(gaussians: Gaussians3D, camera_params: CameraParams, config: RasterConfig):
  """project_to_image without the host read-back of the visible count (extension): returns
  (points (N, 7), depths (N, 1), indexes (N,), count (1,) int32 on the device); rows past ``count`` are
  uninitialised.  Everything downstream takes the count from the device (the ``*_counted`` entry points), so a whole
  view can be captured in a CUDA graph."""
  position, log_scaling, rotation, alpha_logit = gaussians.shape_tensors()
  dtype = position.dtype
  N.require_cuda(position, log_scaling, rotation, alpha_logit, camera_params.T_camera_world, camera_params.projection)
  return _ProjectStaticFunction.apply(
    position.contiguous(), log_scaling.contiguous(), rotation.contiguous(), alpha_logit.contiguous(),
    camera_params.T_camera_world.to(dtype).contiguous(), camera_params.projection.to(dtype).contiguous(),
    camera_params.image_size, camera_params.depth_range, config.blur_cov, config.clamp_margin, config.alpha_threshold)


@beartype
def apply(position: torch.Tensor, log_scaling: torch.Tensor,
          rotation: torch.Tensor, alpha_logit: torch.Tensor,
          T_camera_world: torch.Tensor,
          projection: torch.Tensor,

          image_size: Tuple[Integral, Integral],
          depth_range: Tuple[float, float],

          blur_cov: float = 0.0,
          clamp_margin: float = 0.15,
          alpha_threshold: float = 1. / 255.,
          after_launch=None
          ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
  """``after_launch(indexes_capacity, count_device, depth_capacity, points_capacity)`` (extension): called after the projection kernel is enqueued and
  before the visible count is read back; see render_gaussians."""
  dtype = position.dtype
  N.require_cuda(position, log_scaling, rotation, alpha_logit, T_camera_world, projection)
  return _ProjectFunction.apply(
    position.contiguous(), log_scaling.contiguous(), rotation.contiguous(), alpha_logit.contiguous(),
    T_camera_world.to(dtype).contiguous(), projection.to(dtype).contiguous(),
    image_size, depth_range, blur_cov, clamp_margin, alpha_threshold, after_launch)


@beartype
def project_to_image(gaussians: Gaussians3D, camera_params: CameraParams, config: RasterConfig, after_launch=None
                     ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
  """
  Project 3D gaussians to 2D gaussians in image space using perspective projection
  (EWA approximation of the projected covariance, Zwicker et al. 2003).

  Returns:
    points:    torch.Tensor (V, 7)  - packed 2D gaussians in image space (mean, axis, sigma, alpha)
    depths:    torch.Tensor (V, 1)  - camera space depth
    indexes:   torch.Tensor (V,)    - indexes of the gaussians that are in view (int64, ascending)
  """
  return apply(
    *gaussians.shape_tensors(),
    camera_params.T_camera_world,
    camera_params.projection,
    camera_params.image_size,
    camera_params.depth_range,
    config.blur_cov,
    config.clamp_margin,
    config.alpha_threshold,
    after_launch)
