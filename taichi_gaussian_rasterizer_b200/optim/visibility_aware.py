"""Optimizers whose per point step follows the point's visibility (optim/visibility_aware.py of the reference): the
step weight is visibility / running visibility (a power-4 mean with decay vis_beta), gradients are divided by
(visibility + vis_smooth), and the parameter moves by lr_step * saturate(weight).  ``step(indexes, visibility,
basis)`` takes the visible gaussians and their visibility as returned by the rasterizer."""
import ctypes
from typing import Optional

import torch

from .. import _native as N
from .fractional import ADAM, LAPROP, _per_point_state, fused_group_step


class VisibilityOptimizer(torch.optim.Optimizer):

  def __init__(self, algorithm: int, param_groups, lr=0.001, betas=(0.9, 0.999), eps=1e-16, vis_beta=0.9,
               vis_smooth: float = 0.01, bias_correction=True, grad_scale: float = 1.0):
    assert lr > 0, f"Invalid learning rate: {lr}"
    assert eps > 0, f"Invalid epsilon: {eps}"
    assert 0.0 <= betas[0] < 1.0, f"Invalid beta1: {betas[0]}"
    assert 0.0 <= betas[1] < 1.0, f"Invalid beta2: {betas[1]}"
    assert 0.0 <= vis_beta < 1.0, f"Invalid visibility beta: {vis_beta}"
    assert vis_smooth >= 0.0, f"Invalid visibility smoothing: {vis_smooth}"
    defaults = dict(lr=lr, betas=betas, eps=eps, mask_lr=None, point_lr=None, type="scalar",
                    bias_correction=bias_correction)
    self.algorithm = algorithm
    self.vis_beta, self.vis_smooth, self.grad_scale = vis_beta, vis_smooth, grad_scale
    super().__init__(param_groups, defaults)

  @torch.no_grad()
  def step(self, indexes: torch.Tensor, visibility: torch.Tensor, basis: Optional[torch.Tensor] = None):
    assert visibility.shape == indexes.shape, f"shape mismatch {visibility.shape} != {indexes.shape}"
    indexes = indexes.to(torch.int64).contiguous()
    visibility = visibility.to(torch.float32).contiguous()
    first = self.param_groups[0]["params"][0]
    n, device = first.shape[0], first.device
    state0 = self.state[first]
    total_weight = _per_point_state(state0, "total_weight", n, device)
    running_vis = _per_point_state(state0, "running_vis", n, device)
    weight = torch.empty_like(visibility)
    # running visibility, step weight and total weight in one pass (visibility_aware.py:37-48,88-89)
    N.call("gs_opt_update_visibility", ctypes.c_int64(indexes.shape[0]), N.ptr(indexes), N.ptr(visibility),
           N.ptr(running_vis), N.ptr(total_weight), N.ptr(weight), ctypes.c_double(self.vis_beta),
           ctypes.c_double(1e-12), N.stream_ptr(device))
    for group in self.param_groups:
      param = group["params"][0]
      assert param.shape[0] == n, f"param shape {param.shape[0]} != {n}"
      fused_group_step(self.algorithm, group, self.state[param], indexes, weight, total_weight, basis,
                       visibility=visibility, grad_scale=self.grad_scale, vis_smooth=self.vis_smooth)


class VisibilityAwareAdam(VisibilityOptimizer):
  def __init__(self, param_groups, lr=0.001, betas=(0.9, 0.999), eps=1e-16, vis_beta=0.5, vis_smooth: float = 0.01,
               bias_correction=True):
    super().__init__(ADAM, param_groups, lr=lr, betas=betas, eps=eps, vis_beta=vis_beta, vis_smooth=vis_smooth,
                     bias_correction=bias_correction)


class VisibilityAwareLaProp(VisibilityOptimizer):
  def __init__(self, param_groups, lr=0.001, betas=(0.9, 0.999), eps=1e-16, vis_beta=0.5, vis_smooth: float = 0.01,
               bias_correction=True):
    super().__init__(LAPROP, param_groups, lr=lr, betas=betas, eps=eps, vis_beta=vis_beta, vis_smooth=vis_smooth,
                     bias_correction=bias_correction)
