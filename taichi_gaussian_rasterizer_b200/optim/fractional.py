"""Sparse optimizers that advance each point by a fractional step `weight` (optim/fractional.py of the reference).

Usage is the reference's: every param group holds ONE tensor whose first dimension indexes the points, and carries
``name``, ``lr`` and ``type`` in {"scalar", "vector", "local_vector"}; optional ``mask_lr`` (per column factors) and
``point_lr`` (per point factors).  ``step(indexes, weight, basis)`` updates only the rows in ``indexes``.

State layout per parameter, kept key-compatible with the reference's state dicts INCLUDING its swapped names
(optim/util.py:5-18 returns ``state['v'], state['m']`` and fractional.py:117,120 unpacks that as ``m, v``): ``v`` (N, D)
holds the FIRST moment, ``m`` the second — (N, D) for scalar groups, (N,) (running squared gradient norm) for vector
groups; ``total_weight`` (N,) (and ``running_vis``) live on the first group's parameter.
"""
import ctypes
from typing import Optional

import torch

from .. import _native as N

ADAM, LAPROP = 0, 1
_GROUP_TYPES = {"scalar": 0, "vector": 1, "local_vector": 2}


def saturate(x: torch.Tensor):
  return 1 - 1 / torch.exp(2 * x)   # optim/fractional.py:150-151


def _moments(state: dict, param2d: torch.Tensor, group_type: str):
  """(first moment (N, D), second moment (N, D) | (N,)) tensors of a parameter, created on first use."""
  if "v" not in state:
    state["v"] = torch.zeros_like(param2d)
    state["m"] = torch.zeros_like(param2d) if group_type == "scalar" else \
      torch.zeros((param2d.shape[0],), dtype=param2d.dtype, device=param2d.device)
  return state["v"], state["m"]      # reference quirk: 'v' is the first moment, 'm' the second


def _per_point_state(state: dict, key: str, n: int, device):
  if key not in state:
    state[key] = torch.zeros(n, device=device, dtype=torch.float32)
  return state[key]


def fused_group_step(algorithm: int, group: dict, state: dict, indexes: torch.Tensor, weight: torch.Tensor,
                     total_weight: torch.Tensor, basis: Optional[torch.Tensor] = None,
                     visibility: Optional[torch.Tensor] = None, grad_scale: float = 1.0, vis_smooth: float = -1.0):
  """One kernel: (visibility rescale) -> (inverse basis) -> moments -> step -> (basis) -> mask_lr / point_lr ->
  ``param[indexes] -= step * saturate(weight)``.  Replaces weighted_step + the update line of the reference's
  ``step`` (optim/fractional.py:108-148,186; optim/visibility_aware.py:95-103)."""
  assert len(group["params"]) == 1, f"expected 1 tensor in group {group.get('name')}, got {len(group['params'])}"
  param = group["params"][0]
  if param.grad is None:
    return
  gtype = group["type"]
  assert gtype in _GROUP_TYPES, f"unknown group type {gtype}"
  N.require_cuda(param, param.grad, indexes, weight)
  assert param.dtype == torch.float32, "the optimizer kernels are float32 (like the reference's)"
  assert param.is_contiguous(), "parameters must be contiguous (updated in place)"
  n = param.shape[0]
  param2d = param.data.view(n, -1)
  grad2d = param.grad.contiguous().view(n, -1)
  d = param2d.shape[1]
  m, v = _moments(state, param2d, gtype)
  if gtype == "local_vector":
    assert basis is not None, "basis is required for local_vector optimizer"
    assert basis.shape == (indexes.shape[0], d, d), f"basis must be (M, {d}, {d}), got {tuple(basis.shape)}"
    basis = basis.to(torch.float32).contiguous()
  else:
    basis = None
  mask_lr, point_lr = group.get("mask_lr"), group.get("point_lr")
  if mask_lr is not None:
    mask_lr = mask_lr.to(device=param.device, dtype=torch.float32).reshape(-1).contiguous()
    assert mask_lr.shape[0] == d
  if point_lr is not None:
    point_lr = point_lr.to(device=param.device, dtype=torch.float32).contiguous()
    assert point_lr.shape == (n,)
  beta1, beta2 = group["betas"]
  p = N.GsOptParams(algorithm, _GROUP_TYPES[gtype], d, int(bool(group["bias_correction"])), n, indexes.shape[0],
                    float(group["lr"]), float(beta1), float(beta2), float(group["eps"]), float(grad_scale),
                    float(vis_smooth))
  N.call("gs_opt_step", ctypes.byref(p), N.ptr(indexes), N.ptr(weight), N.ptr(visibility), N.ptr(grad2d), N.ptr(m),
         N.ptr(v), N.ptr(total_weight), N.ptr(param2d), N.ptr(mask_lr), N.ptr(point_lr), N.ptr(basis),
         N.stream_ptr(param.device))


class FractionalOpt(torch.optim.Optimizer):

  def __init__(self, algorithm: int, param_groups, lr=0.001, betas=(0.9, 0.999), eps=1e-16, bias_correction=True):
    assert lr > 0, f"Invalid learning rate: {lr}"
    assert eps > 0, f"Invalid epsilon: {eps}"
    assert 0.0 <= betas[0] < 1.0, f"Invalid beta1: {betas[0]}"
    assert 0.0 <= betas[1] < 1.0, f"Invalid beta2: {betas[1]}"
    defaults = dict(lr=lr, betas=betas, eps=eps, mask_lr=None, point_lr=None, type="scalar",
                    bias_correction=bias_correction)
    self.algorithm = algorithm
    super().__init__(param_groups, defaults)

  def _first_state(self):
    first = self.param_groups[0]["params"][0]
    return self.state[first], first.shape[0], first.device

  @torch.no_grad()
  def step(self, indexes: torch.Tensor, weight: torch.Tensor, basis: Optional[torch.Tensor] = None):
    assert weight.shape == indexes.shape, f"shape mismatch {weight.shape} != {indexes.shape}"
    indexes = indexes.to(torch.int64).contiguous()
    weight = weight.to(torch.float32).contiguous()
    state0, n, device = self._first_state()
    total_weight = _per_point_state(state0, "total_weight", n, device)
    N.call("gs_opt_accumulate_weight", ctypes.c_int64(indexes.shape[0]), N.ptr(indexes), N.ptr(weight),
           N.ptr(total_weight), N.stream_ptr(device))
    for group in self.param_groups:
      param = group["params"][0]
      assert param.shape[0] == n, f"param shape {param.shape[0]} != {n}"
      fused_group_step(self.algorithm, group, self.state[param], indexes, weight, total_weight, basis)


class FractionalAdam(FractionalOpt):
  def __init__(self, params, lr=0.001, betas=(0.9, 0.999), eps=1e-16, bias_correction=True):
    super().__init__(ADAM, params, lr, betas, eps, bias_correction)


class FractionalLaProp(FractionalOpt):
  def __init__(self, params, lr=0.001, betas=(0.9, 0.999), eps=1e-16, bias_correction=True):
    super().__init__(LAPROP, params, lr, betas, eps, bias_correction)


class _UnitWeight:
  """Sparse* optimizers: every visible point takes a full step (weight 1)."""

  def step(self, indexes: torch.Tensor, basis: Optional[torch.Tensor] = None):
    weight = torch.ones(indexes.shape[0], device=indexes.device, dtype=torch.float32)
    FractionalOpt.step(self, indexes, weight, basis)


class SparseAdam(_UnitWeight, FractionalOpt):
  def __init__(self, params, lr=0.001, betas=(0.9, 0.999), eps=1e-16, bias_correction=True):
    FractionalOpt.__init__(self, ADAM, params, lr, betas, eps, bias_correction)


class SparseLaProp(_UnitWeight, FractionalOpt):
  def __init__(self, params, lr=0.001, betas=(0.9, 0.999), eps=1e-16, bias_correction=True):
    FractionalOpt.__init__(self, LAPROP, params, lr, betas, eps, bias_correction)
