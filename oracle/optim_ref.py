"""CPU restatement of the reference's visibility-weighted sparse optimizers — TEST INFRASTRUCTURE ONLY.

Plain torch (CPU, any float dtype), one function per step, written from the reference's sources
(paths relative to /root/reference/taichi_splatting/):

  adam_kernel / laprop_kernel     optim/fractional_adam.py:7-85, optim/fractional_laprop.py:7-86
  weighted_step                   optim/fractional.py:108-148
  fractional_step                 optim/fractional.py:164-186  (FractionalOpt.step; Sparse*: weight = 1)
  visibility_step                 optim/visibility_aware.py:37-48,72-103 (VisibilityOptimizer.step)

Pinned by tests/golden/optim.npz, which tests/golden/make_golden.py produced by running the reference's own
optimizer classes (their Taichi kernels executed by the emulator).  Only tests may import this module.

A "group" here is a dict: param (N, ...) tensor, grad (same shape) or None, type, lr, betas, eps, bias_correction,
mask_lr, point_lr, and state (dict with the reference's keys).  Everything is updated in place.
"""
from typing import Optional

import torch


def lerp(t, a, b):
  return a * t + b * (1.0 - t)          # taichi_lib/generic.py lerp(t, a, b)


def saturate(x):
  return 1 - 1 / torch.exp(2 * x)       # optim/fractional.py:150-151


def moments(state, param2d, group_type):
  """optim/util.py:5-18 + fractional.py:117,120 (`m, v = get_*_state` while the getters return state['v'], state['m']):
  the FIRST moment lives under 'v' and the second under 'm' ((N, D) for scalar groups, (N,) for vector groups)."""
  if "v" not in state:
    state["v"] = torch.zeros_like(param2d)
    state["m"] = torch.zeros_like(param2d) if group_type == "scalar" else param2d.new_zeros(param2d.shape[0])
  return state["v"], state["m"]


def kernel_step(algorithm, group_type, betas, eps, bias_correction, indexes, weight, m_arr, v_arr, total_weight, grad,
                lr):
  """The per point kernels: returns lr_step (M, D) and updates m_arr / v_arr rows in place."""
  beta1, beta2 = betas
  w = weight.unsqueeze(1)
  tw = total_weight[indexes].unsqueeze(1)
  g = grad[indexes]
  m_old = m_arr[indexes]
  if group_type == "scalar":
    gg, v_old = g * g, v_arr[indexes]
  else:
    gg, v_old = (g * g).sum(dim=1, keepdim=True), v_arr[indexes].unsqueeze(1)
  v = lerp(beta2 ** w, v_old, gg)
  one = torch.ones_like(tw)
  if algorithm == "adam":
    bias = torch.sqrt(1 - beta2 ** tw) / (1 - beta1 ** tw) if bias_correction else one
    m = lerp(beta1 ** w, m_old, g)
    lr_step = m / torch.clamp_min(torch.sqrt(v), eps) * bias * lr
  else:
    bias1 = 1 - beta1 ** tw if bias_correction else one
    bias2 = 1 - beta2 ** tw if bias_correction else one
    m = lerp(beta1 ** w, m_old, g / torch.clamp_min(torch.sqrt(v / bias2), eps))
    lr_step = m * lr / bias1
  m_arr[indexes] = m
  v_arr[indexes] = v if group_type == "scalar" else v.squeeze(1)
  return lr_step


def weighted_step(algorithm, group, grad2d, indexes, weight, total_weight, basis: Optional[torch.Tensor]):
  param2d = group["param"].view(group["param"].shape[0], -1)
  gtype = group["type"]
  m_arr, v_arr = moments(group["state"], param2d, gtype)
  if gtype == "local_vector":
    assert basis is not None
    grad2d = grad2d.clone()
    grad2d[indexes] = torch.einsum("bij,bj->bi", torch.linalg.inv(basis), grad2d[indexes])
  lr_step = kernel_step(algorithm, "scalar" if gtype == "scalar" else "vector", group["betas"], group["eps"],
                        group["bias_correction"], indexes, weight, m_arr, v_arr, total_weight, grad2d, group["lr"])
  if gtype == "local_vector":
    lr_step = torch.einsum("bij,bj->bi", basis, lr_step)
  if group.get("mask_lr") is not None:
    lr_step = lr_step * group["mask_lr"].view(-1).unsqueeze(0)
  if group.get("point_lr") is not None:
    lr_step = lr_step * group["point_lr"][indexes].unsqueeze(1)
  return lr_step


def _total_weight(groups):
  st = groups[0]["state"]
  if "total_weight" not in st:
    st["total_weight"] = groups[0]["param"].new_zeros(groups[0]["param"].shape[0])
  return st["total_weight"]


@torch.no_grad()
def fractional_step(algorithm, groups, indexes, weight, basis=None):
  total_weight = _total_weight(groups)
  total_weight[indexes] += weight
  for group in groups:
    if group["grad"] is None:
      continue
    n = group["param"].shape[0]
    lr_step = weighted_step(algorithm, group, group["grad"].view(n, -1), indexes, weight, total_weight, basis)
    group["param"].view(n, -1)[indexes] -= lr_step * saturate(weight).unsqueeze(1)


@torch.no_grad()
def visibility_step(algorithm, groups, indexes, visibility, basis=None, vis_beta=0.5, vis_smooth=0.01, grad_scale=1.0,
                    eps=1e-12):
  total_weight = _total_weight(groups)
  st = groups[0]["state"]
  if "running_vis" not in st:
    st["running_vis"] = torch.zeros_like(total_weight)
  running = st["running_vis"]
  a, b = visibility ** 4, running[indexes] ** 4
  updated = (a + (b - a) * vis_beta) ** 0.25          # power_lerp(beta, visibility, running, k=4)
  running[indexes] = updated
  weight = visibility / torch.clamp_min(updated, eps)
  total_weight[indexes] += weight
  for group in groups:
    if group["grad"] is None:
      continue
    n = group["param"].shape[0]
    grad2d = torch.zeros_like(group["grad"].view(n, -1))
    grad2d[indexes] = group["grad"].view(n, -1)[indexes] * grad_scale / (visibility.unsqueeze(1) + vis_smooth)
    lr_step = weighted_step(algorithm, group, grad2d, indexes, weight, total_weight, basis)
    group["param"].view(n, -1)[indexes] -= lr_step * saturate(weight).unsqueeze(1)
