"""Gradient sinks: parameter storage -> buffer that a backward kernel ADDS into, instead of returning a fresh dense
gradient for autograd to accumulate (one read-modify-write pass less per parameter and view, and no dense zero fill
of the culled rows).  Installed by ``distributed.GradientBucket.fused_accumulation()`` around a multi-view batch;
used by the spherical-harmonics backward (feature) and the projection backward (position, log_scaling, rotation,
alpha_logit).  Without a registered sink every operator returns ordinary gradients."""
from typing import Optional

import torch

_sinks = {}


def register_grad_sink(param: torch.Tensor, buffer: torch.Tensor):
  assert buffer.shape == param.shape and buffer.dtype == param.dtype and buffer.is_contiguous()
  _sinks[param.data_ptr()] = buffer


def unregister_grad_sink(param: torch.Tensor):
  _sinks.pop(param.data_ptr(), None)


def grad_sink(param: torch.Tensor) -> Optional[torch.Tensor]:
  return _sinks.get(param.data_ptr())
