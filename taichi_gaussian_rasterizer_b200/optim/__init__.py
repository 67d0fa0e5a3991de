"""Visibility-weighted sparse optimizers (surface of taichi_splatting/optim/__init__.py, minus ParameterClass and
restore_grad, which only serve tensordict / Taichi autodiff).  Each parameter group's update over the visible
points is one fused CUDA kernel (csrc/optim_kernels.cu) instead of a Taichi kernel plus a dozen torch ops."""
from .fractional import FractionalAdam, FractionalLaProp, FractionalOpt, SparseAdam, SparseLaProp
from .visibility_aware import VisibilityAwareAdam, VisibilityAwareLaProp, VisibilityOptimizer

__all__ = ["FractionalOpt", "FractionalAdam", "FractionalLaProp", "SparseAdam", "SparseLaProp",
           "VisibilityOptimizer", "VisibilityAwareAdam", "VisibilityAwareLaProp"]
