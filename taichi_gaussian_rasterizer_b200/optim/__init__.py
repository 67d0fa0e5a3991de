"""Visibility-weighted sparse optimizers and ParameterClass (surface of taichi_splatting/optim/__init__.py, minus
restore_grad, which only serves Taichi autodiff).  Each parameter group's update over the visible
points is one fused CUDA kernel (csrc/optim_kernels.cu) instead of a Taichi kernel plus a dozen torch ops."""
from .fractional import FractionalAdam, FractionalLaProp, FractionalOpt, SparseAdam, SparseLaProp
from .parameter_class import ParameterClass
from .visibility_aware import VisibilityAwareAdam, VisibilityAwareLaProp, VisibilityOptimizer

__all__ = ["FractionalOpt", "FractionalAdam", "FractionalLaProp", "SparseAdam", "SparseLaProp",
           "VisibilityOptimizer", "VisibilityAwareAdam", "VisibilityAwareLaProp", "ParameterClass"]
