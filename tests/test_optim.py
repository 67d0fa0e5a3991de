"""Visibility-weighted sparse optimizers.  tests/golden/optim.npz holds four steps of every
reference optimizer class run by tests/golden/make_golden.py (the reference's own Python + its Taichi step kernels
under the emulator): parameters after each step and the final optimizer state.

  * CPU: the oracle restatement (oracle/optim_ref.py) reproduces the fixtures;
  * GPU: taichi_gaussian_rasterizer_b200.optim (one fused CUDA kernel per group) reproduces them too, with
    state dicts key- and shape-compatible with the reference's.
Tolerance: 1e-5 relative L2 per tensor (f32 pow / exp differ in the last ulps between numpy, libm and CUDA)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import optim_ref
from util import rel_l2

GOLDEN = Path(__file__).resolve().parent / "golden" / "optim.npz"
NAMES = ("position", "log_scaling", "feature", "sh")
TYPES = dict(position="local_vector", log_scaling="scalar", feature="vector", sh="scalar")
LRS = dict(position=0.1, log_scaling=0.05, feature=0.02, sh=0.01)
TOL = 1e-5


def T(a):
  return torch.from_numpy(np.ascontiguousarray(a))


def fixture():
  assert GOLDEN.exists(), f"{GOLDEN} is missing: run tests/golden/make_golden.py --only optim"
  return np.load(GOLDEN)


CLASSES = [str(c) for c in fixture()["classes"]]


def extras(d, cls, k, device="cpu"):
  out = {}
  for e in ("mask_lr", "point_lr"):
    key = f"{cls}_{e}_{k}"
    if key in d:
      out[e] = T(d[key]).to(device)
  return out


@pytest.mark.parametrize("cls", CLASSES)
def test_oracle_optimizers_match_reference(cls):
  d = fixture()
  algorithm = "adam" if "Adam" in cls else "laprop"
  groups = [dict(dict(param=T(d[f"{cls}_init_{k}"]).clone(), grad=None, type=TYPES[k], lr=LRS[k], betas=(0.9, 0.999),
                      eps=1e-16, bias_correction=True, mask_lr=None, point_lr=None, state={}), **extras(d, cls, k))
            for k in NAMES]
  for s in range(int(d["num_steps"])):
    idx, basis = T(d[f"{cls}_s{s}_indexes"]), T(d[f"{cls}_s{s}_basis"])
    for g, k in zip(groups, NAMES):
      g["grad"] = T(d[f"{cls}_s{s}_grad_{k}"])
    if cls.startswith("Visibility"):
      optim_ref.visibility_step(algorithm, groups, idx, T(d[f"{cls}_s{s}_visibility"]), basis)
    else:
      w = torch.ones(idx.shape[0]) if cls.startswith("Sparse") else T(d[f"{cls}_s{s}_weight"])
      optim_ref.fractional_step(algorithm, groups, idx, w, basis)
    for g, k in zip(groups, NAMES):
      assert rel_l2(g["param"], T(d[f"{cls}_s{s}_param_{k}"])) < TOL, f"step {s} {k}"
  for g, k in zip(groups, NAMES):
    for sk, sv in g["state"].items():
      ref = T(d[f"{cls}_state_{k}_{sk}"])
      assert sv.shape == ref.shape and rel_l2(sv, ref) < TOL, f"state {k}.{sk}"


@pytest.mark.gpu
@pytest.mark.parametrize("cls", CLASSES)
def test_gpu_optimizers_match_reference(cuda_device, cls):
  from taichi_gaussian_rasterizer_b200 import optim
  d = fixture()
  tensors = {k: torch.nn.Parameter(T(d[f"{cls}_init_{k}"]).to(cuda_device)) for k in NAMES}
  opt = getattr(optim, cls)([dict(params=[tensors[k]], name=k, lr=LRS[k], type=TYPES[k], **extras(d, cls, k, cuda_device))
                             for k in NAMES])
  for s in range(int(d["num_steps"])):
    idx = T(d[f"{cls}_s{s}_indexes"]).to(cuda_device)
    basis = T(d[f"{cls}_s{s}_basis"]).to(cuda_device)
    for k, t in tensors.items():
      t.grad = T(d[f"{cls}_s{s}_grad_{k}"]).to(cuda_device)
    before = {k: t.detach().cpu().clone() for k, t in tensors.items()}
    if cls.startswith("Visibility"):
      opt.step(indexes=idx, visibility=T(d[f"{cls}_s{s}_visibility"]).to(cuda_device), basis=basis)
    elif cls.startswith("Sparse"):
      opt.step(indexes=idx, basis=basis)
    else:
      opt.step(indexes=idx, weight=T(d[f"{cls}_s{s}_weight"]).to(cuda_device), basis=basis)
    for k, t in tensors.items():
      assert rel_l2(t, T(d[f"{cls}_s{s}_param_{k}"])) < TOL, f"step {s} {k}: {rel_l2(t, T(d[f'{cls}_s{s}_param_{k}']))}"
      # rows outside `indexes` are untouched
      mask = torch.ones(t.shape[0], dtype=torch.bool)
      mask[idx.cpu()] = False
      assert torch.equal(t.detach().cpu()[mask], before[k][mask])
  for k, t in tensors.items():     # state dict compatible with the reference's (keys, shapes, values)
    keys = {key[len(f"{cls}_state_{k}_"):] for key in d.files if key.startswith(f"{cls}_state_{k}_")}
    assert set(opt.state[t].keys()) == keys
    for sk in keys:
      ref = T(d[f"{cls}_state_{k}_{sk}"])
      assert opt.state[t][sk].shape == ref.shape and rel_l2(opt.state[t][sk], ref) < TOL, f"state {k}.{sk}"


@pytest.mark.gpu
def test_gpu_optimizer_edge_cases(cuda_device):
  from taichi_gaussian_rasterizer_b200.optim import SparseAdam, VisibilityAwareLaProp
  p = torch.nn.Parameter(torch.randn(10, 3, device=cuda_device))
  q = torch.nn.Parameter(torch.randn(10, 2, device=cuda_device))          # no gradient: skipped
  opt = VisibilityAwareLaProp([dict(params=[p], name="p", type="vector"), dict(params=[q], name="q")], lr=0.1)
  p.grad = torch.randn_like(p)
  before_p, before_q = p.detach().clone(), q.detach().clone()
  empty = torch.zeros(0, dtype=torch.int64, device=cuda_device)
  opt.step(indexes=empty, visibility=torch.zeros(0, device=cuda_device))   # nothing visible: no change
  assert torch.equal(p, before_p)
  opt.step(indexes=torch.tensor([1, 7], device=cuda_device), visibility=torch.tensor([0.5, 2.0], device=cuda_device))
  assert torch.equal(q, before_q) and not torch.equal(p[1], before_p[1]) and torch.equal(p[0], before_p[0])
  with pytest.raises(AssertionError, match="basis is required"):
    r = torch.nn.Parameter(torch.randn(4, 2, device=cuda_device))
    r.grad = torch.randn_like(r)
    SparseAdam([dict(params=[r], name="r", type="local_vector")]).step(indexes=torch.tensor([0], device=cuda_device))
  with pytest.raises(RuntimeError, match="no CPU path"):
    c = torch.nn.Parameter(torch.randn(4, 2))
    c.grad = torch.randn_like(c)
    SparseAdam([dict(params=[c], name="c")]).step(indexes=torch.tensor([0]))
