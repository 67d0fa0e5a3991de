"""A small batched dictionary of tensors: the part of ``tensordict.TensorDict`` that the reference's
``ParameterClass`` and split / prune code relies on (taichi_splatting/optim/parameter_class.py:12-260,
examples/fit_image_gaussians.py:190-215).  ``tensordict`` is not a dependency of this package.

Semantics kept from tensordict, because callers depend on them:
  * every leaf shares the leading ``batch_size`` dimensions; values may be tensors or nested TensorDicts;
  * ``td[str]`` reads a leaf, ``td[tensor | slice | mask]`` indexes every leaf along the batch dimension;
  * ``torch.cat([a, b])`` concatenates leaf by leaf (``__torch_function__``), as the reference's ``append_tensors`` does;
  * ``new_zeros(*shape)`` makes a TensorDict with the same keys whose leaves are zeros of shape ``shape + leaf.shape[batch_dims:]``;
  * ``reshape(*shape)`` reshapes the batch dimensions only;
  * ``from_dict(d, batch_dims=1)`` infers the batch size from the leaves, ``to_dict()`` returns plain nested dicts.
"""
from collections.abc import Callable, Iterable, Mapping, Sequence
from typing import Dict, Optional, Union

import torch

Leaf = Union[torch.Tensor, "TensorDict"]


def _as_size(batch_size) -> torch.Size:
  if batch_size is None:
    return torch.Size(())
  if isinstance(batch_size, int):
    return torch.Size((batch_size,))
  return torch.Size(tuple(batch_size))


class TensorDict(Mapping):
  def __init__(self, source: Optional[Mapping[str, Leaf]] = None, batch_size=None):
    self._d: Dict[str, Leaf] = {}
    self._batch_size = _as_size(batch_size)
    for k, v in (source or {}).items():
      self[k] = v

  # ------------------------------------------------------------------ construction
  @classmethod
  def from_dict(cls, d: Mapping, batch_dims: Optional[int] = None, batch_size=None) -> "TensorDict":
    """Nested dicts become nested TensorDicts; the batch size is the common leading shape of the leaves
    (its first ``batch_dims`` dimensions when given)."""
    if isinstance(d, TensorDict) and batch_dims is None and batch_size is None:
      return d
    items = {k: (cls.from_dict(v, batch_dims=batch_dims, batch_size=batch_size) if isinstance(v, Mapping) else v)
             for k, v in d.items()}
    if batch_size is None:
      shapes = [tuple(v.batch_size) if isinstance(v, TensorDict) else tuple(v.shape) for v in items.values()]
      common = []
      if shapes:
        for dims in zip(*shapes):
          if all(x == dims[0] for x in dims):
            common.append(dims[0])
          else:
            break
      if batch_dims is not None:
        assert len(common) >= batch_dims or not shapes, f"leaves do not share {batch_dims} leading dimensions: {shapes}"
        common = common[:batch_dims]
      batch_size = common
    return cls(items, batch_size=batch_size)

  # ------------------------------------------------------------------ mapping protocol
  def __setitem__(self, key: str, value: Leaf):
    assert isinstance(key, str), "only string keys can be assigned"
    if isinstance(value, Mapping) and not isinstance(value, TensorDict):
      value = TensorDict.from_dict(value, batch_size=self._batch_size)
    lead = tuple(value.batch_size) if isinstance(value, TensorDict) else tuple(value.shape)
    n = len(self._batch_size)
    assert lead[:n] == tuple(self._batch_size), f"{key}: shape {lead} does not start with batch size {tuple(self._batch_size)}"
    self._d[key] = value

  def __getitem__(self, idx):
    if isinstance(idx, str):
      return self._d[idx]
    if isinstance(idx, tuple) and idx and all(isinstance(i, str) for i in idx):
      out = self
      for k in idx:
        out = out[k]
      return out
    return self._index(idx)

  def _index(self, idx) -> "TensorDict":
    assert len(self._batch_size) >= 1, "cannot index a TensorDict without batch dimensions"
    items = {k: v[idx] for k, v in self._d.items()}
    probe = torch.empty(self._batch_size, device="meta")[idx]
    return TensorDict(items, batch_size=probe.shape)

  def __iter__(self):
    return iter(self._d)

  def __len__(self):
    return len(self._d)

  def __contains__(self, key):
    return key in self._d

  def keys(self):
    return self._d.keys()

  def values(self):
    return self._d.values()

  def items(self):
    return self._d.items()

  # ------------------------------------------------------------------ properties
  @property
  def batch_size(self) -> torch.Size:
    return self._batch_size

  @property
  def shape(self) -> torch.Size:
    return self._batch_size

  @property
  def batch_dims(self) -> int:
    return len(self._batch_size)

  @property
  def device(self):
    for v in self._d.values():
      return v.device
    return None

  # ------------------------------------------------------------------ transforms
  def apply(self, fn: Callable[[torch.Tensor], torch.Tensor], batch_size=None) -> "TensorDict":
    items = {k: (v.apply(fn, batch_size=batch_size) if isinstance(v, TensorDict) else fn(v)) for k, v in self._d.items()}
    if batch_size is None:
      return TensorDict.from_dict(items, batch_dims=self.batch_dims) if items else TensorDict({}, self._batch_size)
    return TensorDict(items, batch_size=batch_size)

  def to(self, *args, **kwargs) -> "TensorDict":
    return TensorDict({k: v.to(*args, **kwargs) for k, v in self._d.items()}, batch_size=self._batch_size)

  def detach(self) -> "TensorDict":
    return TensorDict({k: v.detach() for k, v in self._d.items()}, batch_size=self._batch_size)

  def clone(self) -> "TensorDict":
    return TensorDict({k: v.clone() for k, v in self._d.items()}, batch_size=self._batch_size)

  def replace(self, *args, **kwargs) -> "TensorDict":
    """A new TensorDict with some leaves replaced (tensordict's out-of-place ``replace``)."""
    d = dict(self._d)
    for a in args:
      d.update(a)
    d.update(kwargs)
    return TensorDict(d, batch_size=self._batch_size)

  def update(self, other: Mapping[str, Leaf]) -> "TensorDict":
    for k, v in other.items():
      self[k] = v
    return self

  def to_dict(self) -> Dict:
    return {k: (v.to_dict() if isinstance(v, TensorDict) else v) for k, v in self._d.items()}

  def new_zeros(self, *shape) -> "TensorDict":
    if len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size)):
      shape = tuple(shape[0])
    n = self.batch_dims
    items = {k: (v.new_zeros(*shape) if isinstance(v, TensorDict) else v.new_zeros(tuple(shape) + tuple(v.shape[n:])))
             for k, v in self._d.items()}
    return TensorDict(items, batch_size=shape)

  def reshape(self, *shape) -> "TensorDict":
    if len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size)):
      shape = tuple(shape[0])
    n = self.batch_dims
    items = {k: (v.reshape(*shape) if isinstance(v, TensorDict) else v.reshape(tuple(shape) + tuple(v.shape[n:])))
             for k, v in self._d.items()}
    probe = torch.empty(self._batch_size, device="meta").reshape(shape)
    return TensorDict(items, batch_size=probe.shape)

  def view(self, *shape) -> "TensorDict":
    return self.reshape(*shape)

  # ------------------------------------------------------------------ torch.cat / torch.stack on TensorDicts
  @classmethod
  def __torch_function__(cls, func, types, args=(), kwargs=None):
    kwargs = kwargs or {}
    if func is torch.cat:
      return _cat(*args, **kwargs)
    return NotImplemented

  def __repr__(self):
    inner = ", ".join(f"{k}: {tuple(v.shape)}" for k, v in self._d.items())
    return f"TensorDict({{{inner}}}, batch_size={tuple(self._batch_size)})"


def _cat(tds: Sequence[TensorDict], dim: int = 0) -> TensorDict:
  first = tds[0]
  assert all(set(t.keys()) == set(first.keys()) for t in tds), f"key mismatch in cat: {[list(t.keys()) for t in tds]}"
  assert 0 <= dim < max(first.batch_dims, 1), "TensorDicts concatenate along a batch dimension"
  items = {k: torch.cat([t[k] for t in tds], dim=dim) for k in first.keys()}
  size = list(first.batch_size)
  size[dim] = sum(t.batch_size[dim] for t in tds)
  return TensorDict(items, batch_size=size)


def cat(tds: Iterable[TensorDict], dim: int = 0) -> TensorDict:
  return _cat(list(tds), dim=dim)


__all__ = ["TensorDict", "cat"]
