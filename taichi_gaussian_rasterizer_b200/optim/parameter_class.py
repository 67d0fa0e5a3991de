"""ParameterClass: rows of named tensors (one row per gaussian), some of them trainable, together with an optimizer
whose per-row state follows the rows when they are filtered or appended — what densification (split / prune)
operates on.

Public surface of taichi_splatting/optim/parameter_class.py:12-260, name for name: the constructor
``ParameterClass(tensors, parameter_groups, optimizer_state=None, optimizer=..., **optim_kwargs)``, the properties
``parameter_groups / learning_rates / tensor_state / other_state / optimizer_state / batch_size / batch_dims``, and the
methods ``set_learning_rate, update_group(s), state_dict, from_state_dict, zero_grad, step, keys, optimized_keys, items,
modify_tensors, apply, to, replace, detach, to_dict, __getitem__, append_tensors, append``.

Differences, deliberate:
  * the row container is ``taichi_gaussian_rasterizer_b200.tensor_dict.TensorDict`` (``tensordict`` is not installed);
  * the reference's ``apply`` / ``to`` call a method ``modify`` that does not exist (parameter_class.py:151-155) — here
    they work, through ``modify_tensors``;
  * every derived object (index, append, replace, modify) is produced by one private constructor path (``_derive``)
    that re-creates the optimizer of the same class with the same keyword arguments and re-seats the state.
"""
from collections.abc import Callable, Iterable, Mapping

import torch
import torch.optim as optim
from beartype import beartype
from beartype.typing import Dict, Optional, Tuple

from ..tensor_dict import TensorDict

_GROUP_INTERNAL = ('params', 'name')


def as_parameters(tensors: TensorDict | Mapping[str, torch.Tensor], keys: Iterable[str]) -> TensorDict:
  """The same rows with the tensors named in ``keys`` wrapped as leaf ``nn.Parameter``s (parameter_class.py:240-250)."""
  wanted = set(keys)
  if not isinstance(tensors, TensorDict):
    tensors = TensorDict.from_dict(tensors, batch_dims=1)
  out = TensorDict({}, batch_size=tensors.batch_size)
  for name, value in tensors.items():
    out[name] = torch.nn.Parameter(value.detach(), requires_grad=True) if name in wanted else value
  missing = wanted - set(out.keys())
  assert not missing, f"parameter groups {sorted(missing)} have no tensor (have {list(out.keys())})"
  return out


def replace_dict(d: Mapping, **kwargs) -> dict:
  merged = dict(d)
  merged.update(kwargs)
  return merged


def _split_state(state: Mapping, rows: int) -> Tuple[dict, dict]:
  """Per-row tensors (leading dimension = number of rows) versus everything else; torch's own optimizers keep their
  step counter as a 0-d tensor, which belongs with the scalars."""
  def per_row(v):
    return torch.is_tensor(v) and v.dim() >= 1 and v.shape[0] == rows
  return ({k: v for k, v in state.items() if per_row(v)}, {k: v for k, v in state.items() if not per_row(v)})


class ParameterClass:

  @beartype
  def __init__(self, tensors: TensorDict | Mapping[str, torch.Tensor],
               parameter_groups: Dict[str, Dict],
               optimizer_state: Optional[Tuple[TensorDict, Dict]] = None,
               optimizer=optim.Optimizer,
               **optim_kwargs):
    self.tensors = as_parameters(tensors, parameter_groups.keys())
    self.optim_kwargs = optim_kwargs
    self.optimizer = optimizer(
      [{'params': [self.tensors[name]], 'name': name, **options} for name, options in parameter_groups.items()],
      **optim_kwargs)
    if optimizer_state is not None:
      self._seat_state(*optimizer_state)

  def _seat_state(self, tensor_state: TensorDict, other_state: Mapping[str, Mapping]):
    for name in tensor_state.keys():
      assert name in self.tensors.keys(), f"state parameter {name} not in {list(self.tensors.keys())}"
      entry = dict(tensor_state[name].to_dict())
      # 0-d tensors (torch's step counters) are updated in place by the optimizer: the new optimizer gets its own
      entry.update({k: (v.clone() if torch.is_tensor(v) else v) for k, v in other_state.get(name, {}).items()})
      self.optimizer.state[self.tensors[name]] = entry

  def _derive(self, tensors: TensorDict, state: Optional[Tuple[TensorDict, Dict]]) -> 'ParameterClass':
    return ParameterClass(tensors, self.parameter_groups, state, optimizer=type(self.optimizer), **self.optim_kwargs)

  # ---------------------------------------------------------------------------- groups and learning rates
  def _group(self, name: str) -> dict:
    for group in self.optimizer.param_groups:
      if group['name'] == name:
        return group
    raise ValueError(f"Group {name} not found in optimizer")

  @property
  def parameter_groups(self) -> Dict[str, Dict]:
    return {g['name']: {k: v for k, v in g.items() if k not in _GROUP_INTERNAL} for g in self.optimizer.param_groups}

  @property
  def learning_rates(self) -> Dict[str, float]:
    return {g['name']: g['lr'] for g in self.optimizer.param_groups}

  @beartype
  def set_learning_rate(self, **kwargs: float):
    for name, lr in kwargs.items():
      self._group(name)['lr'] = lr
    return self

  @beartype
  def update_group(self, name: str, **kwargs):
    self._group(name).update(kwargs)

  @beartype
  def update_groups(self, **kwargs):
    for name, options in kwargs.items():
      self.update_group(name, **options)
    return {name: options['lr'] for name, options in kwargs.items()}

  # ---------------------------------------------------------------------------- optimizer state by kind
  def _states(self):
    for name, tensor in self.tensors.items():
      if tensor in self.optimizer.state:
        yield name, self.optimizer.state[tensor]

  @property
  def tensor_state(self) -> TensorDict:
    """Per-row optimizer state (moments, running visibility, ...), keyed like the tensors."""
    return TensorDict.from_dict({name: _split_state(s, self.batch_size[0])[0] for name, s in self._states()}, batch_size=self.batch_size)

  @property
  def other_state(self) -> Dict[str, Dict]:
    """The rest of the optimizer state (step counters, ...)."""
    return {name: _split_state(s, self.batch_size[0])[1] for name, s in self._states()}

  @property
  def optimizer_state(self) -> Tuple[TensorDict, Dict]:
    return self.tensor_state, self.other_state

  # ---------------------------------------------------------------------------- serialisation
  def state_dict(self) -> Dict:
    return dict(tensors=self.tensors.to_dict(),
                optimizer=(self.tensor_state.to_dict(), self.other_state),
                parameter_groups=self.parameter_groups)

  @staticmethod
  def from_state_dict(state: dict, optimizer=optim.Adam, **optim_kwargs) -> 'ParameterClass':
    per_row, scalars = state['optimizer']
    rows = TensorDict.from_dict(state['tensors'], batch_dims=1)
    return ParameterClass(rows, parameter_groups=state['parameter_groups'],
                          optimizer_state=(TensorDict.from_dict(per_row, batch_size=rows.batch_size), scalars),
                          optimizer=optimizer, **optim_kwargs)

  def __getstate__(self):
    return self.__dict__

  def __setstate__(self, state):
    self.__dict__.update(state)

  # ---------------------------------------------------------------------------- optimizer pass-through
  def zero_grad(self):
    self.optimizer.zero_grad()

  def step(self, **kwargs):
    self.optimizer.step(**kwargs)

  # ---------------------------------------------------------------------------- dictionary surface
  def keys(self):
    return self.tensors.keys()

  def optimized_keys(self):
    return self.parameter_groups.keys()

  def items(self):
    return self.tensors.items()

  def __getattr__(self, name):
    rows = self.__dict__.get('tensors')
    if rows is not None and name in rows.keys():
      return rows[name]
    raise AttributeError(f"{type(self).__name__} has no attribute or tensor {name!r}")

  @property
  def batch_size(self):
    return self.tensors.batch_size

  @property
  def batch_dims(self):
    return self.tensors.batch_dims

  def detach(self) -> TensorDict:
    return self.tensors.detach()

  def to_dict(self):
    return self.tensors.to_dict()

  # ---------------------------------------------------------------------------- derived objects
  def modify_tensors(self, f: Callable[[TensorDict], TensorDict]) -> 'ParameterClass':
    """``f`` applied to the rows and to the per-row optimizer state alike."""
    return self._derive(f(self.tensors), (f(self.tensor_state), self.other_state))

  def apply(self, f: Callable[[torch.Tensor], torch.Tensor]) -> 'ParameterClass':
    return self.modify_tensors(lambda rows: rows.apply(f))

  def to(self, device) -> 'ParameterClass':
    return self.modify_tensors(lambda rows: rows.to(device))

  def replace(self, **kwargs) -> 'ParameterClass':
    """New values for some tensors, optimizer state kept (returns a new object, like the reference)."""
    return self._derive(self.tensors.replace(**kwargs), self.optimizer_state)

  @beartype
  def __getitem__(self, idx: torch.Tensor | str):
    if isinstance(idx, str):
      return self.tensors[idx]
    return self._derive(self.tensors[idx], (self.tensor_state[idx], self.other_state))

  def append_tensors(self, tensors: TensorDict | Mapping[str, torch.Tensor],
                     tensor_state: Optional[TensorDict] = None) -> 'ParameterClass':
    """Rows appended at the end; their optimizer state is ``tensor_state`` or zeros."""
    if not isinstance(tensors, TensorDict):
      tensors = TensorDict.from_dict(tensors, batch_dims=1)
    assert set(tensors.keys()) == set(self.tensors.keys()), f"{list(tensors.keys())} != {list(self.tensors.keys())}"
    own_state = self.tensor_state
    if tensor_state is None:
      tensor_state = own_state.new_zeros(tensors.batch_size[0])
    assert tensors.shape == tensor_state.shape, f"{tensors.shape} != {tensor_state.shape}"
    rows = torch.cat([self.tensors.detach(), tensors.detach()])
    return self._derive(rows, (torch.cat([own_state, tensor_state]), self.other_state))

  def append(self, params: 'ParameterClass') -> 'ParameterClass':
    return self.append_tensors(params.tensors)
