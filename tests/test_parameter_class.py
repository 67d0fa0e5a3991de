"""ParameterClass (optim/parameter_class.py:12-260), the split helpers
(misc/renderer2d.py:60-132) and the split / prune policy (examples/fit_image_gaussians.py:153-228).
Host logic: runs on CPU with torch.optim.Adam; the GPU test runs the whole fit loop on the CUDA path."""
import math
import pickle

import pytest
import torch

from taichi_gaussian_rasterizer_b200.data_types import Gaussians2D, RasterConfig
from taichi_gaussian_rasterizer_b200.examples import fit_image_gaussians as fit
from taichi_gaussian_rasterizer_b200.misc import renderer2d
from taichi_gaussian_rasterizer_b200.optim.parameter_class import ParameterClass, as_parameters
from taichi_gaussian_rasterizer_b200.synthetic import random_2d_gaussians
from taichi_gaussian_rasterizer_b200.tensor_dict import TensorDict

GROUPS = dict(position=dict(lr=0.1), log_scaling=dict(lr=0.05), rotation=dict(lr=0.2), alpha_logit=dict(lr=0.1),
              feature=dict(lr=0.3))


def make_params(n=50, seed=0):
  torch.manual_seed(seed)
  g = random_2d_gaussians(n, (64, 48))
  return ParameterClass(g.to_tensordict(), GROUPS, optimizer=torch.optim.Adam, betas=(0.8, 0.9))


def adam_step(params):
  params.zero_grad()
  loss = sum((t ** 2).sum() for k, t in params.items() if k in params.optimized_keys())
  loss.backward()
  params.step()


# ----------------------------------------------------------------------------------------------- TensorDict
def test_tensordict_index_cat_zeros_reshape():
  td = TensorDict.from_dict(dict(a=torch.arange(12.).view(6, 2), b=dict(c=torch.arange(6), d=torch.ones(6, 3, 2))),
                            batch_dims=1)
  assert tuple(td.batch_size) == (6,) and tuple(td['b'].batch_size) == (6,)
  mask = torch.tensor([True, False, True, True, False, False])
  sub = td[mask]
  assert tuple(sub.batch_size) == (3,) and torch.equal(sub['b']['c'], torch.tensor([0, 2, 3]))
  assert torch.equal(td[torch.tensor([5, 0])]['a'], td['a'][[5, 0]])
  both = torch.cat([td, sub])
  assert tuple(both.batch_size) == (9,) and both['b']['d'].shape == (9, 3, 2)
  z = td.new_zeros(4, 2)
  assert tuple(z.batch_size) == (4, 2) and z['a'].shape == (4, 2, 2) and z['b']['c'].dtype == torch.int64
  r = z.reshape(8)
  assert tuple(r.batch_size) == (8,) and r['b']['d'].shape == (8, 3, 2)
  d = td.to_dict()
  assert isinstance(d['b'], dict) and torch.equal(TensorDict.from_dict(d, batch_dims=1)['b']['c'], td['b']['c'])
  with pytest.raises(AssertionError):
    td['bad'] = torch.zeros(5)


def test_gaussians_round_trip_through_tensordict():
  g = random_2d_gaussians(10, (32, 32))
  td = g.to_tensordict()
  assert isinstance(td, TensorDict) and tuple(td.batch_size) == (10,)
  g2 = Gaussians2D.from_tensordict(td[torch.arange(3)])
  assert g2.batch_size == (3,) and torch.equal(g2.position, g.position[:3])


# ----------------------------------------------------------------------------------------------- ParameterClass
def test_parameters_groups_and_learning_rates():
  p = make_params()
  assert set(p.optimized_keys()) == set(GROUPS) and 'z_depth' in p.keys() and 'z_depth' not in p.optimized_keys()
  assert isinstance(p.position, torch.nn.Parameter) and not isinstance(p.z_depth, torch.nn.Parameter)
  assert p.learning_rates['feature'] == 0.3 and p.parameter_groups['position']['lr'] == 0.1
  p.set_learning_rate(position=0.01)
  assert p.learning_rates['position'] == 0.01 and p.learning_rates['feature'] == 0.3
  assert p.update_groups(rotation=dict(lr=0.5, weight_decay=0.1)) == dict(rotation=0.5)
  assert p.parameter_groups['rotation']['weight_decay'] == 0.1
  with pytest.raises(ValueError):
    p.update_group('nope', lr=1.0)
  with pytest.raises(AttributeError):
    p.nope


def test_filter_keeps_the_optimizer_state_of_surviving_rows():
  p = make_params()
  adam_step(p)
  adam_step(p)
  state = p.tensor_state
  assert set(state.keys()) == set(GROUPS) and state['position']['exp_avg'].shape == (50, 2)
  assert float(p.other_state['position']['step']) == 2.0      # torch's 0-d step counter is not per-row state
  keep = torch.rand(50) > 0.4
  q = p[keep]
  assert q.batch_size[0] == int(keep.sum()) and type(q.optimizer) is torch.optim.Adam
  assert torch.equal(q.position.detach(), p.position.detach()[keep])
  assert torch.equal(q.tensor_state['feature']['exp_avg_sq'], state['feature']['exp_avg_sq'][keep])
  assert q.optimizer.defaults['betas'] == (0.8, 0.9) and q.learning_rates == p.learning_rates
  # the filtered object trains on: a further step matches stepping the full set and filtering afterwards
  adam_step(p)
  adam_step(q)
  assert torch.allclose(q.position.detach(), p.position.detach()[keep], atol=1e-6)


def test_append_zero_state_and_given_state():
  p = make_params(20)
  adam_step(p)
  extra = random_2d_gaussians(5, (64, 48)).to_tensordict()
  q = p.append_tensors(extra)
  assert q.batch_size[0] == 25 and torch.equal(q.feature.detach()[20:], extra['feature'])
  assert torch.equal(q.tensor_state['position']['exp_avg'][:20], p.tensor_state['position']['exp_avg'])
  assert (q.tensor_state['position']['exp_avg'][20:] == 0).all()
  given = p.tensor_state[torch.arange(5)]
  q2 = p.append_tensors(extra, given)
  assert torch.equal(q2.tensor_state['rotation']['exp_avg_sq'][20:], given['rotation']['exp_avg_sq'])
  assert p.append(p[torch.arange(3)]).batch_size[0] == 23
  with pytest.raises(AssertionError):
    p.append_tensors(TensorDict(dict(position=torch.zeros(2, 2)), batch_size=[2]))


def test_replace_modify_state_dict_pickle():
  p = make_params(12)
  adam_step(p)
  r = p.replace(rotation=torch.nn.functional.normalize(p.rotation.detach() * 3))
  assert torch.allclose(r.rotation.norm(dim=1), torch.ones(12)) and isinstance(r.rotation, torch.nn.Parameter)
  assert torch.equal(r.tensor_state['rotation']['exp_avg'], p.tensor_state['rotation']['exp_avg'])
  d = p.apply(lambda t: t.double())
  assert d.position.dtype == torch.float64 and d.tensor_state['position']['exp_avg'].dtype == torch.float64
  assert p.to(torch.device('cpu')).batch_size[0] == 12
  sd = p.state_dict()
  back = ParameterClass.from_state_dict(sd, optimizer=torch.optim.Adam, betas=(0.8, 0.9))
  assert torch.equal(back.feature.detach(), p.feature.detach())
  assert torch.equal(back.tensor_state['feature']['exp_avg'], p.tensor_state['feature']['exp_avg'])
  assert back.learning_rates == p.learning_rates
  again = pickle.loads(pickle.dumps(p))
  assert torch.equal(again.position.detach(), p.position.detach())
  det = p.detach()
  assert not det['position'].requires_grad and set(p.to_dict()) == set(p.keys())


def test_as_parameters_rejects_unknown_groups():
  with pytest.raises(AssertionError):
    as_parameters(dict(a=torch.zeros(3, 1)), ['a', 'b'])


# ----------------------------------------------------------------------------------------------- split helpers
def test_uniform_split_places_children_along_the_long_axis():
  torch.manual_seed(1)
  g = random_2d_gaussians(30, (64, 64))
  kids = renderer2d.uniform_split_gaussians2d(g, n=2, sep=0.7, depth_noise=0.0)
  assert kids.batch_size == (60,)
  axis = torch.argmax(g.log_scaling, dim=1)
  basis = renderer2d.point_basis(g)                      # columns: sigma-scaled principal axes
  along = basis[torch.arange(30), :, axis]               # (30, 2)
  pos = kids.position.view(30, 2, 2)
  assert torch.allclose(pos[:, 0], g.position - 0.7 * along, atol=1e-4)
  assert torch.allclose(pos[:, 1], g.position + 0.7 * along, atol=1e-4)
  sc = kids.scaling.view(30, 2, 2)
  shrink = math.sqrt(2) / 2
  expect = g.scaling.clone()
  expect[torch.arange(30), axis] *= shrink
  assert torch.allclose(sc[:, 0], expect, rtol=1e-5) and torch.allclose(sc[:, 1], expect, rtol=1e-5)
  assert torch.equal(kids.feature.view(30, 2, -1)[:, 1], g.feature)
  assert torch.allclose(kids.z_depth.view(30, 2), g.z_depth.expand(30, 2).clamp_min(1e-6))


def test_sampled_split_statistics():
  torch.manual_seed(2)
  g = random_2d_gaussians(4, (64, 64))
  kids = renderer2d.split_gaussians2d(g, n=4000)
  assert kids.batch_size == (16000,)
  assert torch.allclose(kids.log_scaling.view(4, 4000, 2)[:, 0], g.log_scaling + math.log(1 / math.sqrt(4000)))
  off = kids.position.view(4, 4000, 2) - g.position.unsqueeze(1)
  cov = torch.einsum('pki,pkj->pij', off, off) / 4000
  assert torch.allclose(cov, 0.25 * renderer2d.point_covariance(g), rtol=0.15, atol=0.05)
  one = renderer2d.sample_gaussians(g)
  assert one.shape == (4, 2)
  assert torch.allclose(renderer2d.point_rotation(g) @ renderer2d.point_rotation(g).transpose(1, 2),
                        torch.eye(2).expand(4, 2, 2), atol=1e-5)


# ----------------------------------------------------------------------------------------------- policy
def test_make_epochs_and_masks():
  epochs = fit.make_epochs(2000, 8, 32)
  assert sum(epochs) == 2000 and epochs[0] == 8 and max(epochs[:-1]) <= 32 and epochs == sorted(epochs[:-1]) + epochs[-1:]
  t = torch.tensor([5., 1., 4., 2., 3.])
  assert fit.take_n(t, 2).tolist() == [False, True, False, True, False]
  assert fit.take_n(t, 2, descending=True).tolist() == [True, False, True, False, False]
  assert fit.take_n(t, 0).sum() == 0
  torch.manual_seed(0)
  m = fit.randomize_n(torch.tensor([0., 1., 1., 0., 1.]), 2)
  assert m.sum() == 2 and not m[0] and not m[3]


def test_find_split_prune_reaches_the_target_and_masks_are_disjoint():
  torch.manual_seed(3)
  n, target = 100, 120
  cost, score = torch.rand(n), torch.rand(n)
  split, prune = fit.find_split_prune(n, target, 10, cost, score)
  assert not (split & prune).any()
  both = fit.take_n(cost, 10) & fit.take_n(score, 30, descending=True)
  # every split adds one point: n - prune + split == target when nothing was selected twice
  assert n - int(prune.sum()) + int(split.sum()) == target - 0 * int(both.sum())
  assert int(prune.sum()) == 10 - int(both.sum())
  split, prune = fit.find_split_prune(100, 50, 0, cost, score)       # above target, no pruning: nothing happens
  assert split.sum() == 0 and prune.sum() == 0


@pytest.mark.gpu
def test_fit_loop_with_split_prune_on_the_cuda_path(cuda_device):
  ref = fit.synthetic_target((256, 192), seed=0, device=cuda_device)
  config = RasterConfig(compute_point_heuristic=True, compute_visibility=True)
  params, history = fit.fit(ref, n=1500, target=2500, iters=60, config=config, seed=0, epoch=6, max_epoch=12)
  assert history[-1]['psnr'] > history[0]['psnr'] + 2.0, history
  assert history[-1]['n'] > 1500 and params.batch_size[0] <= 2500 + 1
  assert any(e.get('split', 0) > 0 for e in history) and any(e.get('prune', 0) > 0 for e in history)
  state = params.tensor_state
  assert state['position']['running_vis'].shape[0] == params.batch_size[0]
  for name, t in params.items():
    assert torch.isfinite(t).all(), name


# ------------------------------------------------------------------------------- fixtures made from the reference itself
# tests/golden/pclass.npz (make_golden.py --only pclass): the reference's ParameterClass driven by its own SparseAdam
# (Taichi step kernels under the emulator, tensordict stand-in of ti_emu.py) through a scripted sequence — two steps, a
# row filter, a step, an append with zero state, a step — and its split operations with a seeded generator.
import numpy as np   # noqa: E402
from pathlib import Path   # noqa: E402

PCLASS = Path(__file__).resolve().parent / "golden" / "pclass.npz"
FIELDS = ("position", "z_depth", "log_scaling", "rotation", "alpha_logit", "feature")


def _T(a):
  return torch.from_numpy(np.ascontiguousarray(a))


def _rel(a, b):
  a, b = a.detach().double().cpu(), b.double()
  return ((a - b).norm() / max(b.norm().item(), 1e-30)).item()


def test_split_operations_match_reference():
  """split_gaussians2d / uniform_split_gaussians2d / point_basis / point_covariance against the reference's functions
  run on the same inputs with the same generator seed (misc/renderer2d.py:36-132)."""
  d = np.load(PCLASS)
  g = Gaussians2D(**{k: _T(d[f"split_in_{k}"]) for k in FIELDS}, batch_size=(d["split_in_position"].shape[0],))
  cases = dict(split2=lambda: renderer2d.split_gaussians2d(g, n=2),
               split3s=lambda: renderer2d.split_gaussians2d(g, n=3, scaling=0.6),
               uniform2=lambda: renderer2d.uniform_split_gaussians2d(g, n=2),
               uniform3r=lambda: renderer2d.uniform_split_gaussians2d(g, n=3, random_axis=True, sep=0.5))
  for name, fn in cases.items():
    torch.manual_seed(11)
    res = fn()
    for k in FIELDS:
      ref = _T(d[f"{name}_{k}"])
      assert getattr(res, k).shape == ref.shape, f"{name}.{k}"
      assert torch.allclose(getattr(res, k), ref, rtol=1e-6, atol=1e-7), f"{name}.{k}: {_rel(getattr(res, k), ref)}"
  assert torch.allclose(renderer2d.point_basis(g), _T(d["point_basis"]), rtol=1e-6, atol=1e-7)
  assert torch.allclose(renderer2d.point_covariance(g), _T(d["point_covariance"]), rtol=1e-6, atol=1e-7)


@pytest.mark.gpu
def test_parameter_class_sequence_matches_reference(cuda_device):
  """Rows and optimizer state after every operation of the scripted sequence equal the reference's: construction,
  step, filter (state follows the rows), append (new rows get zero state), step on the derived objects."""
  from taichi_gaussian_rasterizer_b200.optim import SparseAdam
  d = np.load(PCLASS)
  names = ("position", "log_scaling", "feature", "z_depth")
  groups = dict(position=dict(lr=0.1, type="vector"), log_scaling=dict(lr=0.05, type="scalar"),
                feature=dict(lr=0.02, type="vector"))
  tensors = {k: _T(d[f"init_{k}"]).to(cuda_device) for k in names}
  pc = ParameterClass(TensorDict.from_dict(tensors, batch_dims=1), groups, optimizer=SparseAdam, betas=(0.9, 0.95),
                      eps=1e-12, bias_correction=True)

  def check(tag, pc):
    for k in names:
      ref = _T(d[f"{tag}_tensor_{k}"])
      assert pc.tensors[k].shape == ref.shape and _rel(pc.tensors[k], ref) < 1e-5, f"{tag} tensor {k}"
    state = pc.tensor_state.to_dict()
    keys = {key[len(f"{tag}_state_"):] for key in d.files if key.startswith(f"{tag}_state_")}
    ours = {f"{k}_{sk}" for k, st in state.items() for sk in st}
    assert ours == keys, f"{tag}: state entries {ours} != {keys}"
    for k, st in state.items():
      for sk, sv in st.items():
        ref = _T(d[f"{tag}_state_{k}_{sk}"])
        assert sv.shape == ref.shape and _rel(sv, ref) < 1e-5, f"{tag} state {k}.{sk}: {_rel(sv, ref)}"

  def step(i, pc):
    for k in groups:
      pc.tensors[k].grad = _T(d[f"s{i}_grad_{k}"]).to(cuda_device)
    pc.step(indexes=_T(d[f"s{i}_indexes"]).to(cuda_device))
    check(f"after_step{i}", pc)

  step(0, pc)
  step(1, pc)
  pc = pc[_T(d["keep"]).to(cuda_device)]
  check("after_filter", pc)
  step(2, pc)
  pc = pc.append_tensors(TensorDict.from_dict({k: _T(d[f"new_{k}"]).to(cuda_device) for k in names}, batch_dims=1))
  check("after_append", pc)
  step(3, pc)
