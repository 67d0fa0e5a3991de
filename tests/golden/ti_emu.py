"""A minimal interpreter for the subset of Taichi that the reference's kernels use.  FIXTURE GENERATION ONLY.

Taichi cannot be installed in this environment (no wheel, no network), and the reference's rasterizer and
tile mapper exist only as Taichi kernel SOURCE (``/root/reference/taichi_splatting/rasterizer/forward.py``,
``.../backward.py``, ``.../mapper/tile_mapper.py``, ``.../taichi_lib/*.py``).  This module registers a fake
``taichi`` package (plus stand-ins for ``tensordict`` and the reference's CUB extension ``cuda_lib``) so that
the UNMODIFIED reference modules import and their ``@ti.kernel`` / ``@ti.func`` bodies execute as Python:

* scalars are numpy float32 / float64 scalars and Python ints, vectors / matrices are ``Vec`` (a thin wrapper
  around a numpy array with Taichi's type promotion: an f32 value combined with an int or a Python constant
  stays f32), every arithmetic step is a single IEEE operation (no FMA contraction, no BLAS);
* Taichi's value semantics are restored by rewriting each function's AST (``x = expr`` copies vectors,
  ``ti.atomic_add(a[i], v)`` becomes an in-place add on ``a``), ``ti.template()`` arguments stay by reference;
* the outermost ``for`` of a kernel is the parallel loop.  Kernels that use ``ti.simt`` run one Python thread
  per GPU thread, ``block_dim`` threads per block, with barriers behind ``block.sync``, ``sync_all_nonzero``,
  ``warp.all_nonzero / any_nonzero`` and ``warp.shfl_down_f32`` (so shared-memory staging, warp votes and the
  shuffle reductions behave as on the GPU, including the stale shared slots of SURVEY.md Q1); other kernels
  run their iterations serially in index order.

What it cannot do: Taichi's reverse-mode autodiff (``kernel.grad``), so projection / SH backward passes are
pinned through the reference's own torch implementation (torch_lib) instead; and it does not reproduce the
CUDA backend's fast-math transcendental approximations (the reference does not pin those either).

Used by tests/golden/make_golden.py, which needs /root/reference; nothing else imports this file.
"""
import ast
import inspect
import math
import struct as _struct
import sys
import textwrap
import threading
import types

import numpy as np

# ----------------------------------------------------------------------------------------------- dtypes


class DType:
  def __init__(self, name, np_type, kind, bits):
    self.name, self.np_type, self.kind, self.bits = name, np_type, kind, bits

  def __call__(self, x=0):
    if isinstance(x, Vec):
      return x.astype(self)
    if self.kind == "f":
      return self.np_type(x)
    v = int(x)
    if self.kind == "u":
      return v & ((1 << self.bits) - 1)
    v &= (1 << self.bits) - 1
    return v - (1 << self.bits) if v >= (1 << (self.bits - 1)) else v

  def __repr__(self):
    return f"ti.{self.name}"


f16 = DType("f16", np.float16, "f", 16)
f32 = DType("f32", np.float32, "f", 32)
f64 = DType("f64", np.float64, "f", 64)
i8, i16, i32, i64 = (DType(n, t, "i", b) for n, t, b in (("i8", np.int8, 8), ("i16", np.int16, 16),
                                                          ("i32", np.int32, 32), ("i64", np.int64, 64)))
u8, u16, u32, u64 = (DType(n, t, "u", b) for n, t, b in (("u8", np.uint8, 8), ("u16", np.uint16, 16),
                                                          ("u32", np.uint32, 32), ("u64", np.uint64, 64)))


def _is_float_scalar(x):
  return isinstance(x, (float, np.floating))


def _scalar(x):
  """numpy array element -> emulator scalar (np.float32/np.float64 stay, integers become Python ints)."""
  if isinstance(x, np.floating):
    return x
  if isinstance(x, (np.integer, np.bool_)):
    return int(x) if not isinstance(x, np.bool_) else bool(x)
  return x


# ----------------------------------------------------------------------------------------------- Vec


def _float_type(*ops):
  """Taichi promotion: f64 if any operand is f64, else f32 if any operand is floating (Python floats are weak)."""
  ft = None
  for o in ops:
    if isinstance(o, Vec):
      o = o.a
    if isinstance(o, np.ndarray):
      if o.dtype == np.float64:
        return np.float64
      if o.dtype.kind == "f":
        ft = np.float32
    elif isinstance(o, np.float64):
      return np.float64
    elif isinstance(o, (float, np.floating)):
      ft = ft or np.float32
  return ft


def _raw(o, ft):
  if isinstance(o, Vec):
    o = o.a
  if ft is not None:
    return np.asarray(o, dtype=ft) if isinstance(o, np.ndarray) else ft(o)
  return o


class Vec:
  """Vector / matrix value.  ``a`` is a numpy array (float32 / float64 / int64 / bool), possibly a view."""
  __slots__ = ("a",)
  __array_priority__ = 1000

  def __init__(self, a):
    self.a = a

  # -- structure
  @property
  def n(self):
    return self.a.shape[0]

  @property
  def m(self):
    return self.a.shape[1]

  def get_shape(self):
    return self.a.shape

  def __len__(self):
    return self.a.shape[0]

  def __iter__(self):
    if self.a.ndim == 1:
      return iter([_scalar(v) for v in self.a])
    return iter([Vec(self.a[i].copy()) for i in range(self.a.shape[0])])

  def copy(self):
    return Vec(self.a.copy())

  def astype(self, dtype):
    if dtype.kind == "f":
      return Vec(self.a.astype(dtype.np_type))
    return Vec(np.trunc(self.a).astype(np.int64) if self.a.dtype.kind == "f" else self.a.astype(np.int64))

  def __getitem__(self, idx):
    r = self.a[idx]
    if isinstance(r, np.ndarray):
      return Vec(r.copy())
    return _scalar(r)

  def __setitem__(self, idx, v):
    self.a[idx] = v.a if isinstance(v, Vec) else v

  def _swz(self, idxs):
    if len(idxs) == 1:
      return _scalar(self.a[idxs[0]])
    return Vec(self.a[list(idxs)].copy())

  x = property(lambda s: s._swz((0,)))
  y = property(lambda s: s._swz((1,)))
  z = property(lambda s: s._swz((2,)))
  w = property(lambda s: s._swz((3,)))
  xy = property(lambda s: s._swz((0, 1)))
  xyz = property(lambda s: s._swz((0, 1, 2)))

  # -- arithmetic: one IEEE operation per element, Taichi promotion
  def _bin(self, o, fn, rev=False):
    ft = _float_type(self, o)
    a, b = _raw(self, ft), _raw(o, ft)
    return Vec(np.asarray(fn(b, a) if rev else fn(a, b)))

  def __add__(self, o): return self._bin(o, np.add)
  def __radd__(self, o): return self._bin(o, np.add, True)
  def __sub__(self, o): return self._bin(o, np.subtract)
  def __rsub__(self, o): return self._bin(o, np.subtract, True)
  def __mul__(self, o): return self._bin(o, np.multiply)
  def __rmul__(self, o): return self._bin(o, np.multiply, True)

  def _div(self, o, rev=False):
    ft = _float_type(self, o) or np.float32   # int / int is a float division in Taichi (default fp is f32)
    a, b = _raw(self, ft), _raw(o, ft)
    return Vec(np.asarray(np.divide(b, a) if rev else np.divide(a, b)))

  def __truediv__(self, o): return self._div(o)
  def __rtruediv__(self, o): return self._div(o, True)
  def __floordiv__(self, o): return self._bin(o, np.floor_divide)
  def __rfloordiv__(self, o): return self._bin(o, np.floor_divide, True)
  def __mod__(self, o): return self._bin(o, np.mod)

  def __pow__(self, k):
    if isinstance(k, int) and k >= 1:
      r = self
      for _ in range(k - 1):
        r = r * self
      return r
    return self._bin(k, np.power)

  def __neg__(self): return Vec(-self.a)
  def __abs__(self): return Vec(np.abs(self.a))

  def _inplace(self, r):
    self.a[...] = r.a
    return self

  def __iadd__(self, o): return self._inplace(self + o)
  def __isub__(self, o): return self._inplace(self - o)
  def __imul__(self, o): return self._inplace(self * o)
  def __itruediv__(self, o): return self._inplace(self / o)

  def __gt__(self, o): return self._bin(o, np.greater)
  def __lt__(self, o): return self._bin(o, np.less)
  def __ge__(self, o): return self._bin(o, np.greater_equal)
  def __le__(self, o): return self._bin(o, np.less_equal)

  def all(self): return bool(self.a.all())
  def any(self): return bool(self.a.any())

  def sum(self):
    flat = self.a.reshape(-1)
    acc = _scalar(flat[0])
    for v in flat[1:]:
      acc = acc + _scalar(v)
    return acc

  def min(self): return _scalar(self.a.min())
  def max(self): return _scalar(self.a.max())

  def dot(self, o):
    p = self * o
    return p.sum()

  def norm(self):
    return sqrt(self.dot(self))

  def normalized(self):
    return self / self.norm()

  def transpose(self):
    return Vec(self.a.T.copy())

  def __matmul__(self, o):
    """Sequential sum of products, left to right, one rounding per operation."""
    ft = _float_type(self, o)
    a, b = _raw(self, ft), _raw(o, ft)
    if b.ndim == 1:
      out = np.zeros(a.shape[0], dtype=a.dtype)
      for i in range(a.shape[0]):
        acc = a[i, 0] * b[0]
        for k in range(1, a.shape[1]):
          acc = acc + a[i, k] * b[k]
        out[i] = acc
      return Vec(out)
    out = np.zeros((a.shape[0], b.shape[1]), dtype=a.dtype)
    for i in range(a.shape[0]):
      for j in range(b.shape[1]):
        acc = a[i, 0] * b[0, j]
        for k in range(1, a.shape[1]):
          acc = acc + a[i, k] * b[k, j]
        out[i, j] = acc
    return Vec(out)

  def __repr__(self):
    return f"Vec({self.a!r})"


class VectorType:
  def __init__(self, n, dtype):
    self.n, self.dtype = n, dtype

  def get_shape(self):
    return (self.n,)

  @property
  def np_dtype(self):
    return self.dtype.np_type if self.dtype.kind == "f" else np.int64

  def __call__(self, *args):
    if len(args) == 1 and hasattr(args[0], "detach") and hasattr(args[0], "tolist"):   # a torch tensor
      args = (args[0].detach().cpu().numpy(),)
    if len(args) == 1 and isinstance(args[0], (list, tuple, Vec, np.ndarray)):
      args = list(args[0])
    vals = []
    for a in args:
      if isinstance(a, Vec):
        vals.extend(list(a))
      else:
        vals.append(a)
    if len(vals) == 1 and self.n > 1:
      vals = vals * self.n
    assert len(vals) == self.n, f"vector({self.n}) built from {len(vals)} values"
    return Vec(np.array([self.dtype(v) for v in vals], dtype=self.np_dtype))


class MatrixType(VectorType):
  def __init__(self, n, m, dtype):
    self.n, self.m, self.dtype = n, m, dtype

  def get_shape(self):
    return (self.n, self.m)

  def __call__(self, *args):
    if len(args) == 1 and isinstance(args[0], (list, tuple)):
      rows = args[0]
      if len(rows) and isinstance(rows[0], (list, tuple, Vec)):
        vals = [v for r in rows for v in r]
      else:
        vals = list(rows)
    else:
      vals = list(args)
    if len(vals) == 1:
      vals = vals * (self.n * self.m)
    assert len(vals) == self.n * self.m
    return Vec(np.array([self.dtype(v) for v in vals], dtype=self.np_dtype).reshape(self.n, self.m))


class StructType:
  """``@ti.dataclass``: fields from the class annotations, ``@ti.func`` methods bound to instances."""

  def __init__(self, cls):
    self.__dict__["members"] = dict(getattr(cls, "__annotations__", {}))
    self.__dict__["_methods"] = {k: v for k, v in cls.__dict__.items() if callable(v)}
    self.__dict__["__name__"] = cls.__name__

  def __setattr__(self, k, v):
    self.__dict__[k] = v

  def __call__(self, *args, **kwargs):
    inst = types.SimpleNamespace()
    for name, value in list(zip(self.members, args)) + list(kwargs.items()):
      setattr(inst, name, _copy(value))
    for name, fn in self._methods.items():
      setattr(inst, name, types.MethodType(fn, inst))
    return inst


# ----------------------------------------------------------------------------------------------- math


def _unary(fn):
  def f(x):
    if isinstance(x, Vec):
      return Vec(fn(x.a))
    if type(x) in (int, float):   # a Python constant: Taichi's default floating type (np.float64 is a float subclass!)
      x = np.float32(x)
    return fn(x)
  return f


sqrt = _unary(np.sqrt)
exp = _unary(np.exp)
log = _unary(np.log)
ti_abs = _unary(np.abs)


def _round_to(fn):
  def f(x, dtype=None):
    if isinstance(x, Vec):
      r = fn(x.a)
      return Vec(r.astype(np.int64)) if dtype is not None and dtype.kind != "f" else Vec(r)
    r = fn(x)
    return int(r) if dtype is not None and dtype.kind != "f" else r
  return f


floor = _round_to(np.floor)
ceil = _round_to(np.ceil)


def _minmax(fn):
  def f(*args):
    acc = args[0]
    for b in args[1:]:
      if isinstance(acc, Vec) or isinstance(b, Vec):
        ft = _float_type(acc, b)
        acc = Vec(np.asarray(fn(_raw(acc, ft), _raw(b, ft))))
      elif _is_float_scalar(acc) or _is_float_scalar(b):
        ft = _float_type(acc, b)
        acc = fn(ft(acc), ft(b))
      else:
        acc = int(fn(acc, b))
    return acc
  return f


ti_max = _minmax(np.maximum)
ti_min = _minmax(np.minimum)


def clamp(x, lo, hi):
  return ti_min(ti_max(x, lo), hi)


def normalize(v):
  return v.normalized()


def cast(x, dtype):
  if isinstance(x, Vec):
    return x.astype(dtype)
  if dtype.kind != "f" and _is_float_scalar(x):
    return dtype(int(x))   # float -> int conversion truncates
  return dtype(x)


def bit_cast(x, dtype):
  if isinstance(x, (np.float32, float)) and dtype.bits == 32:
    return dtype(_struct.unpack("<I", _struct.pack("<f", float(x)))[0])
  raise NotImplementedError("bit_cast")


def static(x, *rest):
  return x


def template():
  return "template"


class _NdRange:
  def __init__(self, *dims):
    self.dims = [d if isinstance(d, tuple) else (0, int(d)) for d in dims]

  def __iter__(self):
    def rec(i, prefix):
      if i == len(self.dims):
        yield tuple(prefix)
        return
      lo, hi = self.dims[i]
      for v in range(lo, hi):
        yield from rec(i + 1, prefix + [v])
    return rec(0, [])

  def __len__(self):
    return int(np.prod([hi - lo for lo, hi in self.dims]))


def ndrange(*dims):
  return _NdRange(*dims)


def grouped(r):
  return (Vec(np.array(t, dtype=np.int64)) for t in r)


# ----------------------------------------------------------------------------------------------- ndarray arguments


class NdArrayAnn:
  def __init__(self, elem, ndim):
    self.elem, self.ndim = elem, ndim


def ndarray_ann(dtype=None, ndim=None, **kw):
  return NdArrayAnn(dtype, ndim)


class NdArray:
  """Kernel view of a torch tensor (shares memory through .numpy())."""

  def __init__(self, tensor, ann):
    import torch
    t = tensor.detach()
    assert t.device.type == "cpu" and t.is_contiguous(), "the emulator runs on contiguous CPU tensors"
    self.a = t.numpy() if t.dtype != torch.bfloat16 else None
    self.ndim = ann.ndim
    self.shape = tuple(self.a.shape[:ann.ndim])
    self.scalar_elem = isinstance(ann.elem, DType)

  def __iter__(self):
    """``for i in ndarray`` in a Taichi kernel ranges over the array's indices (struct-for)."""
    if self.ndim == 1:
      return iter(range(self.shape[0]))
    return iter(np.ndindex(*self.shape))

  def __getitem__(self, idx):
    r = self.a[idx]
    if isinstance(r, np.ndarray):
      return Vec(r)          # a VIEW: assignment sites copy, ti.template() arguments stay by reference
    return _scalar(r)

  def __setitem__(self, idx, v):
    self.a[idx] = v.a if isinstance(v, Vec) else v


def _convert_arg(value, ann):
  import torch
  if isinstance(ann, NdArrayAnn):
    assert isinstance(value, torch.Tensor), f"expected a tensor for an ndarray argument, got {type(value)}"
    return NdArray(value, ann)
  if isinstance(ann, VectorType) and not isinstance(value, Vec):
    return ann(value)
  if isinstance(ann, DType):
    return ann(value)
  return value


# ----------------------------------------------------------------------------------------------- SIMT


class _Block:
  def __init__(self, nthreads, first_global_tid):
    self.n = nthreads
    self.first = first_global_tid
    self.barrier = threading.Barrier(nthreads)
    self.shared = []
    self.lock = threading.Lock()
    self.votes = [0] * nthreads
    nw = (nthreads + 31) // 32
    self.warp_barrier = [threading.Barrier(min(32, nthreads - w * 32)) for w in range(nw)]
    self.warp_slots = [[None] * 32 for _ in range(nw)]
    self.error = None


_tls = threading.local()
_atomic_lock = threading.Lock()


def _ctx():
  c = getattr(_tls, "ctx", None)
  if c is None:
    raise RuntimeError("ti.simt used outside an emulated SIMT kernel")
  return c


class _SharedArray:
  def __init__(self, shape, dtype):
    shape = tuple(shape) if isinstance(shape, (tuple, list)) else (int(shape),)
    if isinstance(dtype, VectorType):
      self.a = np.zeros(shape + tuple(dtype.get_shape()), dtype=dtype.np_dtype)
    else:
      self.a = np.zeros(shape, dtype=dtype.np_type if dtype.kind == "f" else np.int64)

  def __getitem__(self, idx):
    r = self.a[idx]
    return Vec(r) if isinstance(r, np.ndarray) else _scalar(r)   # view (see NdArray.__getitem__)

  def __setitem__(self, idx, v):
    self.a[idx] = v.a if isinstance(v, Vec) else v


def shared_array(shape, dtype):
  """The k-th SharedArray call of a thread names the block's k-th shared array.  Shared memory is NOT
  re-initialised between uses: like the hardware, slots keep what the last writer left there."""
  c = _ctx()
  k = c["shared_calls"]
  c["shared_calls"] += 1
  blk = c["block"]
  with blk.lock:
    if k >= len(blk.shared):
      blk.shared.append(_SharedArray(shape, dtype))
  return blk.shared[k]


def block_sync():
  _ctx()["block"].barrier.wait()


def sync_all_nonzero(pred):
  c = _ctx()
  blk = c["block"]
  blk.votes[c["tid"]] = int(pred)
  blk.barrier.wait()
  r = int(all(v != 0 for v in blk.votes))
  blk.barrier.wait()
  return r


def sync_any_nonzero(pred):
  c = _ctx()
  blk = c["block"]
  blk.votes[c["tid"]] = int(pred)
  blk.barrier.wait()
  r = int(any(v != 0 for v in blk.votes))
  blk.barrier.wait()
  return r


def global_thread_idx():
  c = _ctx()
  return c["block"].first + c["tid"]


def _warp_exchange(value):
  c = _ctx()
  blk, tid = c["block"], c["tid"]
  w, lane = tid // 32, tid % 32
  slots = blk.warp_slots[w]
  slots[lane] = value
  blk.warp_barrier[w].wait()
  snapshot = list(slots[:blk.warp_barrier[w].parties])
  blk.warp_barrier[w].wait()
  return snapshot, lane


def warp_all_nonzero(mask, pred):
  snap, _ = _warp_exchange(int(pred))
  return int(all(v != 0 for v in snap))


def warp_any_nonzero(mask, pred):
  snap, _ = _warp_exchange(int(pred))
  return int(any(v != 0 for v in snap))


def warp_shfl_down(mask, val, offset):
  snap, lane = _warp_exchange(val)
  src = lane + int(offset)
  return snap[src] if src < len(snap) else val


def warp_shfl_up(mask, val, offset):
  snap, lane = _warp_exchange(val)
  src = lane - int(offset)
  return snap[src] if src >= 0 else val


def warp_shfl_sync(mask, val, src):
  snap, _ = _warp_exchange(val)
  return snap[int(src)]


_loop_cfg = {"block_dim": None}


def loop_config(block_dim=None, **kw):
  _loop_cfg["block_dim"] = block_dim


def _parallel_for(iterable, body, uses_simt):
  block_dim = _loop_cfg["block_dim"]
  _loop_cfg["block_dim"] = None
  items = list(iterable)

  def call(item):
    if isinstance(item, tuple):
      body(*item)
    else:
      body(item)

  if not uses_simt:
    for it in items:
      call(it)
    return
  assert block_dim, "a SIMT kernel needs ti.loop_config(block_dim=...)"
  for b0 in range(0, len(items), block_dim):
    chunk = items[b0:b0 + block_dim]
    blk = _Block(len(chunk), b0)

    def run(tid, item):
      _tls.ctx = {"block": blk, "tid": tid, "shared_calls": 0}
      try:
        call(item)
      except threading.BrokenBarrierError:
        pass
      except BaseException as e:   # noqa: BLE001 - report the first failure, release the other threads
        blk.error = blk.error or e
        blk.barrier.abort()
        for wb in blk.warp_barrier:
          wb.abort()
      finally:
        _tls.ctx = None

    threads = [threading.Thread(target=run, args=(i, it)) for i, it in enumerate(chunk)]
    for t in threads:
      t.start()
    for t in threads:
      t.join()
    if blk.error is not None:
      raise blk.error


# ----------------------------------------------------------------------------------------------- value semantics


def _copy(v):
  if isinstance(v, Vec):
    return Vec(v.a.copy())
  if isinstance(v, tuple):
    return tuple(_copy(x) for x in v)
  return v


def _atomic_add(container, idx, value):
  with _atomic_lock:
    if isinstance(container, Vec):
      old = container.a[idx]
      container.a[idx] = old + (value.a if isinstance(value, Vec) else value)
    else:
      old = container.a[idx].copy() if isinstance(container.a[idx], np.ndarray) else container.a[idx]
      container.a[idx] = container.a[idx] + (value.a if isinstance(value, Vec) else value)
  return old


class _Rewrite(ast.NodeTransformer):
  """x = expr  ->  x = __ti_copy__(expr);  ti.atomic_add(a[i], v)  ->  __ti_atomic_add__(a, i, v)."""

  def visit_Assign(self, node):
    self.generic_visit(node)
    node.value = ast.Call(func=ast.Name(id="__ti_copy__", ctx=ast.Load()), args=[node.value], keywords=[])
    return node

  def visit_Call(self, node):
    self.generic_visit(node)
    f = node.func
    if (isinstance(f, ast.Attribute) and f.attr == "atomic_add" and len(node.args) == 2
        and isinstance(node.args[0], ast.Subscript)):
      sub = node.args[0]
      return ast.Call(func=ast.Name(id="__ti_atomic_add__", ctx=ast.Load()),
                      args=[sub.value, sub.slice, node.args[1]], keywords=[])
    return node

  def visit_FunctionDef(self, node):
    for a in node.args.args + node.args.kwonlyargs:
      a.annotation = None
    node.returns = None
    node.decorator_list = []
    self.generic_visit(node)
    return node


def _closure_namespace(fn):
  ns = dict(fn.__globals__)
  if fn.__closure__:
    for name, cell in zip(fn.__code__.co_freevars, fn.__closure__):
      try:
        ns[name] = cell.cell_contents
      except ValueError:
        pass
  ns["__ti_copy__"] = _copy
  ns["__ti_atomic_add__"] = _atomic_add
  ns["__ti_parallel_for__"] = _parallel_for
  return ns


def _compile(fn, kernel):
  src = textwrap.dedent(inspect.getsource(fn))
  tree = ast.parse(src)
  fdef = tree.body[0]
  assert isinstance(fdef, ast.FunctionDef)
  fdef = _Rewrite().visit(fdef)
  if kernel:
    uses_simt = "simt" in src
    new_body = []
    for stmt in fdef.body:
      if isinstance(stmt, ast.For):
        tgt = stmt.target
        names = [tgt.id] if isinstance(tgt, ast.Name) else [e.id for e in tgt.elts]
        body_fn = ast.FunctionDef(
          name="__ti_body__", args=ast.arguments(posonlyargs=[], args=[ast.arg(arg=n) for n in names], vararg=None,
                                                kwonlyargs=[], kw_defaults=[], kwarg=None, defaults=[]),
          body=stmt.body, decorator_list=[], returns=None, type_params=[])
        new_body.append(body_fn)
        new_body.append(ast.Expr(ast.Call(func=ast.Name(id="__ti_parallel_for__", ctx=ast.Load()),
                                          args=[stmt.iter, ast.Name(id="__ti_body__", ctx=ast.Load()),
                                                ast.Constant(uses_simt)], keywords=[])))
      else:
        new_body.append(stmt)
    fdef.body = new_body
  mod = ast.Module(body=[fdef], type_ignores=[])
  ast.fix_missing_locations(mod)
  ns = _closure_namespace(fn)
  exec(compile(mod, filename=f"<ti_emu:{fn.__qualname__}>", mode="exec"), ns)
  return ns[fdef.name]


class _Func:
  def __init__(self, fn, kernel=False):
    self.fn, self.kernel, self.compiled = fn, kernel, None
    self.__name__ = fn.__name__
    self.__qualname__ = fn.__qualname__
    self.__doc__ = fn.__doc__

  def _get(self):
    if self.compiled is None:
      self.compiled = _compile(self.fn, self.kernel)
    return self.compiled

  def __get__(self, obj, objtype=None):
    return self if obj is None else types.MethodType(self, obj)

  def __call__(self, *args, **kwargs):
    f = self._get()
    if self.kernel:
      anns = self.fn.__annotations__
      names = list(inspect.signature(self.fn).parameters)
      args = [_convert_arg(a, anns.get(n)) for a, n in zip(args, names)]
    return f(*args, **kwargs)

  @property
  def grad(self):
    raise NotImplementedError("the emulator has no reverse-mode autodiff (kernel.grad)")


def func(fn):
  return _Func(fn, kernel=False)


def kernel(fn):
  return _Func(fn, kernel=True)


def dataclass(cls):
  return StructType(cls)


# ----------------------------------------------------------------------------------------------- module assembly


def _module(name, **attrs):
  m = types.ModuleType(name)
  m.__dict__.update(attrs)
  return m


def build_taichi_module():
  ti = _module("taichi")
  for d in (f16, f32, f64, i8, i16, i32, i64, u8, u16, u32, u64):
    setattr(ti, d.name, d)
  ti.int8, ti.int16, ti.int32, ti.int64 = i8, i16, i32, i64   # long spellings used by the optimizer kernels
  ti.uint8, ti.uint16, ti.uint32, ti.uint64 = u8, u16, u32, u64
  ti.float16, ti.float32, ti.float64 = f16, f32, f64
  ti.cpu, ti.cuda, ti.gpu = "cpu", "cuda", "gpu"
  ti.init = lambda *a, **k: None
  ti.reset = lambda *a, **k: None
  ti.func, ti.kernel, ti.dataclass, ti.pyfunc = func, kernel, dataclass, (lambda f: f)
  ti.static, ti.template, ti.cast, ti.bit_cast = static, template, cast, bit_cast
  ti.sqrt, ti.exp, ti.log, ti.abs = sqrt, exp, log, ti_abs
  ti.floor, ti.ceil, ti.max, ti.min = floor, ceil, ti_max, ti_min
  ti.ndrange, ti.grouped, ti.loop_config = ndrange, grouped, loop_config
  ti.atomic_add = lambda *a: (_ for _ in ()).throw(RuntimeError("atomic_add must be rewritten by the emulator"))
  ti.Vector = lambda vals, dt=None: Vec(np.array(list(vals)))

  class Matrix:
    @staticmethod
    def cols(vs):
      return Vec(np.stack([v.a for v in vs], axis=1))

    @staticmethod
    def rows(vs):
      return Vec(np.stack([v.a for v in vs], axis=0))

  ti.Matrix = Matrix

  tmath = _module("taichi.math", pi=math.pi, clamp=clamp, normalize=normalize, dot=lambda a, b: a.dot(b),
                  isinf=lambda x: bool(np.isinf(x)), isnan=lambda x: bool(np.isnan(x)))
  for n in (1, 2, 3, 4):
    if n > 1:
      setattr(tmath, f"vec{n}", VectorType(n, f32))
      setattr(tmath, f"ivec{n}", VectorType(n, i32))
      setattr(tmath, f"uvec{n}", VectorType(n, u32))
      setattr(tmath, f"mat{n}", MatrixType(n, n, f32))
  ti.math = tmath

  ttypes = _module("taichi.types",
                   vector=lambda n, dtype: VectorType(n, dtype),
                   matrix=lambda n, m, dtype: MatrixType(n, m, dtype),
                   ndarray=ndarray_ann, struct=None)
  ti.types = ttypes

  block = _module("taichi.lang.simt.block", SharedArray=shared_array, sync=block_sync,
                  sync_all_nonzero=sync_all_nonzero, sync_any_nonzero=sync_any_nonzero,
                  global_thread_idx=global_thread_idx)
  warp = _module("taichi.lang.simt.warp", all_nonzero=warp_all_nonzero, any_nonzero=warp_any_nonzero,
                 shfl_down_f32=warp_shfl_down, shfl_down_i32=warp_shfl_down, shfl_up_i32=warp_shfl_up,
                 shfl_up_f32=warp_shfl_up, shfl_sync_i32=warp_shfl_sync, shfl_sync_f32=warp_shfl_sync)
  simt = _module("taichi.lang.simt", block=block, warp=warp)
  lang_struct = _module("taichi.lang.struct", StructType=StructType)
  lang_matrix = _module("taichi.lang.matrix", VectorType=VectorType, MatrixType=MatrixType)
  lang = _module("taichi.lang", simt=simt, struct=lang_struct, matrix=lang_matrix)
  ti.lang, ti.simt = lang, simt
  return {"taichi": ti, "taichi.math": tmath, "taichi.types": ttypes, "taichi.lang": lang,
          "taichi.lang.simt": simt, "taichi.lang.simt.block": block, "taichi.lang.simt.warp": warp,
          "taichi.lang.struct": lang_struct, "taichi.lang.matrix": lang_matrix}


def _fake_tensordict():
  """Stand-in for the part of ``tensordict`` the reference uses on the host side: ``@tensorclass`` containers
  (constructed with a batch_size keyword, ``apply``) and a ``TensorDict`` with leading batch dimensions (string and
  batch indexing, nesting, ``from_dict`` / ``to_dict``, ``new_zeros``, ``torch.cat``) - what optim/parameter_class.py and
  misc/renderer2d.py need.  Written for the fixture generator only; the product package has its own
  (taichi_gaussian_rasterizer_b200/tensor_dict.py), deliberately not imported here."""
  import dataclasses
  import torch

  def tensorclass(cls):
    fields = list(getattr(cls, "__annotations__", {}))
    dc = dataclasses.dataclass(cls)
    orig_init = dc.__init__

    def __init__(self, *args, batch_size=None, **kwargs):
      orig_init(self, *args, **kwargs)
      self.batch_size = tuple(batch_size) if batch_size is not None else ()

    def apply(self, f, batch_size=None):
      return type(self)(**{k: f(getattr(self, k)) for k in fields},
                        batch_size=self.batch_size if batch_size is None else batch_size)

    dc.__init__ = __init__
    dc._fields = fields
    if not hasattr(dc, "apply"):
      dc.apply = apply
    return dc

  class TensorDict:
    def __init__(self, data, batch_size):
      self._d = dict(data)
      self.batch_size = torch.Size(batch_size)

    @classmethod
    def from_dict(cls, d, batch_dims=None, batch_size=None):
      def leaves(x):
        for v in x.values():
          if isinstance(v, (dict, TensorDict)):
            yield from leaves(v if isinstance(v, dict) else v._d)
          else:
            yield v
      if batch_size is None:
        shapes = [tuple(t.shape) for t in leaves(d)]
        common = 0
        if shapes:
          common = min(len(sh) for sh in shapes)
          for i in range(common):
            if len({sh[i] for sh in shapes}) != 1:
              common = i
              break
        nd = common if batch_dims is None else min(batch_dims, common)
        batch_size = shapes[0][:nd] if shapes else ()
      data = {k: (cls.from_dict(v, batch_size=batch_size) if isinstance(v, dict) else v) for k, v in d.items()}
      return cls(data, batch_size)

    def to_dict(self):
      return {k: (v.to_dict() if isinstance(v, TensorDict) else v) for k, v in self._d.items()}

    def _map(self, f, batch_size=None):
      return TensorDict({k: (v._map(f, batch_size) if isinstance(v, TensorDict) else f(v)) for k, v in self._d.items()},
                        self.batch_size if batch_size is None else batch_size)

    keys = lambda self: self._d.keys()        # noqa: E731
    values = lambda self: self._d.values()    # noqa: E731
    items = lambda self: self._d.items()      # noqa: E731
    __contains__ = lambda self, k: k in self._d   # noqa: E731
    __len__ = lambda self: self.batch_size[0] if len(self.batch_size) else 0   # noqa: E731

    @property
    def batch_dims(self):
      return len(self.batch_size)

    @property
    def shape(self):
      return self.batch_size

    def __getitem__(self, idx):
      if isinstance(idx, str):
        return self._d[idx]
      probe = torch.empty(self.batch_size)[idx]
      return self._map(lambda t: t[idx], batch_size=probe.shape)

    def __setitem__(self, k, v):
      self._d[k] = v

    def apply(self, f, batch_size=None):
      return self._map(f, batch_size)

    def detach(self):
      return self._map(lambda t: t.detach())

    def to(self, *a, **kw):
      return self._map(lambda t: t.to(*a, **kw))

    def replace(self, **kw):
      return TensorDict({**self._d, **kw}, self.batch_size)

    def new_zeros(self, *size):
      size = tuple(size[0]) if len(size) == 1 and isinstance(size[0], (tuple, list)) else tuple(size)
      nb = len(self.batch_size)
      return self._map(lambda t: t.new_zeros((*size, *t.shape[nb:])), batch_size=size)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
      kwargs = kwargs or {}
      if func is torch.cat:
        tds = list(args[0])
        dim = kwargs.get("dim", args[1] if len(args) > 1 else 0)

        def cat(parts):
          first = parts[0]
          out = {k: (cat([p._d[k] for p in parts]) if isinstance(first._d[k], TensorDict)
                     else torch.cat([p._d[k] for p in parts], dim=dim)) for k in first._d}
          bs = list(first.batch_size)
          bs[dim] = sum(p.batch_size[dim] for p in parts)
          return TensorDict(out, bs)
        return cat(tds)
      return NotImplemented

  return _module("tensordict", tensorclass=tensorclass, TensorDict=TensorDict)


def _fake_cuda_lib():
  """The reference's CUB wrappers (cuda_lib/full_cumsum.cu, radix_sort_pairs.cu) restated with torch / numpy on
  the CPU: exclusive sum with the total appended, and a STABLE sort on key bits [start_bit, end_bit)."""
  import torch

  def full_cumsum(x):
    if x.shape[0] == 0:
      return x.new_zeros((1,)), 0
    out = x.new_zeros((x.shape[0] + 1,))
    out[1:] = torch.cumsum(x, 0)
    return out, int(out[-1])

  def radix_sort_pairs(keys, values, start_bit=0, end_bit=None):
    k = keys.numpy()
    bits = k.dtype.itemsize * 8
    end = bits if end_bit in (None, -1) else end_bit
    mask = ((1 << end) - 1) & ~((1 << start_bit) - 1)
    order = np.argsort(k & k.dtype.type(mask), kind="stable")
    return torch.from_numpy(k[order].copy()), values[torch.from_numpy(order)]

  return _module("taichi_splatting.cuda_lib", full_cumsum=full_cumsum, radix_sort_pairs=radix_sort_pairs,
                 segmented_sort_pairs=None)


def install(reference_root="/root/reference"):
  """Register the fake modules and make ``taichi_splatting.*`` importable from the reference checkout WITHOUT
  running its package ``__init__`` (which JIT-compiles the CUB extension with nvcc at import)."""
  import importlib
  for name, mod in build_taichi_module().items():
    sys.modules[name] = mod
  sys.modules.setdefault("tensordict", _fake_tensordict())
  pkg = types.ModuleType("taichi_splatting")
  pkg.__path__ = [f"{reference_root}/taichi_splatting"]
  sys.modules["taichi_splatting"] = pkg
  sys.modules["taichi_splatting.cuda_lib"] = _fake_cuda_lib()
  pkg.cuda_lib = sys.modules["taichi_splatting.cuda_lib"]
  queue = importlib.import_module("taichi_splatting.taichi_queue")
  queue.TaichiQueue.init()
  return pkg
