"""Compatibility shim for taichi_splatting/taichi_queue.py:34-90.

The reference funnels every Taichi launch through ``TaichiQueue`` because the Taichi
runtime is not thread-safe and must be initialised once (``TaichiQueue.init(arch=...)``
replaces ``ti.init``).  The sm_100a kernels here are stateless and stream-ordered on
``torch.cuda.current_stream()``, so nothing needs queueing: host code written against
the reference keeps working, the calls simply run inline.
"""
from concurrent.futures import Future


class TaichiQueue:
  _initialised = False
  options = {}

  @classmethod
  def init(cls, *args, threaded: bool = False, **kwargs) -> None:
    cls._initialised = True
    cls.options = dict(kwargs, threaded=threaded)

  @classmethod
  def stop(cls) -> None:
    cls._initialised = False

  @staticmethod
  def thread_id():
    return None

  @staticmethod
  def run_async(func, *args, **kwargs) -> Future:
    args = [a.result() if isinstance(a, Future) else a for a in args]
    future = Future()
    future.set_result(func(*args, **kwargs))
    return future

  @staticmethod
  def run_sync(func, *args, **kwargs):
    return TaichiQueue.run_async(func, *args, **kwargs).result()


class _QueueContext:
  def __init__(self, *args, **kwargs):
    self.args, self.kwargs = args, kwargs

  def __enter__(self):
    TaichiQueue.init(*self.args, **self.kwargs)

  def __exit__(self, exc_type, exc_value, traceback):
    TaichiQueue.stop()


def taichi_queue(*args, **kwargs):
  return _QueueContext(*args, **kwargs)


def queued(kernel):
  def f(*args, **kwargs):
    return TaichiQueue.run_sync(kernel, *args, **kwargs)
  return f
