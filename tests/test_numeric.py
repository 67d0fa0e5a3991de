"""The shared numeric contract (include/gs_numeric.h): software expf / logf accuracy."""
import math
import struct

import numpy as np

import oracle


def _ulp(x: float) -> float:
  x = abs(float(np.float32(x)))
  if x == 0:
    return 2.0 ** -149
  return 2.0 ** (math.floor(math.log2(x)) - 23)


def test_expf_accuracy():
  rng = np.random.default_rng(0)
  xs = np.concatenate([rng.uniform(-87, 88, 20000), rng.uniform(-1, 1, 5000), [0.0, -0.0, 1.0, -1.0, 88.5]])
  worst = 0.0
  for x in xs.astype(np.float32):
    want = math.exp(float(x))
    got = oracle.expf(float(x))
    worst = max(worst, abs(got - want) / _ulp(want))
  assert worst < 1.0, f"gs_expf max error {worst} ulp"


def test_logf_accuracy():
  rng = np.random.default_rng(1)
  xs = np.concatenate([np.exp(rng.uniform(-20, 20, 20000)), rng.uniform(1e-3, 3, 5000), [1.0, 255.0, 0.5]])
  worst = 0.0
  for x in xs.astype(np.float32):
    want = math.log(float(x))
    got = oracle.logf(float(x))
    worst = max(worst, abs(got - want) / max(_ulp(want), 2.0 ** -149))
  assert worst < 1.0, f"gs_logf max error {worst} ulp"


def test_special_values():
  assert oracle.expf(0.0) == 1.0
  assert oracle.expf(-200.0) == 0.0
  assert math.isinf(oracle.expf(100.0))
  assert oracle.logf(1.0) == 0.0
  assert oracle.logf(0.0) == -math.inf
  assert math.isnan(oracle.logf(-1.0))
