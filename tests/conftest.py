import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(REPO, "tests", "golden")
if REPO not in sys.path:
  sys.path.insert(0, REPO)


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: needs a real B200 (run with `-m gpu` under gpurun)")


def pytest_collection_modifyitems(config, items):
  """GPU tests never run silently on a CPU box: without CUDA they are skipped with a reason."""
  import torch
  if torch.cuda.is_available():
    return
  skip = pytest.mark.skip(reason="no CUDA device (derl_b200 has no CPU fallback)")
  for item in items:
    if "gpu" in item.keywords:
      item.add_marker(skip)


class Golden:
  """tests/golden/<name>.npz with `case(i)` access to the c<i>_* groups."""

  def __init__(self, name):
    with np.load(os.path.join(GOLDEN, name), allow_pickle=False) as f:
      self.data = {k: f[k] for k in f.files}

  def __getitem__(self, key):
    return self.data[key]

  @property
  def ncases(self):
    return int(self.data["ncases"])

  def case(self, i):
    prefix = f"c{i}_"
    return {k[len(prefix):]: v for k, v in self.data.items() if k.startswith(prefix)}


@pytest.fixture(scope="session")
def golden():
  cache = {}

  def load(name):
    if name not in cache:
      cache[name] = Golden(name)
    return cache[name]
  return load


def reference_root():
  for cand in (os.environ.get("DERL_REF"), "/root/reference"):
    if cand and os.path.isdir(os.path.join(cand, "derl")):
      return cand
  return None


@pytest.fixture(scope="session")
def ref():
  """The live reference package (dev container only; absent on the GPU box)."""
  root = reference_root()
  if root is None:
    pytest.skip("reference tree not present (expected on the GPU box)")
  for p in (os.path.join(REPO, "tests", "_stubs"), root):
    if p not in sys.path:
      sys.path.insert(0, p)
  import derl
  import derl.summary
  derl.summary.stop_recording()
  return derl


@pytest.fixture(autouse=True)
def _quiet_summaries():
  from derl_b200 import summary
  summary.stop_recording()
  yield
  summary.stop_recording()
