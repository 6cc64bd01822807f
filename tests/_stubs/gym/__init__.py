"""Import-only stand-in for `gym` (absent in this image, no network).

Exists solely so that the read-only reference tree can be imported as a live
oracle in the dev container (tests/ref_oracle.py).  It carries no simulator:
only the class names the reference's `derl/env/*` modules subclass at import
time.  Never imported by the product package.
"""
from . import spaces  # noqa: F401


class Env:
  metadata = {}
  unwrapped = property(lambda self: self)


class Wrapper(Env):
  def __init__(self, env=None):
    self.env = env

  @property
  def unwrapped(self):
    return getattr(self.env, "unwrapped", self.env)


class ObservationWrapper(Wrapper):
  pass


class RewardWrapper(Wrapper):
  pass


class ActionWrapper(Wrapper):
  pass


Space = spaces.Space


def make(*args, **kwargs):
  raise RuntimeError("gym stub: no simulators in this image")
