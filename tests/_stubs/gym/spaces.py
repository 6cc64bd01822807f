"""Space name stubs (type checks only), see gym/__init__.py."""


class Space:
  def __init__(self, shape=None, dtype=None):
    self.shape = shape
    self.dtype = dtype


class Discrete(Space):
  def __init__(self, n):
    super().__init__((), int)
    self.n = n


class Box(Space):
  def __init__(self, low=None, high=None, shape=None, dtype=None):
    super().__init__(shape, dtype)
    self.low, self.high = low, high
