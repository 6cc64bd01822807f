"""Import-only stand-in for `atari_py` (see tests/_stubs/gym)."""


def list_games():
  return []
