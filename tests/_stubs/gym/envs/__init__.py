from . import atari  # noqa: F401
