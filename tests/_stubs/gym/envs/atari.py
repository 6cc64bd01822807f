class AtariEnv:
  pass
