"""Recording gate + writer forwarding with the contract of the reference's derl/summary.py.

Reference behaviour kept (derl/summary.py:13-61): recording starts ENABLED; every
SummaryWriter method is reachable as `summary.<method>(...)` and raises ValueError while no
writer is set (:49-51); `should_record()` is the global gate that PPOLoss / Trainer consult
before logging (derl/alg/ppo.py:56, derl/alg/common.py:61).  Host-side Python only.
"""
_state = {"record": True, "writer": None}


def should_record():
  return _state["record"]


def start_recording():
  _state["record"] = True


def stop_recording():
  _state["record"] = False


def set_recording(flag):
  _state["record"] = bool(flag)


def set_writer(summary_writer):
  _state["writer"] = summary_writer


def get_writer():
  return _state["writer"]


def make_writer(*args, **kwargs):
  from torch.utils.tensorboard import SummaryWriter
  set_writer(SummaryWriter(*args, **kwargs))


def __getattr__(name):
  """Forward `summary.add_scalar(...)` and friends to the writer (module __getattr__)."""
  if name.startswith("__"):
    raise AttributeError(name)

  def forward(*args, **kwargs):
    if _state["writer"] is None:
      raise ValueError("summary.writer cannot be None, call set_writer or "
                       "make_writer to set writer")
    return getattr(_state["writer"], name)(*args, **kwargs)
  return forward
