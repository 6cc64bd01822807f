"""A wide rollout column that starts its life in pinned host memory.

The reference re-uploads every minibatch from NumPy (`torch.from_numpy(...).to(device)`,
derl/models.py:79-88).  The resident-rollout design uploads once — but a blocking copy of a
14.8 GB frame-stack column puts ~0.3 s of PCIe time in front of the update.  `HostColumn`
removes that serial phase: during the first epoch each minibatch's rows are pulled straight
from the pinned host array by the TMA gather kernel running on a side stream (one minibatch
ahead of the consumer, i.e. concurrently with the previous minibatch's update) and written BOTH to the minibatch and to their home
position in the device-resident copy (`derl_b200_gather_rows_upload`).  One epoch's
minibatches partition the rollout, so after the first epoch the column is fully resident and
later epochs gather from HBM.  Any other access pattern falls back to one bulk upload.
"""
import numpy as np
import torch

from .. import ops  # noqa: F401

_K = torch.ops.derl_b200
MIN_BYTES = 256 << 20     # smaller columns are simply uploaded
MIN_ROW_BYTES = 2048      # TMA bulk path of the gather kernel


_SIDE_STREAMS = {}


def _side_stream(device):
  """ONE high-priority upload stream per device, shared by every HostColumn: the caching
  allocator pools freed blocks per stream, so a fresh stream per rollout would have to
  cudaMalloc its multi-GB minibatch buffers again every time (a host-blocking call that used to
  delay the first uploads of every rollout by 50-100 ms: `tools/trace_e2e.py`)."""
  index = device.index if device.index is not None else torch.cuda.current_device()
  stream = _SIDE_STREAMS.get(index)
  if stream is None:
    stream = _SIDE_STREAMS[index] = torch.cuda.Stream(torch.device("cuda", index), priority=-1)
  return stream


def eligible(array):
  """NumPy array -> pinned CPU tensor view if `array` qualifies for lazy upload, else None."""
  if not isinstance(array, np.ndarray) or array.ndim < 2 or not array.flags.c_contiguous:
    return None
  if array.dtype == object or array.dtype.kind in "USV" or array.nbytes < MIN_BYTES:
    return None
  host = torch.from_numpy(array)
  if not host.is_pinned() or host.data_ptr() % 16:
    return None
  return host


class HostColumn:
  def __init__(self, host, device, max_ctas=8, prefetch=True):
    self.host, self.device, self.max_ctas = host, torch.device(device), max_ctas
    self.prefetch, self._ahead = prefetch, None
    self.resident = None
    self.complete = False
    self._perm, self._covered = None, 0
    self._stream = None

  # ---- just enough of the array interface for the runner wrappers
  shape = property(lambda self: self.host.shape)
  ndim = property(lambda self: self.host.ndim)
  dtype = property(lambda self: self.host.dtype)

  def reshape(self, *shape):
    shape = shape[0] if len(shape) == 1 and isinstance(shape[0], (tuple, list)) else shape
    self.host = self.host.reshape(shape)
    if self.resident is not None:
      self.resident = self.resident.reshape(shape)
    return self

  def _row_ok(self):
    row_bytes = self.host[0].numel() * self.host.element_size()
    return row_bytes >= MIN_ROW_BYTES and row_bytes % 16 == 0

  def _ensure_resident(self):
    if self.resident is None:
      self.resident = torch.empty(self.host.shape, dtype=self.host.dtype, device=self.device)
    return self.resident

  def materialize(self):
    """One bulk upload (blocking semantics of an ordinary `.to(device)`)."""
    if not self.complete:
      if self._stream is not None:
        torch.cuda.current_stream(self.device).wait_stream(self._stream)
      self._ensure_resident().copy_(self.host, non_blocking=False)
      self.complete = True
    return self.resident

  def _issue(self, perm, start, count):
    """Enqueue one upload + gather on the side stream -> (rows, completion event)."""
    resident = self._ensure_resident()
    with torch.cuda.stream(self._stream):
      rows = _K.gather_rows_upload(self.host.data_ptr(), perm, start, count, resident,
                                   self.max_ctas)
      done = self._stream.record_event()
    return rows, done

  def gather(self, perm, start, count, perm_ready=None):
    """Rows perm[start:start+count] as a device tensor.  During the first epoch the NEXT range
    of the same size is enqueued on the side stream right away, so that PCIe stays busy while
    the caller is still enqueueing (or running) the update on the current range."""
    if self.complete:
      return _K.gather_rows(self.resident, perm, start, count)
    nrows = self.host.shape[0]
    sequential = (self._perm is None and start == 0) or (self._perm is perm and
                                                         start == self._covered)
    ahead, self._ahead = self._ahead, None
    if ahead is not None and ahead[0] != (start, count):
      sequential = False                     # the prefetched range is not the one asked for
    if not sequential or not self._row_ok() or perm.numel() != nrows:
      return _K.gather_rows(self.materialize(), perm, start, count)
    self._perm = perm
    main = torch.cuda.current_stream(self.device)
    if self._stream is None:
      self._ensure_resident()
      self._stream = _side_stream(self.device)
      self._stream.wait_stream(main)        # resident allocation / anything before the rollout
    if ahead is not None:
      rows, done = ahead[1], ahead[2]
    else:
      if perm_ready is not None:
        self._stream.wait_event(perm_ready)
      rows, done = self._issue(perm, start, count)
    following = start + count
    if self.prefetch and following < nrows:
      ahead_count = min(count, nrows - following)
      self._ahead = ((following, ahead_count),) + self._issue(perm, following, ahead_count)
    main.wait_event(done)
    rows.record_stream(main)
    self._covered = following
    if self._covered == nrows:
      self.complete = True                   # every row now has its home copy in HBM
    return rows
