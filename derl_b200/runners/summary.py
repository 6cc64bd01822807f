"""PeriodicSummaries: opens the summary gate every `log_period` env steps
(reference: derl/runners/summary.py:6-30).  Host logic only."""
from .. import summary
from .env_runner import RunnerWrapper


class PeriodicSummaries(RunnerWrapper):
  def __init__(self, runner, log_period):
    super().__init__(runner)
    self.log_period = log_period
    self.last_record_step = None

  @classmethod
  def make_with_nlogs(cls, runner, nlogs=1e5):
    if runner.nsteps is None:
      raise ValueError("runner.nsteps cannot be None")
    return cls(runner, int(runner.nsteps / nlogs))

  def run(self, obs=None):
    summary.start_recording()
    self.last_record_step = self.runner.step_count
    for interactions in self.runner.run(obs):
      yield interactions
      upcoming = self.runner.step_count + 1
      due = upcoming - self.last_record_step >= self.log_period
      summary.set_recording(due)
      if due:
        self.last_record_step = upcoming
