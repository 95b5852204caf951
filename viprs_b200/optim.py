"""
Host-side bookkeeping of the EM loop: what ``fit()`` records per iteration and what decides that it stops.

The reference keeps this in ``viprs/utils/OptimizeResult.py`` (``OptimizeResult:36-153``,
``IterationConditionCounter:2-33``).  ``fit()`` here must stop on exactly the same conditions, so the two small
state machines below reproduce the *behaviour* of those classes (same public attribute names, because
``VIPRSGrid`` and callers read ``success / stop_iteration / nit / fun / message / valid_optim_result``), written
from that behaviour rather than from the reference's code:

``Streak``       length of the current run of CONSECUTIVE iterations on which a condition held.
``FitStatus``    objective value, iteration count, stop / success flags, the termination message, and a detector
                 of objective oscillations (drops on back-to-back iterations).
"""
from dataclasses import dataclass, field
from typing import Optional


class Streak:
    """Run length of a condition over consecutive iteration numbers (reference: IterationConditionCounter)."""

    __slots__ = ("length", "last_iteration")

    def __init__(self):
        self.length = 0
        self.last_iteration = 0

    @property
    def counter(self):
        return self.length

    def update(self, condition, iteration):
        contiguous = iteration == self.last_iteration + 1
        self.length = self.length + 1 if (condition and contiguous) else 0
        self.last_iteration = iteration


@dataclass
class _DropTracker:
    """Objective drops: `pending` is the iteration index right after the latest drop; `oscillations` counts drops
    that came exactly one iteration after the previous one (OptimizeResult.py:129-138)."""
    pending: Optional[int] = None
    oscillations: int = 0

    def observe(self, dropped, nit):
        if dropped:
            if self.pending is not None and nit - self.pending == 1:
                self.oscillations += 1
            self.pending = nit + 1
        elif self.pending is not None and nit > self.pending:
            self.oscillations = 0


@dataclass
class FitStatus:
    """Progress / outcome of one model's optimisation (reference: OptimizeResult)."""
    message: Optional[str] = None
    stop_iteration: Optional[bool] = None       # None until reset(): the reference distinguishes "never started"
    success: Optional[bool] = None
    fun: Optional[float] = None
    nit: int = 0
    error_on_termination: bool = False
    _drops: _DropTracker = field(default_factory=_DropTracker, repr=False)

    def reset(self):
        self.message, self.fun, self.nit = None, None, 0
        self.stop_iteration = self.success = self.error_on_termination = False
        self._drops = _DropTracker()

    # read-only aliases the reference exposes
    iterations = property(lambda self: self.nit)
    objective = property(lambda self: self.fun)
    converged = property(lambda self: self.success)
    oscillation_counter = property(lambda self: self._drops.oscillations)

    @property
    def valid_optim_result(self):
        """Converged, or stopped without an error (e.g. ran out of iterations)."""
        return self.success or (self.stop_iteration and not self.error_on_termination)

    def update(self, fun, stop_iteration=False, success=False, message=None, increment=True):
        self._drops.observe(self.fun is not None and fun < self.fun, self.nit)
        self.fun, self.stop_iteration, self.success, self.message = fun, stop_iteration, success, message
        if increment:
            self.nit += 1
        ran_out = message is not None and "Maximum iterations" in message
        if stop_iteration and not success and not ran_out:
            self.error_on_termination = True

    def __str__(self):
        return (f"FitStatus(nit={self.nit}, fun={self.fun}, stop_iteration={self.stop_iteration}, "
                f"success={self.success}, message={self.message!r})")


# the reference's names, so that code written against viprs.utils.OptimizeResult keeps working
OptimizeResult = FitStatus
IterationConditionCounter = Streak
