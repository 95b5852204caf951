"""
Host-side bookkeeping of the EM loop, mirroring /root/reference/viprs/utils/OptimizeResult.py (same fields, same
update rules) so that ``fit()`` here stops on exactly the conditions the reference's ``VIPRS.fit`` does.
"""


class IterationConditionCounter:
    """Counts CONSECUTIVE iterations on which a condition held (OptimizeResult.py:2-33)."""

    def __init__(self):
        self._counter = 0
        self._nit = 0

    @property
    def counter(self):
        return self._counter

    def update(self, condition, iteration):
        self._counter = self._counter + 1 if (condition and iteration == self._nit + 1) else 0
        self._nit = iteration


class OptimizeResult:
    """Progress / outcome of one model's optimisation (OptimizeResult.py:36-153)."""

    def __init__(self):
        self.reset()
        self.stop_iteration = None
        self.success = None

    def reset(self):
        self.message = None
        self.stop_iteration = False
        self.success = False
        self.fun = None
        self.nit = 0
        self.error_on_termination = False
        self._last_drop_iter = None
        self._oscillation_counter = 0

    @property
    def iterations(self):
        return self.nit

    @property
    def objective(self):
        return self.fun

    @property
    def converged(self):
        return self.success

    @property
    def valid_optim_result(self):
        return self.success or (self.stop_iteration and not self.error_on_termination)

    @property
    def oscillation_counter(self):
        return self._oscillation_counter

    def _reset_oscillation_counter(self):
        self._oscillation_counter = 0

    def update(self, fun, stop_iteration=False, success=False, message=None, increment=True):
        if self.fun is not None and fun < self.fun:                     # a drop: maybe an oscillation (:129-133)
            if self._last_drop_iter is not None and self.nit - self._last_drop_iter == 1:
                self._oscillation_counter += 1
            self._last_drop_iter = self.nit + 1
        elif self._last_drop_iter is not None and self.nit > self._last_drop_iter:
            self._reset_oscillation_counter()
        self.fun = fun
        self.stop_iteration = stop_iteration
        self.success = success
        self.message = message
        self.nit += int(increment)
        if stop_iteration and not success and "Maximum iterations" not in (message or ""):
            self.error_on_termination = True

    def __str__(self):
        return str(self.__dict__)
