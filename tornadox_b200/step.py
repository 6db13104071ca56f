"""Step-size control (host scalars).  Mirrors tornadox/step.py: ``ConstantSteps`` (:28-52),
``AdaptiveSteps`` (:55-108), ``propose_first_dt`` (:111-115).

The d-wide reduction of ``AdaptiveSteps.scale_error_estimate`` is fused into the step kernels
(the solvers hand the pre-reduced sum to ``norm_from_sq_sum``); the tensor version below is kept
for API parity and for user-defined solvers.
"""

import abc
import math

import torch


class StepRule(abc.ABC):
    """Step-size selection rules for ODE solvers (step.py:10-25)."""

    @abc.abstractmethod
    def suggest(self, previous_dt, scaled_error_estimate, local_convergence_rate=None):
        raise NotImplementedError

    @abc.abstractmethod
    def is_accepted(self, scaled_error_estimate):
        raise NotImplementedError

    def scale_error_estimate(self, unscaled_error_estimate, reference_state):
        raise NotImplementedError

    def first_dt(self, f, t0, tmax, y0, df, df_diagonal):
        raise NotImplementedError

    # tolerances handed to the fused kernels (0, 0 = no norm requested)
    kernel_tolerances = (0.0, 0.0)


class ConstantSteps(StepRule):
    """Always the same step, always accepted (step.py:28-52)."""

    def __init__(self, dt):
        self.dt = dt

    def __repr__(self):
        return f"{self.__class__.__name__}(dt={self.dt})"

    def suggest(self, previous_dt, scaled_error_estimate, local_convergence_rate=None):
        return self.dt

    def is_accepted(self, scaled_error_estimate):
        return True

    def scale_error_estimate(self, unscaled_error_estimate, reference_state):
        return None  # make sure nobody uses it downstream (step.py:44-46)

    def first_dt(self, f, t0, tmax, y0, df, df_diagonal):
        return self.dt


class AdaptiveSteps(StepRule):
    """Proportional control on the scaled local error (step.py:55-108)."""

    def __init__(self, abstol=1e-4, reltol=1e-2, max_changes=(0.2, 10.0), safety_scale=0.95,
                 min_step=1e-15, max_step=1e15):
        self.abstol = abstol
        self.reltol = reltol
        self.max_changes = max_changes
        self.safety_scale = safety_scale
        self.min_step = min_step
        self.max_step = max_step

    def __repr__(self):
        return f"{self.__class__.__name__}(abstol={self.abstol}, reltol={self.reltol})"

    @property
    def kernel_tolerances(self):
        return float(self.abstol), float(self.reltol)

    def suggest(self, previous_dt, scaled_error_estimate, local_convergence_rate=None):
        if local_convergence_rate is None:
            raise ValueError("Please provide a local convergence rate.")
        lower, upper = self.max_changes
        error = float(scaled_error_estimate)
        factor = self.safety_scale * (1.0 / error if error != 0.0 else math.inf) ** (1.0 / local_convergence_rate)
        factor = max(lower, min(factor, upper))
        return factor * previous_dt

    def is_accepted(self, scaled_error_estimate):
        return bool(scaled_error_estimate < 1)

    def scale_error_estimate(self, unscaled_error_estimate, reference_state):
        """RMS of err / (abstol + reltol * reference) (step.py:93-104); tensors or scalars."""
        err = torch.as_tensor(unscaled_error_estimate)
        ref = torch.as_tensor(reference_state)
        if err.ndim > 0 and err.shape != ref.shape:
            raise ValueError("Unscaled error estimate needs same shape as reference state.")
        ratio = err.to(ref.device) / (self.abstol + self.reltol * ref)
        dim = ratio.numel() if ratio.ndim > 0 else 1
        return float(torch.linalg.vector_norm(ratio.double().reshape(-1)) / math.sqrt(dim))

    def first_dt(self, f, t0, tmax, y0, df, df_diagonal):
        return propose_first_dt(f, t0, y0)


def norm_from_sq_sum(sq_sum, dim):
    """sqrt(sum ratio_i^2 / d): finish the norm whose sum(s) the kernel epilogue(s) produced.

    ``sq_sum`` is a device tensor with one entry per launch that contributed (interior / rim tiles of a sharded
    step); if it carries a ``ready_event`` (all-reduce on the communication stream) that is awaited first.
    """
    event = getattr(sq_sum, "ready_event", None)
    if event is not None:
        event.synchronize()
    total = float(sq_sum.sum()) if hasattr(sq_sum, "sum") else float(sq_sum)
    return math.sqrt(total / dim)


def propose_first_dt(f, t0, y0):
    """0.01 * ||y0|| / ||f(t0, y0)||  (step.py:111-115); f runs on the GPU."""
    from .ivp import as_device_tensor

    y0 = as_device_tensor(y0)
    norm_y0 = torch.linalg.vector_norm(y0)
    norm_dy0 = torch.linalg.vector_norm(f(t0, y0))
    return float(0.01 * norm_y0 / norm_dy0)
