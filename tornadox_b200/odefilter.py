"""Solver driver: the host loop around the fused ``attempt_step`` kernels.

Mirrors the public interface of tornadox/odefilter.py -- ``ODEFilterState`` (:17-20), ``ODESolution``
(:23-28), ``ODEFilter`` with ``solve`` (:51-71), ``simulate_final_state`` (:73-78),
``solution_generator`` (:80-159), ``perform_full_step`` (:171-223) and the ``stop_at`` time stopper
(:355-373).  The reference's jitted ``perform_full_step_compiled`` (:225-344) is the same
accept/reject logic inside ``lax.while_loop``; here the loop runs on the host and reads ONE fp64 scalar
per attempt (the error norm the kernel epilogue reduced), so ``compile_step`` is accepted and ignored.
"""

import dataclasses
import math
from abc import ABC, abstractmethod
from collections import namedtuple
from typing import Dict, Iterable

import torch

from . import init, step


class DeferredAbsRow:
    """``|mean[0]|`` computed when somebody asks for it.  KroneckerEK0's reference_state (ek0.py:184) is implied by the
    new mean; the fused kernel already consumed it for the error norm, so writing it to HBM every attempt (8 B per
    dimension) only pays off if the caller reads it."""

    def __init__(self, mean):
        self._mean, self._value = mean, None

    def get(self):
        if self._value is None:
            self._value = self._mean[0].abs()
            self._mean = None
        return self._value


class ODEFilterState(namedtuple("_ODEFilterState", "t y error_estimate reference_state")):
    """(t, y, error_estimate, reference_state) as in odefilter.py:17-20.

    Instances produced by the fused solvers additionally carry ``scaled_error_norm_sq_sum``: the device
    scalar the kernel epilogue reduced (None otherwise); see ``ODEFilter._scaled_error_norm``.
    ``reference_state`` may be stored as a ``DeferredAbsRow``; attribute access resolves it.
    """

    scaled_error_norm_sq_sum = None
    scaled_error_norm_dim = None

    @property
    def reference_state(self):
        raw = tuple.__getitem__(self, 3)
        return raw.get() if isinstance(raw, DeferredAbsRow) else raw


@dataclasses.dataclass(frozen=False)
class ODESolution:
    t: torch.Tensor
    mean: torch.Tensor
    cov_sqrtm: torch.Tensor
    info: Dict


_COUNTERS = ("num_f_evaluations", "num_df_evaluations", "num_df_diagonal_evaluations")


class ODEFilter(ABC):
    """Interface for filtering-based ODE solvers (odefilter.py:31-49)."""

    def __init__(self, *, steprule=None, num_derivatives=4, initialization=None, dtype=torch.float64,
                 process_group=None):
        self.steprule = steprule or step.AdaptiveSteps()
        self.num_derivatives = num_derivatives
        self.iwp = None
        self.init = initialization or init.TaylorMode()
        self.dtype = dtype
        self.process_group = process_group

    def __repr__(self):
        return (
            f"{self.__class__.__name__}(num_derivatives={self.num_derivatives}, "
            f"steprule={self.steprule}, initialization={self.init})"
        )

    # ---- user-facing drivers ----------------------------------------------------------------------
    #: ``solve`` materialises KroneckerEK0's dense kron(I_d, C) factor per state (as the reference does) only while it
    #: stays below this many entries; beyond that the (n, n) factor itself is stored per state
    dense_output_limit = 1 << 22

    def solve(self, *args, offload=False, **kwargs):
        """Collect every accepted state (odefilter.py:51-71).

        ``offload=True`` streams the accepted states to pinned host memory as they are produced (asynchronous device-to-host
        copies behind the step kernels), so that a long solve of a large problem does not have to fit on the GPU; the
        returned tensors then live on the CPU.  For KroneckerEK0 the reference stacks the dense ``kron(I_d, C)`` factor of
        every state (O(d^2 n^2) each); here that happens only for small problems (``dense_output_limit``), otherwise
        ``cov_sqrtm`` holds the (n, n) factors."""
        from . import ek0

        def keep(x):
            if not offload or not isinstance(x, torch.Tensor) or x.device.type == "cpu":
                return x
            host = torch.empty(x.shape, dtype=x.dtype, pin_memory=True)
            host.copy_(x, non_blocking=True)
            return host

        times, means, cov_sqrtms, info = [], [], [], dict()
        for state, info in self.solution_generator(*args, **kwargs):
            times.append(float(state.t))
            means.append(keep(state.y.mean))
            if isinstance(self, ek0.KroneckerEK0):
                n, d = state.y.mean.shape
                if (n * d) ** 2 <= self.dense_output_limit:
                    cov_sqrtms.append(keep(state.y.dense_cov_sqrtm()))   # as upstream: small problems only
                else:
                    cov_sqrtms.append(keep(state.y.cov_sqrtm_2))
            else:
                cov_sqrtms.append(keep(state.y.cov_sqrtm))
        if offload and torch.cuda.is_available():
            torch.cuda.synchronize()
        return ODESolution(
            t=torch.tensor(times, dtype=torch.float64),
            mean=torch.stack(means),
            cov_sqrtm=torch.stack(cov_sqrtms),
            info=info,
        )

    #: attempts enqueued per look at the device-resident controller (``simulate_final_state`` fast path)
    device_loop_batch = 64
    #: set False to force the host-driven loop everywhere
    device_loop = True

    def _device_loop_final_state(self, ivp):
        """Solvers with a device-resident step controller override this; None = not available for this configuration."""
        return None

    def simulate_final_state(self, *args, **kwargs):
        """Drain the generator, return the last (state, info) (odefilter.py:73-78).

        Nobody looks at the intermediate states here, so when the solver has a device-resident step controller the
        whole solve runs without the host in the loop (batches of attempts, one look at the controller per batch)."""
        ivp = args[0] if args else kwargs.get("ivp")
        plain = len(args) <= 1 and not kwargs.get("stop_at") and not kwargs.get("progressbar")
        if self.device_loop and plain and ivp is not None:
            out = self._device_loop_final_state(ivp)
            if out is not None:
                return out
        state, info = None, None
        for state, info in self.solution_generator(*args, **kwargs):
            pass
        return state, info

    def solution_generator(self, ivp, stop_at=None, progressbar=False, compile_step=True, compile_init=False):
        """Yield (state, info) after initialisation and after every accepted step (odefilter.py:80-159)."""
        del compile_step, compile_init  # no tracing compiler here; the step is a CUDA kernel either way
        stopper = _TimeStopper(stop_at) if stop_at is not None else None
        state = self.initialize(*ivp)
        info = dict(num_f_evaluations=0, num_df_evaluations=0, num_df_diagonal_evaluations=0, num_steps=0,
                    num_attempted_steps=0)
        yield state, info

        dt = self.steprule.first_dt(*ivp)
        bar = _ProgressBar(ivp.tmax) if progressbar else None
        while state.t < ivp.tmax:
            if bar is not None:
                bar.update(state.t, dt)
            if stopper is not None:
                dt = stopper.adjust_dt_to_time_stops(state.t, dt)
            state, dt, step_info = self.perform_full_step(state, dt, *ivp)
            info["num_steps"] += 1
            info["num_attempted_steps"] += step_info["num_attempted_steps"]
            for key in _COUNTERS:
                info[key] += step_info[key]
            yield state, info
        if bar is not None:
            bar.close()

    # ---- accept / reject loop ---------------------------------------------------------------------
    def perform_full_step(self, state, initial_dt, f, t0, tmax, y0, df, df_diagonal):
        """Attempt steps from ``state`` until the step rule accepts one (odefilter.py:171-223)."""
        dt = initial_dt
        rate = self.num_derivatives + 1
        step_info = dict(num_f_evaluations=0, num_df_evaluations=0, num_df_diagonal_evaluations=0,
                         num_attempted_steps=0)
        while True:
            proposed, attempt_info = self.attempt_step(state, dt, f, t0, tmax, y0, df, df_diagonal)
            step_info["num_attempted_steps"] += 1
            for key in _COUNTERS:
                step_info[key] += attempt_info.get(key, 0)

            error_norm = self._scaled_error_norm(proposed, dt)
            if error_norm is not None and math.isnan(error_norm):
                # the reference would spin forever here (NaN < 1 is False, odefilter.py:179-219)
                raise FloatingPointError(f"error estimate is NaN at t={state.t}, dt={dt}")
            accepted = self.steprule.is_accepted(error_norm)
            suggested = self.steprule.suggest(dt, error_norm, local_convergence_rate=rate)
            remaining = tmax - (proposed.t if accepted else state.t)
            dt = min(suggested, remaining)
            assert dt >= 0, f"Invalid step size: dt={dt}"
            if accepted:
                return proposed, dt, step_info

    # the reference's jitted variant has identical semantics; keep the name for drop-in use
    perform_full_step_compiled = perform_full_step

    def _scaled_error_norm(self, proposed, dt):
        """steprule.scale_error_estimate(dt * error, reference) (odefilter.py:205-210).

        The fused solvers already reduced sum_i ratio_i^2 on the device; only the sqrt remains.
        """
        if isinstance(self.steprule, step.ConstantSteps):
            return None
        sq_sum = getattr(proposed, "scaled_error_norm_sq_sum", None)
        if sq_sum is not None:
            return step.norm_from_sq_sum(sq_sum, proposed.scaled_error_norm_dim)
        if proposed.error_estimate is None:
            return self.steprule.scale_error_estimate(None, proposed.reference_state)
        return self.steprule.scale_error_estimate(dt * proposed.error_estimate, proposed.reference_state)

    @abstractmethod
    def initialize(self, f, t0, tmax, y0, df, df_diagonal):
        raise NotImplementedError

    @abstractmethod
    def attempt_step(self, state, dt, f, t0, tmax, y0, df, df_diagonal):
        raise NotImplementedError


class _TimeStopper:
    """Clip dt so that the solver lands on the requested time points (odefilter.py:355-373)."""

    def __init__(self, locations: Iterable):
        self._locations = iter(locations)
        self._next_location = next(self._locations)

    def adjust_dt_to_time_stops(self, t, dt):
        if t >= self._next_location:
            self._next_location = next(self._locations, math.inf)
        if t + dt > self._next_location:
            dt = self._next_location - t
        return dt


class _ProgressBar:
    """tqdm bar over [0, tmax] in 100 ticks (odefilter.py:121-132, 157-159)."""

    def __init__(self, tmax, ticks=100):
        from tqdm import tqdm

        self.bar = tqdm(total=ticks)
        self.increment = tmax / ticks
        self.threshold = self.increment

    def update(self, t, dt):
        while t + dt >= self.threshold:
            self.bar.update()
            self.threshold += self.increment
        self.bar.set_description(f"t={t:.4f}, dt={dt:.2E}")

    def close(self):
        self.bar.update()
        self.bar.close()
