"""Solver driver around the fused ``attempt_step`` kernels.

Mirrors the public interface of tornadox/odefilter.py -- ``ODEFilterState`` (:17-20), ``ODESolution``
(:23-28), ``ODEFilter`` with ``solve`` (:51-71), ``simulate_final_state`` (:73-78),
``solution_generator`` (:80-159), ``perform_full_step`` (:171-223) and the ``stop_at`` time stopper
(:355-373).  The reference's jitted ``perform_full_step_compiled`` (:225-344) compiles the accept/reject
loop into ``lax.while_loop``; the counterpart here is the device-resident step controller: ONE persistent
cooperative kernel runs attempt after attempt (``tdx_*_run_attempts``), deciding accept / reject, the next
step and the time stops on the device.  ``simulate_final_state`` runs whole batches of attempts per launch;
``solve`` / ``solution_generator`` follow the running kernel from the host through a ring of state buffers
(dense output without a host round trip per step).  A host-driven loop (one scalar read per attempt) remains
for user-defined step rules and for states held in host memory.
"""

import dataclasses
import math
import time
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

        ivp = args[0] if args else kwargs.get("ivp")
        stop_at = kwargs.get("stop_at", args[1] if len(args) > 1 else None)
        if self.device_loop and ivp is not None and len(args) <= 2 and not kwargs.get("progressbar"):
            out = self._solve_following(ivp, stop_at, offload)
            if out is not None:
                return out

        def host_keep(x):
            # states of the host-buffer path alias the solver's two alternating pinned output sets: keep a copy
            if isinstance(x, torch.Tensor) and x.device.type == "cpu" and x.is_pinned():
                return x.clone()
            return keep(x)

        times, means, cov_sqrtms, info = [], [], [], dict()
        for state, info in self.solution_generator(*args, **dict(kwargs, device_loop=False)):
            times.append(float(state.t))
            means.append(host_keep(state.y.mean))
            if isinstance(self, ek0.KroneckerEK0):
                n, d = state.y.mean.shape
                if (n * d) ** 2 <= self.dense_output_limit:
                    cov_sqrtms.append(keep(state.y.dense_cov_sqrtm()))   # as upstream: small problems only
                else:
                    cov_sqrtms.append(keep(state.y.cov_sqrtm_2))
            else:
                cov_sqrtms.append(host_keep(state.y.cov_sqrtm))
        if offload and torch.cuda.is_available():
            torch.cuda.synchronize()
        return ODESolution(
            t=torch.tensor(times, dtype=torch.float64),
            mean=torch.stack(means),
            cov_sqrtm=torch.stack(cov_sqrtms),
            info=info,
        )

    def _solve_following(self, ivp, stop_at, offload):
        """``solve`` on top of the device-resident controller: persistent launches on their own stream, each running until
        ``ring_size - 1`` further steps have been accepted; the host follows the kernel's progress through mapped pinned memory
        and copies every accepted state out of the state ring (copy engine) while the kernel is already working on the next
        attempts.  None: not available for this configuration."""
        from . import ek0

        state = self.initialize(*ivp)
        if not self._controller_available(state):
            return None
        kron = isinstance(self, ek0.KroneckerEK0)
        dense = kron and (state.y.mean.shape[0] * state.y.mean.shape[1]) ** 2 <= self.dense_output_limit
        factor0 = state.y.cov_sqrtm_2 if kron else state.y.cov_sqrtm

        def like(src):
            if offload:
                return torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
            return torch.empty_like(src)

        def grab(src):
            dst = like(src)
            dst.copy_(src, non_blocking=True)
            return dst

        times, means, factors = [float(state.t)], [grab(state.y.mean)], [grab(factor0)]
        info = self._ctl_info(0, 0)
        if state.t < ivp.tmax:
            follower = _Follower(self, state, self.steprule.first_dt(*ivp), ivp.tmax, stop_at)
            try:
                for j, t, attempts, dst in follower.states(lambda: (like(state.y.mean), like(factor0))):
                    times.append(t)
                    means.append(dst[0])
                    factors.append(dst[1])
                    info = self._ctl_info(attempts, j)
            finally:
                follower.close()
        if torch.cuda.is_available():
            torch.cuda.current_stream(state.y.mean.device).synchronize()
        if dense:       # as upstream: the dense kron(I_d, C) factor of every state, small problems only
            d = state.y.mean.shape[1]
            factors = [torch.kron(torch.eye(d, dtype=f.dtype, device=f.device), f) for f in factors]
        return ODESolution(t=torch.tensor(times, dtype=torch.float64), mean=torch.stack(means), cov_sqrtm=torch.stack(factors),
                           info=info)

    #: attempts enqueued per look at the device-resident controller (``simulate_final_state`` fast path)
    device_loop_batch = 64
    #: set False to force the host-driven loop everywhere
    device_loop = True

    #: state buffers the controller rotates through when the host follows the solve (``solve``, ``solution_generator``)
    ring_size = 4

    # ---- device-resident controller: hooks of the fused solvers ---------------------------------------------------------
    engine = None
    _has_controller = False

    def _controller_available(self, state):
        eng = self.engine
        return (eng is not None and eng.native_steprule(self.steprule) is not None and state.y.mean.device.type == "cuda"
                and eng.can_run_full_steps())

    def _ctl_ring(self, state, nbuf, clone_first=False):
        raise NotImplementedError

    def _ctl_launch(self, k, ring, stream=None):
        raise NotImplementedError

    def _ctl_state(self, ring, slot, t):
        raise NotImplementedError

    def _ctl_info(self, attempts, accepted):
        raise NotImplementedError

    def _device_loop_final_state(self, ivp, stop_at=None):
        """The whole solve under the device-resident step controller; None = not available for this configuration."""
        if not self._has_controller:
            return None
        state = self.initialize(*ivp)
        if not self._controller_available(state):
            return None
        if not state.t < ivp.tmax:
            return state, self._ctl_info(0, 0)
        return self.run_device_loop(state, self.steprule.first_dt(*ivp), ivp.tmax, stop_at=stop_at)

    def run_device_loop(self, state, dt0, tmax, stop_at=None):
        """Advance ``state`` to ``tmax`` under the device-resident step controller, starting with step ``dt0``: launches of
        ``device_loop_batch`` attempts each, ONE persistent kernel per launch, one look at the controller per launch.  The
        input state's buffers become slot 0 of the state ring.  Returns (final state, info)."""
        eng = self.engine
        rule = eng.native_steprule(self.steprule)
        ring = self._ctl_ring(state, 2)
        eng.ctl_begin(rule, state.t, dt0, tmax, nbuf=2, stop_at=stop_at)
        while True:
            self._ctl_launch(self.device_loop_batch, ring)
            ctl = eng.ctl_read()
            eng.raise_if_failed(ctl)
            if ctl.done:
                break
        return self._ctl_state(ring, ctl.cur, ctl.t), self._ctl_info(int(ctl.attempts), int(ctl.accepted))

    def simulate_final_state(self, *args, **kwargs):
        """Drain the generator, return the last (state, info) (odefilter.py:73-78).

        Nobody looks at the intermediate states here, so when the solver has a device-resident step controller the
        whole solve runs without the host in the loop (batches of attempts per launch, one look at the controller per
        launch); ``stop_at`` becomes the controller's list of time stops."""
        ivp = args[0] if args else kwargs.get("ivp")
        stop_at = kwargs.get("stop_at", args[1] if len(args) > 1 else None)
        plain = len(args) <= 2 and not kwargs.get("progressbar")
        if self.device_loop and plain and ivp is not None:
            out = self._device_loop_final_state(ivp, stop_at=stop_at)
            if out is not None:
                return out
        state, info = None, None
        for state, info in self.solution_generator(*args, **dict(kwargs, device_loop=False)):
            pass
        return state, info

    def solution_generator(self, ivp, stop_at=None, progressbar=False, compile_step=True, compile_init=False, device_loop=None):
        """Yield (state, info) after initialisation and after every accepted step (odefilter.py:80-159).

        With a device-resident controller (``device_loop``; default: the solver's ``device_loop`` attribute) the persistent
        kernel runs ahead on its own stream and the states that are yielded are copies taken out of the state ring;
        otherwise the host drives one full step at a time and yields the step's own output buffers."""
        del compile_step, compile_init  # no tracing compiler here; the step is a CUDA kernel either way
        use_ctl = self.device_loop if device_loop is None else device_loop
        state = self.initialize(*ivp)
        info = dict(num_f_evaluations=0, num_df_evaluations=0, num_df_diagonal_evaluations=0, num_steps=0,
                    num_attempted_steps=0)
        yield state, info

        dt = self.steprule.first_dt(*ivp)
        if use_ctl and not progressbar and state.t < ivp.tmax and self._controller_available(state):
            follower = _Follower(self, state, dt, ivp.tmax, stop_at, clone_first=True)
            try:
                for j, t, attempts, copy in follower.states(None):
                    yield self._ctl_state(copy, 0, t), self._ctl_info(attempts, j)
            finally:
                follower.close()
            return

        stopper = _TimeStopper(stop_at) if stop_at is not None else None
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
        # the kernel-reduced sum implements exactly AdaptiveSteps.scale_error_estimate: a subclass that overrides it (or the
        # tolerances' meaning) must see its own method
        fused_ok = type(self.steprule).scale_error_estimate is step.AdaptiveSteps.scale_error_estimate
        if sq_sum is not None and fused_ok:
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


class _Follower:
    """Host side of a solve whose attempts run in persistent kernels on their own stream (``tdx_*_run_attempts`` with
    ``follow``).  Every launch goes on until ``nbuf - 1`` further steps have been accepted (or the solve is over), so no ring
    slot is written twice within a launch; the host reads the kernel's progress from mapped pinned memory and copies newly
    accepted states out of the ring on a second stream while the kernel is already working on the next attempts.  Everything
    that may synchronise the device (memory allocation, the controller read) happens BETWEEN launches."""

    def __init__(self, solver, state, dt0, tmax, stop_at, clone_first=False):
        eng = solver.engine
        self.solver, self.eng = solver, eng
        self.nbuf = max(3, min(int(solver.ring_size), 8))
        self.ring = solver._ctl_ring(state, self.nbuf, clone_first=clone_first)
        dev = state.y.mean.device
        self.loop_stream = torch.cuda.Stream(device=dev)
        self.copy_stream = torch.cuda.Stream(device=dev)
        current = torch.cuda.current_stream(dev)
        self.loop_stream.wait_stream(current)
        self.copy_stream.wait_stream(current)
        eng.ctl_begin(eng.native_steprule(solver.steprule), state.t, dt0, tmax, nbuf=self.nbuf, stop_at=stop_at, follow=True)
        self.seen = 0
        self.running = None        # event behind the launch in flight
        self.closed = False

    def _launch(self):
        import ctypes

        self.solver._ctl_launch(2 ** 30, self.ring, stream=ctypes.c_void_p(self.loop_stream.cuda_stream),
                                max_accepted=self.nbuf - 1)
        self.running = torch.cuda.Event()
        self.running.record(self.loop_stream)

    def states(self, make_dst):
        """Yield (j, t_j, attempts so far, copy) for every accepted state, in order.  ``make_dst()`` returns the (mean, factor)
        tensors a state is copied into; None: copies of every ring array of the slot (a dict like the ring, one-element lists)."""
        eng, m = self.eng, self.nbuf - 1
        current = torch.cuda.current_stream(self.copy_stream.device)
        while True:
            # destinations first: allocating device / pinned memory may synchronise with a running kernel
            if make_dst is None:
                dsts = [{k: (None if v is None else [None if v[0] is None else torch.empty_like(v[0])]) for k, v in self.ring.items()}
                        for _ in range(m)]
            else:
                dsts = [make_dst() for _ in range(m)]
            self._launch()
            got, events = 0, []
            while True:
                ended = self.running.query()             # read BEFORE the counters: nothing follows a finished launch
                accepted, _ = eng.ctl_poll()
                while self.seen < accepted and got < m:
                    self.seen += 1
                    slot = self.seen % self.nbuf
                    t, _, attempts = eng.ctl_history(self.seen)
                    dst = dsts[got]
                    got += 1
                    with torch.cuda.stream(self.copy_stream):
                        if make_dst is None:
                            for key, arrs in self.ring.items():
                                if arrs is not None and arrs[slot] is not None:
                                    dst[key][0].copy_(arrs[slot], non_blocking=True)
                        else:
                            dst[0].copy_(self.ring["means"][slot], non_blocking=True)
                            dst[1].copy_(self.ring["factors"][slot], non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record(self.copy_stream)
                    events.append((self.seen, t, attempts, dst, ev))
                # hand over what has been copied completely, in order
                while events and events[0][4].query():
                    j, t, attempts, dst, ev = events.pop(0)
                    current.wait_event(ev)
                    yield j, t, attempts, dst
                if ended and self.seen >= eng.ctl_poll()[0]:
                    break
                time.sleep(0)
            for j, t, attempts, dst, ev in events:
                ev.synchronize()
                current.wait_event(ev)
                yield j, t, attempts, dst
            self.loop_stream.synchronize()
            self.running = None
            ctl = eng.ctl_read()
            eng.raise_if_failed(ctl)
            if ctl.done:
                break

    def close(self):
        if self.closed:
            return
        self.closed = True
        if self.running is not None:
            # the consumer stopped early, or an exception is on its way: stop the launch at its next attempt boundary
            if not self.running.query():
                self.eng.ctl_abort()
            self.loop_stream.synchronize()
            self.running = None
        self.copy_stream.synchronize()
        torch.cuda.current_stream(self.copy_stream.device).wait_stream(self.copy_stream)


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
