"""EK1 solvers on the fused CUDA step.  Mirrors tornadox/ek1.py: ``BatchedEK1`` (:181-270) and
``DiagonalEK1`` (:273-369).

The whole reference chain -- precondition, predict_mean, evaluate_ode (f and df_diagonal), estimate_error,
predict_cov_sqrtm (batched QR), observe_cov_sqrtm, correct_cov_sqrtm, correct_mean, un-precondition,
reference_state -- is ONE kernel launch per attempt (``tdx_dek1_attempt_step``), which also reduces the
scaled error norm of ``AdaptiveSteps.scale_error_estimate`` (step.py:93-104).
The dense ``ReferenceEK1`` family (:11-178) is O(d^3) and out of scope.
"""

import torch

from . import _lib, odefilter, rv
from .engine import StepEngine
from .init import _spec_of
from .ivp import as_device_tensor


class BatchedEK1(odefilter.ODEFilter):
    """Common functionality for EK1 variations that act on batched multivariate normals (ek1.py:181-270)."""

    #: when False the per-coordinate error_estimate / reference_state vectors are not written to HBM
    #: (the scaled error norm is fused regardless); state.error_estimate / reference_state are then None
    materialize_error_vectors = True
    _has_controller = True

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.phi_1d = None
        self.sq_1d = None
        self.engine = None

    def _make_engine(self, f, device, y0=None):
        spec = _spec_of(f, y0)
        group = self.process_group
        rank = world = None
        if group is not None:
            import torch.distributed as dist

            rank, world = dist.get_rank(group), dist.get_world_size(group)
        return StepEngine(spec, self.num_derivatives, dtype=self.dtype, device=device, group=group,
                          rank=rank or 0, world=world or 1)

    def initialize(self, f, t0, tmax, y0, df, df_diagonal):
        """ek1.py:190-205.  ``y0`` is the GLOBAL initial value; a sharded solver keeps its own slice."""
        y0 = as_device_tensor(y0, dtype=self.dtype)
        self.engine = self._make_engine(f, y0.device, y0)
        A, _ = _lib.prior(self.num_derivatives)
        Ql = self.engine.Ql
        self.phi_1d = torch.as_tensor(A, dtype=self.dtype, device=y0.device)
        self.sq_1d = torch.as_tensor(Ql, dtype=self.dtype, device=y0.device)
        extended_dy0, cov_sqrtm = self.init(f=f, df=df, y0=y0, t0=t0, num_derivatives=self.num_derivatives)
        extended_dy0 = self._own_slice(extended_dy0.to(self.dtype))
        state = BatchedEK1.y0_to_initial_state(extended_dy0, cov_sqrtm.to(self.dtype), d=extended_dy0.shape[1], t0=t0)
        self.engine.barrier()      # sharded: nobody starts stepping before every rank has its initial state
        return state

    def _own_slice(self, mean_global):
        eng = self.engine
        if eng.world == 1:
            return mean_global.contiguous()
        from .engine import local_slices

        parts = [mean_global[:, s] for s in local_slices(eng.spec, eng.begin, eng.end)]
        return torch.cat(parts, dim=1).contiguous()

    @staticmethod
    def y0_to_initial_state(extended_dy0, cov_sqrtm, d, t0):
        """ek1.py:216-226: the (n, n) factor repeated d times, NaN error / reference."""
        blocks = cov_sqrtm.unsqueeze(0).expand(d, -1, -1).contiguous()
        nan = torch.full((d,), float("nan"), dtype=extended_dy0.dtype, device=extended_dy0.device)
        return odefilter.ODEFilterState(
            t=t0,
            y=rv.BatchedMultivariateNormal(extended_dy0, blocks),
            error_estimate=nan,
            reference_state=nan.clone(),
        )

    def attempt_step(self, state, dt, f, t0, tmax, y0, df, df_diagonal):
        """One fused launch (ek1.py:228-260)."""
        if self.engine is None:
            raise RuntimeError("call initialize() first")
        abstol, reltol = self.steprule.kernel_tolerances
        if state.y.mean.device.type == "cpu":
            return self._attempt_step_host(state, dt, abstol, reltol)
        if self.engine.external:      # a user-defined f: evaluated by the caller's torch code at the predicted state
            self.engine.evaluate_external(f, df_diagonal, state.t, dt, state.y.mean, self._uses_jacobian)
        mean, sc, err, ref, sq = self.engine.dek1_attempt(
            state.t, dt, state.y.mean, state.y.cov_sqrtm, abstol=abstol, reltol=reltol,
            want_err_ref=self.materialize_error_vectors, jacobian=self._uses_jacobian)
        new_state = odefilter.ODEFilterState(
            t=state.t + dt,
            y=rv.BatchedMultivariateNormal(mean, sc),
            error_estimate=err,
            reference_state=ref,
        )
        if abstol != 0.0 or reltol != 0.0:
            new_state.scaled_error_norm_sq_sum = sq
            new_state.scaled_error_norm_dim = self.engine.d_global
        info = dict(num_f_evaluations=1)
        if self._uses_jacobian:
            info["num_df_diagonal_evaluations"] = 1
        return new_state, info

    def perform_full_step(self, state, initial_dt, f, t0, tmax, y0, df, df_diagonal):
        """odefilter.py:171-223.  With one of the reference's own step rules and a device-resident state the whole
        accept / reject loop runs inside the library (``tdx_dek1_full_step``: no Python round trip and no stream
        synchronisation per attempt); anything else takes the generic host loop."""
        rule = self.engine.native_steprule(self.steprule) if self.engine is not None else None
        if rule is None or state.y.mean.device.type != "cuda" or not self.engine.can_run_full_steps():
            return super().perform_full_step(state, initial_dt, f, t0, tmax, y0, df, df_diagonal)
        mean, sc, err, ref, res = self.engine.dek1_full_step(
            rule, state.t, initial_dt, tmax, state.y.mean, state.y.cov_sqrtm,
            want_err_ref=self.materialize_error_vectors, jacobian=self._uses_jacobian)
        proposed = odefilter.ODEFilterState(t=res.t_new, y=rv.BatchedMultivariateNormal(mean, sc), error_estimate=err,
                                            reference_state=ref)
        assert res.dt_next >= 0, f"Invalid step size: dt={res.dt_next}"   # odefilter.py:221
        n_att = int(res.attempts)
        step_info = dict(num_f_evaluations=n_att, num_df_evaluations=0,
                         num_df_diagonal_evaluations=n_att if self._uses_jacobian else 0, num_attempted_steps=n_att)
        return proposed, res.dt_next, step_info

    perform_full_step_compiled = perform_full_step

    # ---- device-resident controller hooks (odefilter.ODEFilter.run_device_loop / _Follower) --------------------------------
    def _ctl_ring(self, state, nbuf, clone_first=False):
        mean0, sc0 = state.y.mean, state.y.cov_sqrtm
        if clone_first:
            mean0, sc0 = mean0.clone(), sc0.clone()
        means = [mean0] + [torch.empty_like(mean0) for _ in range(nbuf - 1)]
        factors = [sc0] + [torch.empty_like(sc0) for _ in range(nbuf - 1)]
        errs = refs = None
        if self.materialize_error_vectors:
            d = mean0.shape[1]
            errs = [torch.full((d,), float("nan"), dtype=mean0.dtype, device=mean0.device) for _ in range(nbuf)]
            refs = [torch.full((d,), float("nan"), dtype=mean0.dtype, device=mean0.device) for _ in range(nbuf)]
        return dict(means=means, factors=factors, errs=errs, refs=refs)

    def _ctl_launch(self, k, ring, stream=None, max_accepted=0):
        self.engine.dek1_run_attempts(k, ring["means"], ring["factors"], ring["errs"], ring["refs"],
                                      jacobian=self._uses_jacobian, stream=stream, max_accepted=max_accepted)

    def _ctl_state(self, ring, slot, t):
        err = ring["errs"][slot] if ring["errs"] is not None else None
        ref = ring["refs"][slot] if ring["refs"] is not None else None
        return odefilter.ODEFilterState(t=t, y=rv.BatchedMultivariateNormal(ring["means"][slot], ring["factors"][slot]),
                                        error_estimate=err, reference_state=ref)

    def _ctl_info(self, attempts, accepted):
        return dict(num_f_evaluations=attempts, num_df_evaluations=0,
                    num_df_diagonal_evaluations=attempts if self._uses_jacobian else 0, num_steps=accepted,
                    num_attempted_steps=attempts)

    def _attempt_step_host(self, state, dt, abstol, reltol):
        """``attempt_step`` for a state held in host memory (CPU tensors): chunked H2D | kernel | D2H pipeline inside the
        library.  The outputs live in two alternating sets of pinned buffers owned by the solver (allocating pinned memory
        per call would cost more than the step), so a returned state stays valid for ONE further host attempt."""
        n, d = state.y.mean.shape
        sets = getattr(self, "_host_out", None)
        if sets is None or sets[0][0].shape != (n, d) or sets[0][0].dtype != state.y.mean.dtype:
            def pinned(shape):
                return torch.empty(shape, dtype=state.y.mean.dtype, pin_memory=True)
            sets = [(pinned((n, d)), pinned((d, n, n))) for _ in range(2)]
            self._host_out, self._host_turn = sets, 0
        mean_out, sc_out = sets[self._host_turn]
        if mean_out.data_ptr() == state.y.mean.data_ptr():
            self._host_turn ^= 1
            mean_out, sc_out = sets[self._host_turn]
        self._host_turn ^= 1
        mean, sc, err, ref, sq = self.engine.dek1_attempt_host(
            state.t, dt, state.y.mean, state.y.cov_sqrtm, abstol=abstol, reltol=reltol, mean_out=mean_out, sc_out=sc_out,
            want_err_ref=self.materialize_error_vectors, jacobian=self._uses_jacobian)
        new_state = odefilter.ODEFilterState(
            t=state.t + dt, y=rv.BatchedMultivariateNormal(mean, sc), error_estimate=err, reference_state=ref)
        if abstol != 0.0 or reltol != 0.0:
            new_state.scaled_error_norm_sq_sum = sq
            new_state.scaled_error_norm_dim = self.engine.d_global
        info = dict(num_f_evaluations=1)
        if self._uses_jacobian:
            info["num_df_diagonal_evaluations"] = 1
        return new_state, info

    _uses_jacobian = True


class DiagonalEK1(BatchedEK1):
    """EK1 with the Jacobian truncated to its diagonal (ek1.py:273-369)."""

    _uses_jacobian = True
