"""EK0 solvers on the fused CUDA step.  Mirrors tornadox/ek0.py: ``KroneckerEK0`` (:85-193) and
``DiagonalEK0`` (:196-280, = the DiagonalEK1 kernel with the Jacobian diagonal forced to zero).

KroneckerEK0 keeps ONE (n, n) factor for all coordinates; an attempt is two d-wide sweeps over the mean
with the global calibration scalar in between, fused into ONE cooperative launch (``tdx_ek0_attempt_step``;
the three phases ``tdx_ek0_sweep_a`` -> ``tdx_ek0_mid`` -> ``tdx_ek0_sweep_b`` remain for callers that all-reduce
between them themselves).
The dense ``ReferenceEK0`` (:13-82) is O(d^3) and out of scope.
"""

import torch

from . import _lib, odefilter, rv
from .ek1 import BatchedEK1
from .ivp import as_device_tensor


class KroneckerEK0(odefilter.ODEFilter):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.A = None
        self.Ql = None
        self.engine = None

    #: when False (default), reference_state = |mean[0]| is not written by the kernel every attempt: it is implied by the
    #: new mean and is computed on first access of ``state.reference_state`` (``odefilter.DeferredAbsRow``)
    materialize_reference_state = False
    _has_controller = True

    def initialize(self, f, t0, tmax, y0, df, df_diagonal):
        """ek0.py:94-121: single (n, n) factor, scalar NaN error estimate."""
        y0 = as_device_tensor(y0, dtype=self.dtype)
        self.engine = BatchedEK1._make_engine(self, f, y0.device, y0)
        A, _ = _lib.prior(self.num_derivatives)
        Ql = self.engine.Ql
        self.A = torch.as_tensor(A, dtype=self.dtype, device=y0.device)
        self.Ql = torch.as_tensor(Ql, dtype=self.dtype, device=y0.device)
        extended_dy0, cov_sqrtm = self.init(f=f, df=df, y0=y0, t0=t0, num_derivatives=self.num_derivatives)
        mean = BatchedEK1._own_slice(self, extended_dy0.to(self.dtype))
        d = mean.shape[1]
        y = rv.LeftIsotropicMatrixNormal(mean=mean, d=d, cov_sqrtm_2=cov_sqrtm.to(self.dtype).contiguous())
        nan = torch.full((d,), float("nan"), dtype=self.dtype, device=mean.device)
        self.engine.barrier()      # sharded: nobody starts stepping before every rank has its initial state
        return odefilter.ODEFilterState(t=t0, error_estimate=float("nan"), reference_state=nan, y=y)

    def attempt_step(self, state, dt, f, t0, tmax, y0, df, df_diagonal):
        """ek0.py:152-193 as one cooperative launch."""
        if self.engine is None:
            raise RuntimeError("call initialize() first")
        abstol, reltol = self.steprule.kernel_tolerances
        if self.engine.external:      # a user-defined f: evaluated by the caller's torch code at the predicted state
            self.engine.evaluate_external(f, None, state.t, dt, state.y.mean, False)
        mean, Cl, err, ref, itol = self.engine.ek0_attempt(
            state.t, dt, state.y.mean, state.y.cov_sqrtm_2, abstol=abstol, reltol=reltol,
            want_ref=self.materialize_reference_state)
        new_state = odefilter.ODEFilterState(
            t=state.t + dt,
            error_estimate=err,
            reference_state=ref if ref is not None else odefilter.DeferredAbsRow(mean),
            y=rv.LeftIsotropicMatrixNormal(mean, state.y.d, Cl),
        )
        if abstol != 0.0 or reltol != 0.0:
            new_state.scaled_error_norm_sq_sum = itol   # sum_i (dt err / tol_i)^2, reduced by sweep B
            new_state.scaled_error_norm_dim = self.engine.d_global
        return new_state, dict(num_f_evaluations=1)


    def perform_full_step(self, state, initial_dt, f, t0, tmax, y0, df, df_diagonal):
        """odefilter.py:171-223 inside the library (``tdx_ek0_full_step``) when the step rule is one of the reference's."""
        rule = self.engine.native_steprule(self.steprule) if self.engine is not None else None
        if rule is None or state.y.mean.device.type != "cuda" or not self.engine.can_run_full_steps():
            return super().perform_full_step(state, initial_dt, f, t0, tmax, y0, df, df_diagonal)
        mean, Cl, err, ref, res = self.engine.ek0_full_step(
            rule, state.t, initial_dt, tmax, state.y.mean, state.y.cov_sqrtm_2, want_ref=self.materialize_reference_state)
        proposed = odefilter.ODEFilterState(t=res.t_new, error_estimate=err,
                                            reference_state=ref if ref is not None else odefilter.DeferredAbsRow(mean),
                                            y=rv.LeftIsotropicMatrixNormal(mean, state.y.d, Cl))
        assert res.dt_next >= 0, f"Invalid step size: dt={res.dt_next}"   # odefilter.py:221
        n_att = int(res.attempts)
        step_info = dict(num_f_evaluations=n_att, num_df_evaluations=0, num_df_diagonal_evaluations=0,
                         num_attempted_steps=n_att)
        return proposed, res.dt_next, step_info

    perform_full_step_compiled = perform_full_step

    # ---- device-resident controller hooks (odefilter.ODEFilter.run_device_loop / _Follower) --------------------------------
    def _ctl_ring(self, state, nbuf, clone_first=False):
        mean0, Cl0 = state.y.mean, state.y.cov_sqrtm_2
        if clone_first:
            mean0, Cl0 = mean0.clone(), Cl0.clone()
        means = [mean0] + [torch.empty_like(mean0) for _ in range(nbuf - 1)]
        factors = [Cl0] + [torch.empty_like(Cl0) for _ in range(nbuf - 1)]
        err = torch.full((nbuf,), float("nan"), dtype=mean0.dtype, device=mean0.device)
        errs = [err[i] for i in range(nbuf)]
        refs = None
        if self.materialize_reference_state:
            refs = [torch.full((mean0.shape[1],), float("nan"), dtype=mean0.dtype, device=mean0.device) for _ in range(nbuf)]
        return dict(means=means, factors=factors, errs=errs, refs=refs)

    def _ctl_launch(self, k, ring, stream=None, max_accepted=0):
        self.engine.ek0_run_attempts(k, ring["means"], ring["factors"], ring["errs"], ring["refs"], stream=stream,
                                     max_accepted=max_accepted)

    def _ctl_state(self, ring, slot, t):
        mean = ring["means"][slot]
        ref = ring["refs"][slot] if ring["refs"] is not None else odefilter.DeferredAbsRow(mean)
        return odefilter.ODEFilterState(t=t, error_estimate=ring["errs"][slot], reference_state=ref,
                                        y=rv.LeftIsotropicMatrixNormal(mean, mean.shape[1], ring["factors"][slot]))

    def _ctl_info(self, attempts, accepted):
        return dict(num_f_evaluations=attempts, num_df_evaluations=0, num_df_diagonal_evaluations=0, num_steps=accepted,
                    num_attempted_steps=attempts)


class DiagonalEK0(BatchedEK1):
    """ek0.py:196-280: block-diagonal EK0 = DiagonalEK1 with a zero Jacobian diagonal."""

    _uses_jacobian = False
