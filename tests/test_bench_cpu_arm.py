"""CPU: bench.py's reference arm cuts the benchmark's own problem into one strip per core (halo rows and, for KroneckerEK0,
z^T z exchanged through shared memory).  The strips together must reproduce the unsharded oracle -- i.e. the CPU arm really
runs the same configuration as the GPU arm."""

import argparse
import pathlib

import numpy as np
import pytest

from oracle.ref_np import driver as odriver
from oracle.ref_np import filters as ofilters

ROOT = pathlib.Path(__file__).resolve().parent.parent


def _bench():
    import bench          # the repository root is on sys.path (conftest); worker processes import it under the same name

    return bench


@pytest.mark.parametrize("workload,solver,size", [("fhn_2d", "dek1", 12), ("fhn_2d", "ek0", 12), ("brusselator", "dek1", 37),
                                                   ("brusselator", "ek0", 37)])
def test_strip_workers_reproduce_the_unsharded_oracle(workload, solver, size):
    bench = _bench()
    args = argparse.Namespace(workload=workload, grid=size, points=size, solver=solver, dtype="f64")
    n, d = bench.NU + 1, bench.problem_dim(args)
    got = np.zeros((n, d))
    value, cores, sample, ms = bench.cpu_reference(args, solver, steps=2, warmup=1, cores=3, final_mean_out=got)
    assert cores == 3 and value > 0
    y0, prob, series = bench.initial_condition(args)
    dt = 0.5 * odriver.propose_first_dt(prob.f, prob.t0, y0)
    m0, _ = odriver.taylor_init(series, y0, prob.t0, bench.NU)
    if solver == "dek1":
        state = ofilters.FilterState(0.0, m0, np.zeros((d, n, n)), None, None)
        step = lambda s: ofilters.diagonal_ek1_attempt(s, dt, prob.f, prob.df_diagonal, bench.NU)[0]
    else:
        state = ofilters.FilterState(0.0, m0, np.zeros((n, n)), None, None)
        step = lambda s: ofilters.kronecker_ek0_attempt(s, dt, prob.f, bench.NU)[0]
    for _ in range(3):
        state = step(state)
    scale = np.max(np.abs(state.mean))
    assert np.max(np.abs(got - state.mean)) <= 1e-12 * scale
    # both arms describe the same workload
    assert bench.shared_config(args)["dims"] == d
