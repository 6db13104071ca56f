"""CPU, world_size 2 (gloo): the multi-GPU plumbing -- shard ranges, halo exchange (incl. the cyclic
two-rank case), scalar all-reduce -- with the NumPy oracle standing in for the per-shard kernels.
Every rank must reproduce its slice of the UNSHARDED oracle step and the same global error norm."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.ref_np import driver as odriver
from oracle.ref_np import filters as ofilters
from oracle.ref_np import prior as oprior
from oracle.ref_np import problems as oproblems

NU = 3


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _spec_and_problem(kind):
    from tornadox_b200 import ivp as pivp

    if kind == "fhn_2d":
        side = 6
        y0 = np.random.default_rng(2).uniform(size=2 * side * side)
        return pivp.fhn_2d(dx=1.0 / side, y0=y0).spec, oproblems.fhn_2d(dx=1.0 / side, y0=y0)
    if kind == "brusselator":
        return pivp.brusselator(N=9).spec, oproblems.brusselator(N=9)
    if kind == "lorenz96":
        return pivp.lorenz96(num_variables=11).spec, oproblems.lorenz96(num_variables=11)
    raise KeyError(kind)


def _local_f_with_halos(kind, spec, prob, begin, end, halo_lo, halo_hi):
    """f and diag J on the local shard given the neighbours' edge values (what the kernels do with halo buffers)."""
    nf = spec.n_fields
    if kind == "fhn_2d":
        ny = spec.grid_cols
        rows = end - begin

        def extended(y_local):
            fields = y_local.reshape(nf, rows, ny)
            parts, top, bottom = [], halo_lo is not None, halo_hi is not None
            ext = []
            for f in range(nf):
                blocks = []
                if top:
                    blocks.append(halo_lo.reshape(nf, ny)[f][None, :])
                blocks.append(fields[f])
                if bottom:
                    blocks.append(halo_hi.reshape(nf, ny)[f][None, :])
                ext.append(np.concatenate(blocks, axis=0))
            return ext, top, bottom

        dx, a, b, k, tau = spec.params

        def f(_, y_local):
            ext, top, bottom = extended(y_local)
            lap = [oproblems.laplace_2d(e, dx)[(1 if top else 0): (e.shape[0] - (1 if bottom else 0))] for e in ext]
            u, v = y_local.reshape(nf, -1)
            du, dv = lap[0].reshape(-1), lap[1].reshape(-1)
            return np.concatenate((a * du + u - u**3 - v + k, (b * dv + u - v) / tau))

        def jd(_, y_local):
            full = prob.df_diagonal(None, np.zeros(2 * spec.grid_rows * ny))   # Laplacian part is state independent
            u = y_local.reshape(nf, -1)[0]
            npts = spec.grid_rows * ny
            lo, hi = begin * ny, end * ny
            ju = full[lo:hi] + 3.0 * 0.0 - 3.0 * u**2      # full[u] = a*dl + 1 - 0
            jv = full[npts + lo: npts + hi]
            return np.concatenate((ju, jv))

        return f, jd
    if kind == "brusselator":
        N = spec.grid_cols
        c = (1.0 / 50.0) * (N + 1) ** 2

        def f(_, y_local):
            u, v = y_local.reshape(nf, -1)
            ul = halo_lo[0] if halo_lo is not None else 1.0
            vl = halo_lo[1] if halo_lo is not None else 3.0
            ur = halo_hi[0] if halo_hi is not None else 1.0
            vr = halo_hi[1] if halo_hi is not None else 3.0
            ue = np.concatenate([[ul], u, [ur]])
            ve = np.concatenate([[vl], v, [vr]])
            cu = ue[:-2] - 2 * u + ue[2:]
            cv = ve[:-2] - 2 * v + ve[2:]
            return np.concatenate([1.0 + u**2 * v - 4 * u + c * cu, 3 * u - u**2 * v + c * cv])

        def jd(_, y_local):
            u, v = y_local.reshape(nf, -1)
            return np.concatenate([2 * u * v - 4 - 2 * c, -(u**2) - 2 * c])

        return f, jd
    if kind == "lorenz96":
        (F,) = spec.params

        def f(_, y):
            ext = np.concatenate([halo_lo, y, halo_hi])      # [i0-2, i0-1, local..., i0+cols]
            return (ext[3:] - ext[:-3]) * ext[1:-2] - y + F

        def jd(_, y):
            return -np.ones_like(y)

        return f, jd
    raise KeyError(kind)


def _worker(rank, world, port, kind):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tornadox_b200.engine import HaloExchanger, local_slices, shard_range

        spec, prob = _spec_and_problem(kind)
        n, d = NU + 1, prob.y0.shape[0]
        rng = np.random.default_rng(7)
        mean = rng.standard_normal((n, d))
        mean[0] = prob.y0
        sc = 1e-2 * rng.standard_normal((d, n, n))
        dt = 1e-3
        abstol, reltol = 1e-4, 1e-2

        # unsharded oracle (every rank computes it to check its own slice)
        ref, _ = ofilters.diagonal_ek1_attempt(ofilters.FilterState(0.0, mean, sc, None, None), dt, prob.f,
                                                prob.df_diagonal, NU)
        rule = odriver.AdaptiveSteps(abstol, reltol)
        ref_norm = rule.scale_error_estimate(dt * ref.error_estimate, ref.reference_state)

        begin, end = shard_range(spec, rank, world)
        idx = np.concatenate([np.arange(d)[s] for s in local_slices(spec, begin, end)])
        mean_l, sc_l = mean[:, idx], sc[idx]

        # predicted state of the shard edges -> neighbours (tdx_pack_halo + HaloExchanger on GPUs)
        p, pinv = oprior.nordsieck_raw(NU, dt)
        m_at = p[0] * (oprior.transition_1d(NU) @ (pinv[:, None] * mean_l))[0]
        nf = spec.n_fields
        per_field = m_at.reshape(nf, -1)
        first_n, last_n = {"fhn_2d": (spec.grid_cols,) * 2, "brusselator": (1, 1), "lorenz96": (1, 2)}[kind]
        edge_first = torch.from_numpy(np.ascontiguousarray(per_field[:, :first_n].reshape(-1)))
        edge_last = torch.from_numpy(np.ascontiguousarray(per_field[:, per_field.shape[1] - last_n:].reshape(-1)))
        cyclic = kind == "lorenz96"
        ex = HaloExchanger(None, rank, world, cyclic, last_n * nf, first_n * nf, torch.float64, "cpu")
        halo_lo, halo_hi = ex.exchange(edge_first, edge_last)
        halo_lo = halo_lo.numpy() if halo_lo is not None else None
        halo_hi = halo_hi.numpy() if halo_hi is not None else None

        f_l, jd_l = _local_f_with_halos(kind, spec, prob, begin, end, halo_lo, halo_hi)
        out, _ = ofilters.diagonal_ek1_attempt(ofilters.FilterState(0.0, mean_l, sc_l, None, None), dt, f_l, jd_l, NU)
        np.testing.assert_allclose(out.mean, ref.mean[:, idx], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(out.cov_sqrtm, ref.cov_sqrtm[idx], rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(out.error_estimate, ref.error_estimate[idx], rtol=1e-11, atol=1e-13)

        # global scaled error norm from the all-reduced local sums -> identical accept/reject everywhere
        ratio = dt * out.error_estimate / (abstol + reltol * out.reference_state)
        sq = torch.tensor([float(np.sum(ratio**2))], dtype=torch.float64)
        dist.all_reduce(sq)
        norm = float(np.sqrt(sq.item() / d))
        assert abs(norm - ref_norm) <= 1e-12 * abs(ref_norm)
        gathered = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(gathered, torch.tensor([norm], dtype=torch.float64))
        assert all(float(x) == norm for x in gathered), "ranks disagree on the error norm"
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["fhn_2d", "brusselator", "lorenz96"])
def test_two_rank_sharded_step_matches_unsharded_oracle(kind):
    mp.spawn(_worker, args=(2, _free_port(), kind), nprocs=2, join=True)


def test_shard_ranges_cover_the_axis():
    from tornadox_b200 import ivp as pivp
    from tornadox_b200.engine import local_slices, shard_range

    spec = pivp.fhn_2d(dx=1.0 / 10, y0=np.zeros(200)).spec
    for world in (1, 2, 3, 4):
        ranges = [shard_range(spec, r, world) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == 10
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        covered = np.concatenate([np.arange(200)[s] for r in ranges for s in local_slices(spec, *r)])
        assert sorted(covered.tolist()) == list(range(200))
    with pytest.raises(ValueError):
        shard_range(pivp.vanderpol().spec, 0, 2)
