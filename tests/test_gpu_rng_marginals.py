"""Marginal laws of the engine's draw sites (the part of live-mode parity that the draw-tape replay leaves open).

The Philox-mode kernels transform 32-bit words with fp32 MUFU code (``om_u01``, ``om_exp1``, ``om_normal_fast``, the Weibull
via ``sqrtf``, the 24-bit coin, Lemire ``randint``; csrc/om_rng.cuh) and draw the position of a new jam by rejection
(``Sim::sample_jam_position``) where the reference draws from a categorical over 1000 grid points (road.py:185-206).
``om_debug_draws`` / ``om_debug_jam_positions`` run those very device functions; here 2e6 draws per site are tested against
the law the reference draws from (numpy legacy RandomState / ``random``: exact inverse-CDF / polar transforms of uniform
doubles) with Kolmogorov-Smirnov, and the jam sampler against the reference's own probability vector with chi-square.
"""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N = 2_000_000


def _draws(kind, seed=12345, arg=0, n=N):
    import torch
    from openmines_b200 import _lib
    lib = _lib.load()
    out = torch.empty(n, dtype=torch.float64, device="cuda")
    rc = lib.om_debug_draws(kind, seed, arg, n, ctypes.c_void_p(out.data_ptr()), None)
    assert rc == 0
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize("kind,name,cdf", [
    (0, "uniform(0,1]", lambda s: s.uniform(0, 1)),
    (1, "exponential(1)  [np.random.exponential / random.expovariate]", lambda s: s.expon()),
    (2, "normal(0,1)  [np.random.normal]", lambda s: s.norm()),
    (3, "weibull(2)  [np.random.weibull(2)]", lambda s: s.weibull_min(2)),
    (4, "uniform(-0.1,0.1)  [load jitter, truck.py:230]", lambda s: s.uniform(-0.1, 0.2)),
    (6, "random()  [jam coin, road.py:179]", lambda s: s.uniform(0, 1)),
])
def test_draw_site_marginals(kind, name, cdf):
    """KS distance of 2e6 device draws from the exact law: D < 1.95 / sqrt(n) (p ~ 1e-3 for continuous data) -- the 24-bit
    granularity of the uniforms (2^-24 = 6e-8) is far below the resolution of the test (1.4e-3)."""
    from scipy import stats
    x = _draws(kind)
    assert np.isfinite(x).all()
    d = stats.kstest(x, cdf(stats).cdf).statistic
    print("%-70s n=%d  KS D=%.2e  (bound %.2e)  mean %.6f  var %.6f" % (name, len(x), d, 1.95 / np.sqrt(len(x)), x.mean(), x.var()))
    assert d < 1.95 / np.sqrt(len(x)), (name, d)
    # tails the tests of means do not see: extreme quantiles within a few percent
    dist = cdf(stats)
    for q in (1e-4, 1e-3, 0.999, 0.9999):
        emp = np.quantile(x, q)
        assert abs(dist.cdf(emp) - q) < 0.2 * min(q, 1 - q) + 3e-5, (name, q, emp)


def test_randint_is_uniform():
    """RandomDispatcher's randint(0, n) (random_dispatcher.py:16,22,29): chi-square over the n outcomes for n = 2, 5, 20."""
    from scipy import stats
    for n in (2, 5, 20):
        x = _draws(5, seed=99 + n, arg=n).astype(np.int64)
        assert x.min() == 0 and x.max() == n - 1
        cnt = np.bincount(x, minlength=n)
        p = stats.chisquare(cnt).pvalue
        print("randint(0,%d): chi2 p=%.3f" % (n, p))
        assert p > 1e-3, (n, cnt)


def _poke(mine, state, field, i, value, dtype):
    """write element i of a per-truck field of instance 0 into the host copy of the state blob"""
    ly = mine.layout
    off = getattr(ly, field)
    stride = ly.e_stride if field in ("t_time", "t_key", "t_phase") else ly.t_stride
    b = np.array([value], dtype=dtype).tobytes()
    o = off + i * stride
    state[0, o:o + len(b)] = np.frombuffer(b, dtype=np.uint8)


@pytest.mark.parametrize("n_trucks", [2, 5, 20])
def test_jam_position_sampler_vs_reference_categorical(north_pit, n_trucks):
    """N trucks under way on the road charging site -> load site 1 with fixed coverages: 2e5 device draws of the jam
    position against the reference's categorical p_j ~ sum_i N(x_j; coverage_i, 0.1) over x = linspace(0, 1, 1000)
    (road.py:185-206), chi-square over 50 groups of 20 grid points (p > 1e-3); the rejection sampler never gives up."""
    import torch
    from scipy import stats
    import openmines_b200 as om
    from openmines_b200 import _lib
    mine = om.BatchedMine(north_pit, 1)
    mine.reset("NaiveDispatcher", rng_mode="philox", seeds=[4242 + n_trucks])
    mine.run(until=0.5)                       # every truck has its init order; nobody has arrived yet
    now = 3.0
    rng = np.random.default_rng(n_trucks)
    cov = np.sort(rng.uniform(0.02, 0.98, n_trucks))
    if n_trucks == 2:
        cov = np.array([0.05, 0.97])          # both edges of the grid: the truncation of the Gaussians matters
    st = mine.state.cpu().numpy().copy()
    hdr = mine.headers()
    T = mine.pc.T
    trip = (mine.pc.road_distance[0] / mine.pc.truck_speed[0]) * 60
    # Truck.get_route_coverage (truck.py:496-520): (now - journey_start) / ((dist / speed) * 60 + pre_jam_time + repair_time),
    # capped at 1 -- with a jam delay on some trucks and a pending repair time on others
    prejam = np.where(np.arange(T) % 3 == 1, 2.5, 0.0)
    repair = np.where(np.arange(T) % 4 == 2, 7.25, 0.0)
    for i in range(T):
        on = i < n_trucks
        _poke(mine, st, "t_status", i, 2 if on else 3, np.uint8)            # ST_MOVING / ST_WAIT_SHOVEL
        _poke(mine, st, "t_cur", i, 0 if on else 1, np.uint8)
        _poke(mine, st, "t_tgt", i, 1, np.uint8)
        _poke(mine, st, "t_jstart", i, now - cov[i] * (trip + prejam[i] + repair[i]) if on else 0.0, np.float64)
        _poke(mine, st, "t_prejam", i, prejam[i], np.float64)
        _poke(mine, st, "t_repair", i, repair[i], np.float64)
    st[0, mine.layout.hdr:mine.layout.hdr + 8] = np.frombuffer(np.array([now]).tobytes(), dtype=np.uint8)   # InstHdr.now
    mine.state.copy_(torch.from_numpy(st))
    n = 200_000
    out = torch.empty(n, dtype=torch.float64, device=mine.device)
    fb = torch.zeros(1, dtype=torch.int32, device=mine.device)
    cov_dev_t = torch.zeros(T, dtype=torch.float64, device=mine.device)
    rc = _lib.load().om_debug_jam_positions(mine.h, ctypes.c_void_p(mine.state.data_ptr()), 0, 0, n,
                                            ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(fb.data_ptr()),
                                            ctypes.c_void_p(cov_dev_t.data_ptr()), None)
    assert rc == 0, _lib.load().om_last_error(mine.h)
    torch.cuda.synchronize()
    pos = out.cpu().numpy()
    assert int(fb.item()) == 0
    x = np.linspace(0, 1, 1000)
    idx = np.rint(pos * 999).astype(np.int64)
    assert np.array_equal(x[idx], pos)                       # grid values, bit for bit
    # the device's coverage() against the reference formula in fp64 (the device rounds the result to fp32)
    jstart = now - cov * (trip + prejam[:n_trucks] + repair[:n_trucks])
    cov_ref = np.minimum((now - jstart) / ((mine.pc.road_distance[0] / mine.pc.truck_speed[:n_trucks]) * 60 + prejam[:n_trucks]
                                           + repair[:n_trucks]), 1.0)
    got_cov = cov_dev_t.cpu().numpy()
    assert (got_cov[n_trucks:] == -1.0).all()
    assert np.abs(got_cov[:n_trucks] - cov_ref).max() <= 1.2e-7, np.abs(got_cov[:n_trucks] - cov_ref).max()
    cov_dev = got_cov[:n_trucks]
    p = np.zeros_like(x)
    for c in cov_dev:
        p += stats.norm.pdf(x, c, 0.1)
    p /= p.sum()
    obs = np.bincount(idx, minlength=1000).reshape(50, 20).sum(axis=1)
    exp = p.reshape(50, 20).sum(axis=1) * n
    chi = stats.chisquare(obs, exp * obs.sum() / exp.sum())
    print("N=%d trucks: chi2=%.1f (49 dof) p=%.3f, mean pos %.4f vs %.4f" % (n_trucks, chi.statistic, chi.pvalue, pos.mean(), (p * x).sum()))
    assert chi.pvalue > 1e-3
