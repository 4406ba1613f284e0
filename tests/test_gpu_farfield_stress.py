"""GPU parity tests for the far-field line kernel on ragged / degenerate inputs: grids that are not a multiple of any
tile size, one line, two depths, three active hierarchy levels, windows of every size, zero / NaN / infinite line
parameters, and nu shards cut at arbitrary pixels (which must reproduce the columns of the full run bit for bit).
The checker is the CPU oracle (oracle/), sized to finish in seconds."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

C_A = 2.99792458e18


@pytest.fixture(scope="module")
def ctx():
    from stardis_b200.device import DeviceContext

    c = DeviceContext(0)
    yield c
    c.set_farfield(True)
    c.close()


def _case(rng, N, Ln, D, lam0, hw_lo, hw_hi, dw_lo=1.5e9, dw_hi=5e9, step=0.01):
    lam = lam0 + step * np.arange(N)
    nus = C_A / lam
    line_nus = np.sort(rng.uniform(nus.min(), nus.max(), Ln))
    dws = rng.uniform(dw_lo, dw_hi, (Ln, D))
    gam = 10.0 ** rng.uniform(6.5, 10.5, (Ln, D))
    return nus, line_nus, dws, gam, 10.0 ** rng.uniform(hw_lo, hw_hi, (Ln, D))


def _run(ctx, nus, line_nus, dws, gam, al, far, shard=None, stats=False):
    from stardis_b200 import _lib as L

    D = dws.shape[1]
    ctx.set_farfield(far)
    ctx.set_atmosphere(np.full(D, 5000.0))
    if shard is None:
        ctx.set_grid(nus)
    else:
        ctx.set_grid(nus, *shard)
    ctx.set_lines(line_nus, al)
    ctx.set_broadening(gam, dws)
    ctx.set_line_stats(stats)
    ctx.calc_alpha_line(0)
    out = ctx.get(L.BUF_ALPHA_LINE)
    st = ctx.line_stats() if stats else None
    ctx.set_line_stats(False)
    return out, st


@pytest.mark.parametrize("N,Ln,D,hw", [(513, 1, 2, (2.0, 3.5)), (1023, 40, 3, (0.8, 4.0)), (70001, 400, 2, (0.8, 6.0)),
                                        (40000, 3000, 4, (0.8, 5.0)), (300007, 300, 2, (1.0, 7.0))])
def test_far_field_vs_oracle_on_ragged_grids(ctx, oracle, N, Ln, D, hw):
    rng = np.random.default_rng(N + Ln)
    nus, line_nus, dws, gam, target_hw = _case(rng, N, Ln, D, 4000.0, *hw)
    al = target_hw * oracle.d_nu(nus) / 20.0 / (gam + dws)
    ref, evals, hist = oracle.calc_alan_entries(D, nus, line_nus, dws, gam, al, with_stats=True)
    far, st = _run(ctx, nus, line_nus, dws, gam, al, True, stats=True)
    assert st["evals"] == evals and np.array_equal(st["region_evals"], hist)
    np.testing.assert_allclose(far, ref, rtol=1e-10, atol=0)
    direct, _ = _run(ctx, nus, line_nus, dws, gam, al, False)
    np.testing.assert_allclose(direct, ref, rtol=1e-10, atol=0)
    np.testing.assert_allclose(far, direct, rtol=2e-11, atol=0)
    assert np.array_equal(far == 0, ref == 0)
    # arbitrary shards, including a single pixel and ranges that start / end inside a tile
    cuts = sorted(set([0, 1, N // 7, N // 7 + 1, N // 2 + 3, (3 * N) // 4, N - 1, N]) | set(rng.integers(0, N, 3).tolist()))
    for p0, p1 in zip(cuts[:-1], cuts[1:]):
        part, _ = _run(ctx, nus, line_nus, dws, gam, al, True, shard=(p0, p1))
        assert np.array_equal(part, far[:, p0:p1]), (p0, p1)


def test_far_field_narrow_cores_and_blue_grid(ctx, oracle):
    """Doppler widths of about one pixel (every line core is a few pixels wide) on a dense blue grid."""
    rng = np.random.default_rng(99)
    N, Ln, D = 60000, 2500, 3
    nus, line_nus, dws, gam, target_hw = _case(rng, N, Ln, D, 3000.0, 0.8, 5.5, dw_lo=2e9, dw_hi=6e9, step=0.01)
    gam = 10.0 ** rng.uniform(5.0, 8.5, (Ln, D))  # small damping: y ~ 1e-4 ... 1e-1, region IV dominates the cores
    al = target_hw * oracle.d_nu(nus) / 20.0 / (gam + dws)
    ref, evals, hist = oracle.calc_alan_entries(D, nus, line_nus, dws, gam, al, with_stats=True)
    far, st = _run(ctx, nus, line_nus, dws, gam, al, True, stats=True)
    assert st["evals"] == evals and np.array_equal(st["region_evals"], hist)
    assert hist[3] > hist[2] > 0  # region IV is exercised more than region III
    np.testing.assert_allclose(far, ref, rtol=1e-10, atol=0)


def test_far_field_with_degenerate_line_parameters(ctx, oracle):
    """Zero / infinite Doppler widths, infinite and huge alpha (empty windows through the int64 wrap), zero alpha and a
    whole-grid line next to ordinary ones: same NaN pattern, same zeros, same values as the oracle in both modes."""
    rng = np.random.default_rng(5)
    N, Ln, D = 30000, 60, 2
    nus, line_nus, dws, gam, target_hw = _case(rng, N, Ln, D, 5000.0, 1.0, 6.0)
    al = target_hw * oracle.d_nu(nus) / 20.0 / (gam + dws)
    dws[3, 0] = 0.0            # division by zero inside the window
    dws[7, 1] = np.inf         # x = 0, K = 0
    al[11, 0] = np.inf         # int() wraps: contributes nothing
    al[12, 1] = 1e40
    al[13, :] = 0.0            # hw forced to 10, contributes zeros
    al[20, :] = 1e30 * al[20, :].clip(min=1e-30)  # window far beyond the grid on both sides
    with np.errstate(all="ignore"):
        ref, evals, hist = oracle.calc_alan_entries(D, nus, line_nus, dws, gam, al, with_stats=True)
    for far_on in (True, False):
        out, st = _run(ctx, nus, line_nus, dws, gam, al, far_on, stats=True)
        assert np.array_equal(np.isnan(out), np.isnan(ref))
        assert np.array_equal(np.isinf(out), np.isinf(ref))
        ok = np.isfinite(ref)
        np.testing.assert_allclose(out[ok], ref[ok], rtol=1e-10, atol=0)
        assert st["evals"] == evals


@pytest.mark.parametrize("kind", ["jump", "log", "jitter"])
def test_far_field_on_non_uniform_grids(ctx, oracle, kind):
    """The far criterion is an index distance, valid only where the tile widths vary smoothly: a grid with a jump in its
    step must lose the affected hierarchy levels (k_level_check) and still agree with the oracle; a logarithmic grid
    (smooth, strongly non-uniform) and a jittered one keep working through the per-pair convergence tests."""
    rng = np.random.default_rng(17)
    N, Ln, D = 50000, 600, 2
    if kind == "jump":      # 0.01 A steps, then 0.06 A steps: tile widths jump by a factor 6 in the middle of the grid
        lam = 4000.0 + np.concatenate([0.01 * np.arange(N // 2), 0.01 * (N // 2) + 0.06 * np.arange(1, N - N // 2 + 1)])
    elif kind == "log":     # constant resolving power: the step grows by a factor 2.7 along the grid
        lam = 3000.0 * np.exp(np.arange(N) * 2.0e-5)
    else:                   # 30 % jitter of every step
        lam = 4000.0 + np.cumsum(0.01 * rng.uniform(0.7, 1.3, N))
    nus = C_A / lam
    assert (np.diff(nus) < 0).all()
    line_nus = np.sort(rng.uniform(nus.min(), nus.max(), Ln))
    dws = rng.uniform(1.5e9, 5e9, (Ln, D))
    gam = 10.0 ** rng.uniform(6.5, 10.0, (Ln, D))
    target_hw = 10.0 ** rng.uniform(1.0, 6.5, (Ln, D))
    al = target_hw * oracle.d_nu(nus) / 20.0 / (gam + dws)
    ref, evals, hist = oracle.calc_alan_entries(D, nus, line_nus, dws, gam, al, with_stats=True)
    far, st = _run(ctx, nus, line_nus, dws, gam, al, True, stats=True)
    assert st["evals"] == evals and np.array_equal(st["region_evals"], hist)
    np.testing.assert_allclose(far, ref, rtol=1e-10, atol=0)
    direct, _ = _run(ctx, nus, line_nus, dws, gam, al, False)
    np.testing.assert_allclose(far, direct, rtol=2e-11, atol=0)
    assert np.array_equal(far == 0, ref == 0)
    p0, p1 = N // 3 + 5, (2 * N) // 3 - 7
    part, _ = _run(ctx, nus, line_nus, dws, gam, al, True, shard=(p0, p1))
    assert np.array_equal(part, far[:, p0:p1])
    if kind != "jump":  # the far field really is in use on the smooth grids: most evaluations are not done pixel by pixel
        ctx.set_grid(nus)
        ctx.set_line_stats(True)
        ctx.calc_alpha_line(0)
        ex = ctx.line_stats_ex()
        ctx.set_line_stats(False)
        assert ex["far_replaced_evals"] > 0.5 * evals
        assert ex["multipole_expansions"] > 0 and ex["m2l_row_steps"] > 0
