"""Run-level constants of the mapping path (wfmash_b200/csrc/stats_host.cu): the three GNU GSL distribution functions the
reference calls (gsl_cdf_binomial_Q, gsl_ran_hypergeometric_pdf, gsl_cdf_hypergeometric_P; GSL is not vendored in the
reference tree and not installed here, so PARITY IS UNPINNED at that boundary) cross-checked against scipy.stats, and the
integers derived from them against an independent numpy / scipy restatement of map_stats.hpp:56-180 and
computeMap.hpp:234-293."""
import ctypes
import math

import numpy as np
import pytest
from scipy import stats

import wfmash_b200 as wb


def f32(x):
    return np.float32(x)


def j2md(j, k):  # Stat::j2md: float in, double pow, float out
    j = f32(j)
    if j == 0:
        return f32(1.0)
    if j == 1:
        return f32(0.0)
    return f32(1 - math.pow(float(f32(2) * j / (f32(1) + j)), 1.0 / k))


def md2j(d, k):  # Stat::md2j: std::pow(float, int) is evaluated in double, the quotient is rounded to float
    sim = float(f32(1) - f32(d))
    return f32(math.pow(sim, k) / (2 - math.pow(sim, k)))


def md_lower_bound(d, s, k, ci):
    q2 = f32((1.0 - float(f32(ci))) / 2)
    j = md2j(d, k)
    x = max(int(math.ceil(float(f32(s) * j))), 1)
    while x <= s:
        if stats.binom.sf(x - 1, s, float(j)) < float(q2):
            x -= 1
            break
        x += 1
    return j2md(f32(x) / f32(s), k)


def min_hits_relaxed(s, k, pid, ci=0.95):
    pid = f32(pid)
    first = int(math.ceil(1.0 * s * float(md2j(f32(1.0 - float(pid)), k))))
    relaxed = first
    for i in range(first, -1, -1):
        d = j2md(f32(1.0 * i / s), k)
        if f32(1.0 - float(md_lower_bound(d, s, k, ci))) >= pid:
            relaxed = i
        else:
            break
    return relaxed


def test_gsl_restatements_agree_with_scipy():
    L = wb.lib()
    for f in (L.wfb_stat_binomial_Q, L.wfb_stat_hypergeometric_pdf, L.wfb_stat_hypergeometric_P):
        f.restype = ctypes.c_double
    rng = np.random.default_rng(5)
    for _ in range(400):
        n = int(rng.integers(1, 400)); k = int(rng.integers(0, n + 2)); p = float(rng.choice([0.0, 1.0, rng.random(), rng.random() ** 4]))
        got = L.wfb_stat_binomial_Q(k, ctypes.c_double(p), n)
        exp = float(stats.binom.sf(k, n, p))
        assert abs(got - exp) <= 1e-12 + 1e-10 * exp, (k, p, n, got, exp)
    for _ in range(400):
        ss = int(rng.integers(1, 200)); ci = int(rng.integers(0, ss + 1)); y = int(rng.integers(0, ci + 2))
        exp_pdf = float(stats.hypergeom.pmf(y, 2 * ss - ci, ss, ci))
        exp_cdf = float(stats.hypergeom.cdf(y, 2 * ss - ci, ss, ci))
        got_pdf = L.wfb_stat_hypergeometric_pdf(y, ss, ss - ci, ci)
        got_cdf = L.wfb_stat_hypergeometric_P(y, ss, ss - ci, ci)
        assert abs(got_pdf - exp_pdf) <= 1e-13 + 1e-9 * exp_pdf, (y, ss, ci)
        assert abs(got_cdf - exp_cdf) <= 1e-12 + 1e-9 * exp_cdf, (y, ss, ci)


def test_sketch_size_and_minimum_hits():
    # s = 29 / 39 / 59 for p = 0.95 / 0.90 / 0.80 at w = 1000, k = 15
    assert [wb.sketch_size(p, 1000, 15) for p in (0.95, 0.90, 0.80)] == [29, 39, 59]
    for s, k, pid in [(29, 15, 0.95), (39, 15, 0.90), (59, 15, 0.80), (98, 19, 0.70), (25, 15, 0.85), (200, 21, 0.99), (10, 15, 0.5)]:
        assert wb.estimate_minimum_hits_relaxed(s, k, pid) == min_hits_relaxed(s, k, pid), (s, k, pid)
    assert 1 <= wb.estimate_minimum_hits_relaxed(29, 15, 0.95) < 29


def test_l2_relaxed_identity_table():
    for s, k, pid in [(29, 15, 0.95), (39, 15, 0.90), (59, 15, 0.80)]:
        tab = wb.l2_min_shared_relaxed(pid, k, s)
        plain = wb.l2_min_shared(pid, k, s)
        assert tab[0] == 0 and (tab[1:] <= plain[1:]).all() and (tab[1:] >= 0).all()
        for qs in (1, s // 2, s):
            exp = next((v for v in range(qs + 1)
                        if f32(1 - float(md_lower_bound(j2md(f32(1.0 * v / qs), k), qs, k, 0.95))) >= f32(pid) or f32(1 - float(j2md(f32(1.0 * v / qs), k))) >= f32(pid)),
                       qs + 1)
            assert tab[qs] == exp, (s, qs)


def test_sketch_cutoffs_table():
    assert (wb.sketch_cutoffs(29, 15, stage1_top_ani_filter=False) == 1).all()
    for ss, k, delta in [(29, 15, 0.0), (39, 15, 0.0), (20, 15, 0.02)]:
        got = wb.sketch_cutoffs(ss, k, ani_diff=delta)
        min_p = float(f32(1) - f32(0.999))
        exp = np.ones(ss + 1, dtype=np.int32)

        def dist_diff(cmax, ci):
            above = 0.0
            for ymax in range(cmax + 1):
                py = stats.hypergeom.pmf(ymax, 2 * ss - cmax, ss, cmax)
                yc = ymax if delta == 0 else math.floor(float(md2j(f32(float(j2md(f32(ymax / ss), k)) + float(f32(delta))), k) * f32(ss)))
                acc = 1 - (stats.hypergeom.cdf(yc - 1, 2 * ss - ci, ss, ci) if yc - 1 >= 0 else 0.0)
                above += py * acc
                if above > min_p:
                    return True
            return above > min_p

        for cmax in range(1, ss + 1):
            lo, n = 0, ss  # std::upper_bound over [0, ss) with a comparator that only looks at the element
            while n > 0:
                half = n // 2
                if dist_diff(cmax, lo + half):
                    n = half
                else:
                    lo += half + 1
                    n -= half + 1
            exp[cmax] = lo if lo != 0 else 1
        assert got.tolist() == exp.tolist(), (ss, k, delta)
        assert got[ss] > 1 and (np.diff(got[1:]) >= 0).all()


@pytest.mark.ref
def test_minimum_hits_and_bounds_match_compiled_reference_live():
    """Stat::estimateMinimumHitsRelaxed / md_lower_bound / j2md / md2j of the reference's UNMODIFIED map_stats.hpp
    (oracle/_ref/libstatsref.so; the two GSL cdf calls it makes are defined in the driver from the textbook formulas in long
    double, independently of the product's restatement)."""
    from tests import util
    R = util.load_ref("libstatsref.so")
    if R is None:
        pytest.skip("oracle/_ref not built (reference sources absent)")
    R.ref_md_lower_bound.restype = ctypes.c_float
    n = 0
    for k in (15, 19, 21):
        for s in (10, 17, 29, 39, 59, 98, 200, 333):
            for pid in (0.70, 0.80, 0.85, 0.90, 0.95, 0.99):
                assert wb.estimate_minimum_hits_relaxed(s, k, pid) == R.ref_estimate_minimum_hits_relaxed(s, k, ctypes.c_float(pid), ctypes.c_float(0.95)), (s, k, pid)
                n += 1
    assert n == 144
    # the keep_low_pct_id table against the reference's own md_lower_bound
    for s, k, pid in [(29, 15, 0.95), (59, 15, 0.80)]:
        tab = wb.l2_min_shared_relaxed(pid, k, s)
        for qs in (1, 7, s):
            exp = qs + 1
            for v in range(qs + 1):
                md = j2md(f32(1.0 * v / qs), k)
                ub = f32(1 - float(R.ref_md_lower_bound(ctypes.c_float(float(md)), qs, k, ctypes.c_float(0.95))))
                if ub >= f32(pid) or f32(1 - float(md)) >= f32(pid):
                    exp = v
                    break
            assert tab[qs] == exp, (s, qs)
