"""GPU: the CUDA path (through the C ABI) against the committed golden vectors generated from the compiled
reference (tests/golden/vtm_golden.npz)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from tests.test_golden import G, iter_dist, iter_interp, iter_search  # noqa: E402


@pytest.fixture(scope="module")
def ms():
    import vtm_b200
    m = vtm_b200.MotionSearch(0)
    yield m
    m.close()


def test_gpu_dist_golden(ms):
    for w, h, kind, org, cur, vals in iter_dist():
        assert ms.dist_host(0, org, cur, 0) == vals[0], (w, h, kind)
        ss = 1 if (h > 8 and w <= 64) else 0          # subShiftMode 2 (RdCost.cpp:310-316)
        assert ms.dist_host(0, org, cur, ss) == vals[1], (w, h, kind, "subshift")
        assert ms.dist_host(1, org, cur) == vals[2], (w, h, kind, "satd")


def test_gpu_interp_golden(ms):
    src, mid = np.ascontiguousarray(G["if_src"]), np.ascontiguousarray(G["if_mid"])
    for comp, w, h, frac, vert, first, last, alt, want in iter_interp():
        s = src if first else mid
        got = ms.interp_host(comp, vert, s, 8 * 40 + 8, 40, w, h, frac, first, last, 10, alt)
        assert np.array_equal(got, want), (comp, w, h, frac, vert, first, last, alt)


def test_gpu_search_golden(ms):
    from vtm_b200 import Job
    plane = np.ascontiguousarray(G["search_ref"])
    ms.upload_picture(50, plane)
    jobs, want = [], []
    for w, h, x, y, win, pq, imv, alt, ssm, lam, org, res in iter_search():
        ss = 1 if (ssm == 2 and h > 8 and w <= 64) else 0
        jobs.append(Job(50, 50, x, y, w, h, win, pq, imv, ss, 10, 1, alt, 1, lam, org))
        want.append(res)
    got = ms.search(jobs)
    assert got == want


def test_gpu_tz_golden(ms):
    """vtmme_search with vtmme_tz against the reference's own xTZSearch + fractional refinement (committed fixtures)."""
    from tests.test_golden import GA_H, GA_W, _ga_planes, iter_tz
    from vtm_b200 import Job, TzSearch
    cur, refp, m = _ga_planes()
    ms.upload_picture(60, cur)
    ms.upload_picture(61, refp, m)
    jobs, want = [], []
    for w, h, x, y, pq, ssm, t, lam, res in iter_tz():
        ss = 1 if (ssm == 2 and h > 8 and w <= 64) else 0
        tz = TzSearch((t.startX, t.startY), t.searchRange, GA_W, GA_H, tuple((t.seedX[i], t.seedY[i]) for i in range(t.nSeeds)),
                      (t.int2Nx2NX, t.int2Nx2NY) if t.hasInt2Nx2N else None, t.extended, t.fast, t.firstSearchStop)
        jobs.append(Job(60, 61, x, y, w, h, (0, 0, 0, 0), pq, 0, ss, 10, 1, 0, 1, lam, None, None, tz))
        want.append(res)
    assert ms.search(jobs) == want
    assert [ms.search([j])[0] for j in jobs] == want


def test_gpu_tz_selective_golden(ms):
    """vtmme_search with vtmme_tz.selective / stagedSad against the reference's own xTZSearchSelective, resp. xTZSearch
    under subShiftMode 1, + fractional refinement (committed fixtures)."""
    from tests.test_golden import GA_H, GA_W, GS, _ga_planes, iter_tz
    from vtm_b200 import Job, TzSearch
    cur, refp, m = _ga_planes()
    ms.upload_picture(60, cur)
    ms.upload_picture(61, refp, m)
    jobs, want = [], []
    for w, h, x, y, pq, ssm, t, lam, res in iter_tz(GS):
        ss = (4 if h > 32 else 3 if h > 16 else 2 if h > 8 else 1) if ssm == 1 else (1 if (ssm == 2 and h > 8 and w <= 64) else 0)
        tz = TzSearch((t.startX, t.startY), t.searchRange, GA_W, GA_H, tuple((t.seedX[i], t.seedY[i]) for i in range(t.nSeeds)),
                      (t.int2Nx2NX, t.int2Nx2NY) if t.hasInt2Nx2N else None, t.extended, t.fast, t.firstSearchStop, 128,
                      t.selective, int(ssm == 1))
        jobs.append(Job(60, 61, x, y, w, h, (0, 0, 0, 0), pq, 0, ss, 10, 1, 0, 1, lam, None, None, tz))
        want.append(res)
    assert ms.search(jobs) == want
    assert [ms.search([j])[0] for j in jobs] == want


def test_gpu_amvr_golden(ms):
    """vtmme_search fracMode 2 against the reference's own xPatternSearch + xPatternSearchIntRefine (committed fixtures)."""
    from tests.test_golden import GA_H, GA_W, _ga_planes, iter_amvr
    from vtm_b200 import Amvr, Job
    cur, refp, m = _ga_planes()
    ms.upload_picture(60, cur)
    ms.upload_picture(61, refp, m)
    jobs, want = [], []
    for w, h, x, y, pq, win, imv, use_had, io, lam, res in iter_amvr():
        am = Amvr(imv, ((io.candX[0], io.candY[0]), (io.candX[1], io.candY[1])), io.numCand, io.mvpIdx,
                  (io.mvpIdxBits[0], io.mvpIdxBits[1]), io.bits, GA_W, GA_H, io.fWeight)
        jobs.append(Job(60, 61, x, y, w, h, win, pq, imv << 1, 0, 10, use_had, 0, 2, lam, None, am))
        want.append(res)
    pick = lambda t: t[:3] + t[8:]
    assert [pick(t) for t in ms.search(jobs)] == want
    assert [pick(ms.search([j])[0]) for j in jobs] == want
