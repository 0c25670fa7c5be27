"""GPU: motion compensation (vtmme_mc_host / vtmme_mc_batch) and the bi-prediction helpers through the C ABI, against
the oracle's restatement of InterPrediction::xPredInterBlk and against the golden predictions the reference produced
(tests/golden/mc_golden.npz).  Bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import bindings as B  # noqa: E402
from tests.helpers import mc_cases, oracle_mc  # noqa: E402
from tests.test_golden import GM, iter_mc  # noqa: E402


@pytest.fixture(scope="module")
def ms():
    import vtm_b200
    m = vtm_b200.MotionSearch(0)
    yield m
    m.close()


def test_mc_golden(ms):
    for comp, padded, m, blks, bi, alt, want in iter_mc():
        ms.upload_picture(60 + comp, padded, m)
        got = ms.mc_host(comp, [(60 + comp,) + b for b in blks], bi, 10, alt)
        assert np.array_equal(got, want), (comp, bi, alt)


@pytest.mark.parametrize("comp", [0, 1])
@pytest.mark.parametrize("bi,alt,bd", [(0, 0, 10), (1, 0, 10), (0, 1, 10), (1, 1, 10), (0, 0, 8), (1, 0, 8)])
def test_mc_matches_oracle(ms, oracle_lib, comp, bi, alt, bd):
    W, H, M = 384, 256, 192
    cw, ch = (W, H) if comp == 0 else (W // 2, H // 2)
    rng = np.random.default_rng(100 + comp)
    padded = np.ascontiguousarray(np.pad(rng.integers(0, 1 << bd, (ch, cw), dtype=np.int16), M, mode="edge"))
    ms.upload_picture(62, padded, M)
    blks = mc_cases(200 + 10 * comp + bi, comp, cw, ch, 400, max_mv_pel=60)
    got = ms.mc_host(comp, [(62,) + b for b in blks], bi, bd, alt)
    want = oracle_mc(oracle_lib, comp, padded, M, blks, bi, bd, alt)
    assert np.array_equal(got, want)


def test_mc_ragged_sizes(ms, oracle_lib):
    """Widths/heights that are not powers of two (tile edges of 1..15 samples) and the 4x11 / 4x4 coefficient quirk."""
    W, H, M = 256, 128, 192
    rng = np.random.default_rng(7)
    padded = np.ascontiguousarray(np.pad(rng.integers(0, 1024, (H, W), dtype=np.int16), M, mode="edge"))
    ms.upload_picture(63, padded, M)
    blks = []
    for w, h in [(1, 1), (3, 5), (4, 11), (4, 4), (17, 33), (31, 16), (100, 7), (24, 24), (128, 128), (12, 20)]:
        for mv in [(0, 0), (5, 0), (0, 9), (7, 13), (8, 8), (-19, 250), (33, -7)]:
            blks.append((16, 0 if h == 128 else 8, w, h, mv[0], mv[1]))
    for bi in (0, 1):
        got = ms.mc_host(0, [(63,) + b for b in blks], bi, 10, 0)
        assert np.array_equal(got, oracle_mc(oracle_lib, 0, padded, M, blks, bi, 10, 0))


def test_mc_batch_device_bipred_chain(ms, oracle_lib):
    """Device flavour: two bi=1 predictions -> add_avg -> remove_high_freq, everything resident on the GPU."""
    import torch
    from oracle import bindings as B
    W, H, M = 256, 128, 192
    rng = np.random.default_rng(11)
    p0 = np.ascontiguousarray(np.pad(rng.integers(0, 1024, (H, W), dtype=np.int16), M, mode="edge"))
    p1 = np.ascontiguousarray(np.pad(rng.integers(0, 1024, (H, W), dtype=np.int16), M, mode="edge"))
    ms.upload_picture(64, p0, M)
    ms.upload_picture(65, p1, M)
    blks = mc_cases(12, 0, W, H, 300, sizes=[8, 16, 32, 64])
    n = sum(b[2] * b[3] for b in blks) + 3   # odd tail: exercises the scalar remainder of add_avg
    st = torch.cuda.Stream()
    ms.set_stream(st.cuda_stream)
    with torch.cuda.stream(st):
        d0 = torch.zeros(n, dtype=torch.int16, device="cuda")
        d1 = torch.zeros(n, dtype=torch.int16, device="cuda")
        davg = torch.zeros(n, dtype=torch.int16, device="cuda")
        org_h = rng.integers(0, 1024, n, dtype=np.int16)
        dorg = torch.from_numpy(org_h).cuda()
        st.synchronize()
        ms.mc_batch(0, [(64,) + b for b in blks], d0.data_ptr(), 1)
        ms.mc_batch(0, [(65,) + (b[0], b[1], b[2], b[3], -b[4], -b[5]) for b in blks], d1.data_ptr(), 1)
        ms.add_avg(d0.data_ptr(), d1.data_ptr(), davg.data_ptr(), n)
        ms.remove_high_freq(dorg.data_ptr(), davg.data_ptr(), n, 1)
        ms.synchronize()
        got_avg, got_org = davg.cpu().numpy(), dorg.cpu().numpy()
    ms.set_stream(0)
    w0 = np.concatenate([oracle_mc(oracle_lib, 0, p0, M, blks, 1), np.zeros(3, np.int16)])
    w1 = np.concatenate([oracle_mc(oracle_lib, 0, p1, M, [(b[0], b[1], b[2], b[3], -b[4], -b[5]) for b in blks], 1),
                         np.zeros(3, np.int16)])
    wavg = np.zeros(n, np.int16)
    oracle_lib.vo_add_avg(B.ptr(w0), B.ptr(w1), B.ptr(wavg), n, 10)
    worg = org_h.copy()
    oracle_lib.vo_remove_high_freq(B.ptr(worg), B.ptr(wavg), n, 1, 10)
    assert np.array_equal(got_avg, wavg)
    assert np.array_equal(got_org, worg)


def test_bipred_helpers_golden(ms):
    import torch
    s0, s1 = torch.from_numpy(np.ascontiguousarray(GM["avg_s0"])).cuda(), torch.from_numpy(np.ascontiguousarray(GM["avg_s1"])).cuda()
    d = torch.zeros_like(s0)
    torch.cuda.synchronize()
    ms.add_avg(s0.data_ptr(), s1.data_ptr(), d.data_ptr(), s0.numel())
    ms.synchronize()
    assert np.array_equal(d.cpu().numpy(), GM["avg_out"])
    pred = torch.from_numpy(np.ascontiguousarray(GM["hf_pred"])).cuda()
    for clip in (0, 1):
        t = torch.from_numpy(np.ascontiguousarray(GM["hf_org"])).cuda()
        torch.cuda.synchronize()
        ms.remove_high_freq(t.data_ptr(), pred.data_ptr(), t.numel(), clip)
        ms.synchronize()
        assert np.array_equal(t.cpu().numpy(), GM["hf_out%d" % clip])


def test_bcw_helpers(ms, oracle_lib):
    """BCW forms of the bi-prediction helpers: addWeightedAvg for the five weights, removeWeightHighFreq for the weights a
    search can see (g_BcwWeights and 8 - g_BcwWeights), clipped and unclipped, 8 and 10 bit — against the oracle, which is
    pinned on the reference's own AreaBuf members (tests/test_oracle_vs_ref.py)."""
    import torch
    rng = np.random.default_rng(77)
    n = 4096 + 5
    for bd in (8, 10):
        lo, hi = -8192, ((1 << bd) - 1 << (14 - bd)) - 8192        # range of 14-bit intermediate predictions
        a = rng.integers(lo, hi + 1, n).astype(np.int16)
        b = rng.integers(lo, hi + 1, n).astype(np.int16)
        da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
        for idx in range(5):
            d = torch.zeros_like(da)
            torch.cuda.synchronize()
            ms.add_weighted_avg(da.data_ptr(), db.data_ptr(), d.data_ptr(), n, idx, bd)
            ms.synchronize()
            want = np.zeros(n, np.int16)
            oracle_lib.vo_add_weighted_avg(B.ptr(a), B.ptr(b), B.ptr(want), n, bd, idx)
            assert np.array_equal(d.cpu().numpy(), want), (bd, idx)
        org = rng.integers(0, 1 << bd, n).astype(np.int16)
        pred = rng.integers(0, 1 << bd, n).astype(np.int16)
        dpred = torch.from_numpy(pred).cuda()
        for w in (-2, 3, 5, 10, 4):
            for clip in (0, 1):
                t = torch.from_numpy(org.copy()).cuda()
                torch.cuda.synchronize()
                ms.remove_weight_high_freq(t.data_ptr(), dpred.data_ptr(), n, w, clip, bd)
                ms.synchronize()
                want = org.copy()
                oracle_lib.vo_remove_weight_high_freq(B.ptr(want), B.ptr(pred), n, clip, bd, w)
                assert np.array_equal(t.cpu().numpy(), want), (bd, w, clip)


def test_mc_errors(ms):
    from vtm_b200 import VtmmeError
    rng = np.random.default_rng(3)
    padded = np.ascontiguousarray(np.pad(rng.integers(0, 1024, (64, 64), dtype=np.int16), 192, mode="edge"))
    ms.upload_picture(66, padded, 192)
    with pytest.raises(VtmmeError, match="VTMME_ERR_RANGE"):
        ms.mc_host(0, [(66, 0, 0, 16, 16, -400 * 16, 0)])
    with pytest.raises(VtmmeError, match="VTMME_ERR_NOPIC"):
        ms.mc_host(0, [(9999, 0, 0, 16, 16, 0, 0)])
    with pytest.raises(VtmmeError, match="VTMME_ERR_ARG"):
        ms.mc_host(0, [(66, 0, 0, 200, 16, 0, 0)])
    with pytest.raises(VtmmeError, match="VTMME_ERR_ARG"):
        ms.mc_host(2, [(66, 0, 0, 16, 16, 0, 0)])


def test_cand_sad_matches_oracle(ms, oracle_lib):
    """Template-cost / seed distortion: SAD(org, MC(ref, mv)) per candidate == vo_sad on vo_mc_block's output."""
    from oracle import bindings as B
    from vtm_b200.synth import make_pair
    W, H, M = 256, 128, 192
    cur, ref, _ = make_pair(21, W, H, max_global=6, max_local=8, n_rects=2, sigma=3.0)
    refp = np.ascontiguousarray(np.pad(ref, M, mode="edge"))
    ms.upload_picture(70, cur)
    ms.upload_picture(71, refp, M)
    rng = np.random.default_rng(22)
    jobs, want = [], []
    stride = refp.shape[1]
    for i in range(120):
        w, h = int(rng.choice([4, 8, 16, 32, 64, 128])), int(rng.choice([4, 8, 16, 32, 64, 128]))
        x = int(rng.integers(0, (W - w) // w + 1)) * w
        y = int(rng.integers(0, (H - h) // h + 1)) * h
        ncand = int(rng.integers(1, 17))
        mv = rng.integers(-30 * 16, 30 * 16, (ncand, 2)).astype(np.int32)
        ss = 0
        if i % 3 == 0:   # seeds: integer MVs with the search's row sub-sampling
            mv &= ~15
            ss = oracle_lib.vo_subshift(2, w, h)
        host_org = None
        if i % 4 == 1:   # bi-pred style pattern handed over from host memory
            host_org = np.ascontiguousarray((2 * cur[y:y + h, x:x + w].astype(np.int32) - rng.integers(0, 1024, (h, w))).astype(np.int16))
        jobs.append(dict(curPic=70, refPic=71, x=x, y=y, w=w, h=h, mv=mv, subShift=ss, org=host_org))
        org = host_org if host_org is not None else np.ascontiguousarray(cur[y:y + h, x:x + w])
        sads = []
        for c in range(ncand):
            pred = np.zeros((h, w), np.int16)
            oracle_lib.vo_mc_block(0, B.ptr(refp, (M + y) * stride + M + x), stride, w, h, int(mv[c, 0]), int(mv[c, 1]), 0, 10, 0,
                                   B.ptr(pred), w)
            sads.append(int(oracle_lib.vo_sad(B.ptr(org), w, B.ptr(pred), w, w, h, ss)))
        want.append(sads)
    got = ms.cand_sad(jobs)
    assert got == want
    # one by one (different batch composition, same answers)
    assert ms.cand_sad(jobs[:1]) == want[:1]
