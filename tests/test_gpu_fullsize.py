"""GPU: one whole frame pair at the BASELINE size (1080p, SR=64, all 43,020 CUs) against the compiled reference
(oracle/_ref — VTM's own xPatternSearch + xPatternSearchFracDIF, all host cores) when it travelled with the repository,
else against the oracle port on a bounded sample (>= 2,000 CUs).  Run A (zero predictors) and run B (seeded random
quarter-pel predictors within +-16 px) of SURVEY 8(d)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    import bench
    import vtm_b200
    from vtm_b200.synth import make_pair
    m = vtm_b200.MotionSearch(0)
    cur, ref, _ = make_pair(5, bench.WIDTH, bench.HEIGHT)
    m.upload_picture(300, cur)
    m.upload_picture(301, ref)                     # border replicated on the device
    ncu = m.set_frame_size(bench.WIDTH, bench.HEIGHT)
    yield m, cur, ref, ncu
    m.close()


def _compare(got, want):
    bad = [(i, t) for i, t in want if (int(got["mvQx"][i]), int(got["mvQy"][i]), int(got["intX"][i]), int(got["intY"][i]),
                                       int(got["intSad"][i]), int(got["fracCost"][i])) != t]
    return bad


@pytest.mark.parametrize("run", ["A", "B"])
def test_whole_pair_equals_reference(setup, run):
    import bench
    from vtm_b200 import FrameParams
    from vtm_b200.synth import random_predictors
    m, cur, ref, ncu = setup
    pred = None if run == "A" else random_predictors(41, ncu, 16)
    prm = FrameParams(searchRange=bench.SR, lambdaMotion=bench.LAMBDA, predSpread=0 if run == "A" else 33)
    got = m.search_frames([300], [301], prm, None if pred is None else pred[None])[0]
    _, _, desc, kind, _, want = bench.cpu_reference_rate(cur, ref, 1, os.cpu_count() or 1, want_results=True, pred_q=pred)
    assert len(want) >= 2000, desc
    if kind == "reference":
        assert len(want) == ncu == 43020
    bad = _compare(got, want)
    assert not bad, "%d of %d CUs differ from the %s; first: CU %d want %s got %s" % (
        len(bad), len(want), kind, bad[0][0], bad[0][1], got[bad[0][0]])
