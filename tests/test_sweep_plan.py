"""CPU: the frame / lane / slot schedule of CenterPointSweep.infer_many (paddle3d_b200/pipeline.py), which the GPU test
tests/test_gpu_pipeline.py::test_sweep_lanes_match_single_lane exercises end to end."""
import pytest


@pytest.mark.parametrize("lanes", [1, 2, 3, 4])
@pytest.mark.parametrize("n", [0, 1, 2, 5, 9, 16])
def test_sweep_plan_invariants(n, lanes):
    from paddle3d_b200.pipeline import CenterPointSweep
    ops = list(CenterPointSweep.plan(n, lanes))
    submits = [o for o in ops if o[0] == "submit"]
    results = [o for o in ops if o[0] == "result"]
    assert [o[1] for o in submits] == list(range(n)) and [o[1] for o in results] == list(range(n))  # every frame once, in order
    busy = {}
    outstanding = 0
    for kind, frame, lane, slot in ops:
        assert lane == frame % lanes and slot == (frame // lanes) & 1
        if kind == "submit":
            assert (lane, slot) not in busy, "slot resubmitted before its result was read"
            busy[(lane, slot)] = frame
            outstanding += 1
            assert outstanding <= lanes + 1
        else:
            assert busy.pop((lane, slot)) == frame
            outstanding -= 1
    assert not busy
