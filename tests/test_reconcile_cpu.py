"""reconcile_detections as the engine runs it on the host after every frame (csrc/detector.cu: reconcile), through the
test hook b200tag_debug_reconcile -- no GPU needed.  Reference: libapriltag reconcile_detections, declared at
apriltag_detect.cu:31-32 and called at :660 (source not vendored: restated from the published algorithm)."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def D():
    from ros_vision_b200 import build, detector
    build.build_native()
    detector.load_library()
    return detector


def _det(D, id_, hamming, margin, cx, cy, half=20.0, family=0):
    d = np.zeros(1, dtype=D.DETECTION_DT)[0]
    d["id"], d["hamming"], d["decision_margin"], d["family"] = id_, hamming, margin, family
    d["c"] = (cx, cy)
    d["p"] = [(cx - half, cy + half), (cx + half, cy + half), (cx + half, cy - half), (cx - half, cy - half)]
    d["H"] = (half, 0, cx, 0, -half, cy, 0, 0, 1)
    return d


def reconcile(D, dets):
    import ctypes as C
    arr = np.array(dets, dtype=D.DETECTION_DT)
    n = D.load_library().b200tag_debug_reconcile(arr.ctypes.data_as(C.c_void_p), len(arr))
    return arr[:n]


def test_reconcile_overlapping_duplicates(D):
    """reconcile_detections (libapriltag; called at apriltag_detect.cu:660): overlapping detections of the same family
    and id are reduced to one -- lower hamming first, then the larger decision margin; non-overlapping duplicates and
    other ids / families stay; the result is sorted by id."""
    # lower hamming wins, whatever the margins
    out = reconcile(D, [_det(D, 7, 1, 90.0, 100, 100), _det(D, 7, 0, 50.0, 104, 102)])
    assert len(out) == 1 and int(out[0]["hamming"]) == 0 and out[0]["c"][0] == 104
    out = reconcile(D, [_det(D, 7, 0, 50.0, 104, 102), _det(D, 7, 1, 90.0, 100, 100)])
    assert len(out) == 1 and int(out[0]["hamming"]) == 0
    # equal hamming: the larger margin wins
    out = reconcile(D, [_det(D, 7, 1, 60.0, 100, 100), _det(D, 7, 1, 80.0, 103, 100)])
    assert len(out) == 1 and out[0]["decision_margin"] == 80.0
    # one quad inside the other (no edge crossing) still overlaps
    out = reconcile(D, [_det(D, 7, 0, 60.0, 100, 100, half=40), _det(D, 7, 0, 80.0, 100, 100, half=10)])
    assert len(out) == 1 and out[0]["decision_margin"] == 80.0
    # same id far apart: both are real tags
    out = reconcile(D, [_det(D, 7, 0, 60.0, 100, 100), _det(D, 7, 0, 80.0, 400, 100)])
    assert len(out) == 2
    # same place, different ids or families: kept; output sorted by id
    out = reconcile(D, [_det(D, 9, 0, 60.0, 100, 100), _det(D, 7, 0, 80.0, 100, 100), _det(D, 7, 0, 70.0, 100, 100, family=1)])
    assert [int(x) for x in out["id"]] == [7, 7, 9] and sorted(int(x) for x in out["family"][:2]) == [0, 1]
    # a chain: three overlapping copies collapse to the best one
    out = reconcile(D, [_det(D, 3, 2, 99.0, 100, 100), _det(D, 3, 0, 10.0, 102, 100), _det(D, 3, 0, 20.0, 101, 101)])
    assert len(out) == 1 and int(out[0]["hamming"]) == 0 and out[0]["decision_margin"] == 20.0
    # the oracle's reconcile gives the same survivors on a random pile of boxes
    rng = np.random.default_rng(5)
    pile = [_det(D, int(rng.integers(0, 4)), int(rng.integers(0, 3)), float(rng.uniform(10, 200)), float(rng.uniform(50, 300)),
                 float(rng.uniform(50, 300)), half=float(rng.uniform(10, 40))) for _ in range(60)]
    out = reconcile(D, pile)
    assert 0 < len(out) < 60
    for i in range(len(out)):
        for j in range(i + 1, len(out)):
            if int(out[i]["id"]) == int(out[j]["id"]):
                ai, aj = out[i], out[j]
                sep_x = abs(ai["c"][0] - aj["c"][0]) >= (abs(ai["p"][1][0] - ai["c"][0]) + abs(aj["p"][1][0] - aj["c"][0]))
                sep_y = abs(ai["c"][1] - aj["c"][1]) >= (abs(ai["p"][0][1] - ai["c"][1]) + abs(aj["p"][0][1] - aj["c"][1]))
                assert sep_x or sep_y, "two overlapping detections of one id survived"


def test_reconcile_equals_oracle(D, oracle):
    """The product's reconcile (C++) and the oracle's (C, oracle/apriltag_oracle.c orc_i_reconcile) keep the same
    detections on random piles of overlapping boxes."""
    import ctypes as C
    lib = oracle.lib()
    lib.orc_i_reconcile.argtypes = [C.c_void_p, C.c_int]
    lib.orc_i_reconcile.restype = C.c_int
    for seed in range(20):
        rng = np.random.default_rng(100 + seed)
        pile = [_det(D, int(rng.integers(0, 5)), int(rng.integers(0, 3)), float(rng.integers(10, 60)), float(rng.uniform(50, 300)),
                     float(rng.uniform(50, 300)), half=float(rng.uniform(10, 45)), family=int(rng.integers(0, 2))) for _ in range(80)]
        got = reconcile(D, pile)
        ref = np.zeros(len(pile), dtype=oracle.DET_DT)
        for k, d in enumerate(pile):
            for f in ("id", "hamming", "decision_margin", "family", "c", "p", "H"):
                ref[k][f] = d[f]
        # the engine sorts the raw device output first (arbitrary append order); give the oracle the same order
        order = sorted(range(len(pile)), key=lambda k: (int(ref[k]["id"]), float(ref[k]["c"][0]), float(ref[k]["c"][1]),
                                                        int(ref[k]["family"]), int(ref[k]["hamming"])))
        ref = np.ascontiguousarray(ref[order])
        n = lib.orc_i_reconcile(ref.ctypes.data_as(C.c_void_p), len(ref))
        ref = ref[:n]
        key = lambda a: sorted((int(x["id"]), int(x["family"]), float(x["c"][0]), float(x["c"][1]), int(x["hamming"])) for x in a)  # noqa: E731
        assert key(got) == key(ref), seed
