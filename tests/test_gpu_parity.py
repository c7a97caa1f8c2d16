"""GPU parity tests: the CUDA engine (through the C ABI / ctypes host mirror) against the CPU oracle on
identical frames.  Run on the B200 box with `pytest -m gpu`."""
import hashlib

import numpy as np
import pytest

from helpers import load_golden, match_corner_sets
from parity import compare_all, compare_detections, compare_front_end, compare_homography

pytestmark = pytest.mark.gpu

CAM = (905.495617, 609.916016, 907.909470, 352.682645)
DIST = (0.059238, -0.075154, -0.003801, 0.001113, 0.0)


@pytest.fixture(scope="module")
def D():
    from ros_vision_b200 import detector
    detector.load_library()
    return detector


def _pack(gray, fmt, rng=None):
    from ros_vision_b200 import synth
    if fmt == "gray":
        return gray
    if fmt == "yuyv":
        return synth.gray_to_yuyv(gray)
    return synth.gray_to_bgr(gray, rng)


@pytest.mark.parametrize("w,h,fmt,dec,sigma,seed,ntags,side", [
    (640, 480, "gray", 2, 0.0, 11, 4, (60, 140)),
    (640, 480, "yuyv", 2, 0.0, 12, 3, (40, 120)),
    (320, 240, "bgr", 2, 0.0, 13, 2, (40, 90)),
    (320, 240, "gray", 1, 0.0, 14, 3, (30, 70)),
    (320, 240, "bgr", 1, 0.8, 15, 4, (24, 60)),
    (328, 248, "yuyv", 2, 0.0, 16, 2, (40, 80)),     # ragged: not a multiple of the CCL / boundary tile sizes
    (1280, 800, "yuyv", 2, 0.0, 2000, 5, (60, 300)),  # BASELINE config 2 shape
])
def test_every_stage_matches_oracle(D, oracle, w, h, fmt, dec, sigma, seed, ntags, side):
    from ros_vision_b200 import synth
    sc = synth.make_scene(w, h, seed, ntags, side_range=side, noise_sigma=4.0)
    frame = _pack(sc.gray, fmt, np.random.default_rng(seed))
    orc = oracle.detect(oracle.make_config(w, h, fmt, dec, sigma), frame)
    det = D.GpuDetector(w, h, fmt, quad_decimate=dec, quad_sigma=sigma, keep_stages=True)
    det.Detect(frame)
    got = compare_all(det, orc, 0, fmt)
    assert len(got) >= 1
    truth = {t.tag_id: t.corners for t in sc.tags}
    for d in got:
        assert int(d["id"]) in truth
        assert match_corner_sets(d["p"], truth[int(d["id"])]) < 1.0
    det.close()


@pytest.mark.parametrize("name", ["ref_colorimage_crop", "ref_colorimage_notags_crop", "ref_grayimage_crop", "synthetic_cfg1"])
def test_golden_fixtures(D, oracle, name):
    """The reference's known answers (gpu_detector_test.cu:84-157) through the CUDA path."""
    meta, img = load_golden(name)
    w, h = meta["width"], meta["height"]
    cases = meta.get("cases") or {"identity_camera": {"camera": None, "dist": None, "oracle": meta["oracle"]}}
    for case in cases.values():
        det = D.GpuDetector(w, h, "gray", camera_matrix=case["camera"], distortion_coefficients=case["dist"], keep_stages=True)
        det.Detect(img)
        exp = case["oracle"]
        assert hashlib.sha256(det.CopyThresholdedTo().tobytes()).hexdigest() == exp["thresh_sha256"]
        info = det.FrameInfo()
        assert info.num_points == exp["num_points"] and info.num_clusters == exp["num_clusters"]
        assert info.num_blobs == exp["num_selected_clusters"] and info.num_quads == exp["num_corners"]
        got = det.Detections()
        assert [int(x) for x in got["id"]] == [d["id"] for d in exp["detections"]]
        assert [int(x) for x in got["hamming"]] == [d["hamming"] for d in exp["detections"]]
        for g, e in zip(got, exp["detections"]):
            assert np.abs(g["p"] - np.array(e["p"])).max() <= 0.05
            compare_homography(g["H"], e["H"])
        if "known_answer" in meta:
            assert len(got) == meta["known_answer"]["num_detections"]
        orc = oracle.detect(oracle.make_config(w, h, "gray", 2, 0.0, camera=case["camera"], dist=case["dist"]), img)
        compare_all(det, orc, 0, "gray")
        det.close()


def test_batch_matches_single_frames(D, oracle):
    from ros_vision_b200 import synth
    w, h = 640, 480
    frames = [synth.gray_to_yuyv(synth.make_scene(w, h, 100 + i, 1 + i % 3, side_range=(50, 150), noise_sigma=3.0 + i).gray)
              for i in range(5)]
    det = D.GpuDetector(w, h, "yuyv", max_batch=8, keep_stages=True)
    det.DetectBatch(frames)
    for i, fr in enumerate(frames):
        orc = oracle.detect(oracle.make_config(w, h, "yuyv", 2, 0.0), fr)
        compare_all(det, orc, i, "yuyv")
    # the same detector again, fewer frames: state left behind by the previous batch must not leak
    det.DetectBatch(frames[3:])
    for i, fr in enumerate(frames[3:]):
        orc = oracle.detect(oracle.make_config(w, h, "yuyv", 2, 0.0), fr)
        compare_all(det, orc, i, "yuyv")
    det.close()


def test_edge_cases(D, oracle):
    w, h = 64, 48
    for name, img in [("flat", np.full((h, w), 128, np.uint8)), ("black", np.zeros((h, w), np.uint8)),
                      ("white", np.full((h, w), 255, np.uint8)),
                      ("checker", ((np.indices((h, w)).sum(0) % 2) * 255).astype(np.uint8)),
                      ("noise", np.random.default_rng(0).integers(0, 256, (h, w), dtype=np.uint8))]:
        det = D.GpuDetector(w, h, "gray", quad_decimate=1, keep_stages=True)
        det.Detect(img)
        orc = oracle.detect(oracle.make_config(w, h, "gray", 1, 0.0), img)
        compare_all(det, orc, 0, "gray")
        det.close()
    # a tag that fills most of the frame: a blob larger than the in-shared-memory sort capacity
    from ros_vision_b200 import synth
    sc = synth.make_scene(1600, 1200, 5, 1, side_range=(900, 900), max_rot_deg=5, max_tilt_deg=5, noise_sigma=0.0,
                          background=200.0)
    det = D.GpuDetector(1600, 1200, "gray", quad_decimate=1, keep_stages=True)
    det.Detect(sc.gray)
    orc = oracle.detect(oracle.make_config(1600, 1200, "gray", 1, 0.0), sc.gray)
    assert orc.clusters["count"][orc.clusters["selected"] != 0].max() > 4096
    compare_all(det, orc, 0, "gray")
    det.close()


def test_boundary_tile_with_more_points_than_its_list(D, oracle):
    """k_boundary's point list holds two points per pixel of its 64x8 tile.  A comb -- one-pixel teeth whose tips sit on a
    tile's last row -- emits four points at every tip and two everywhere else: 1056 points in a tile, so the production
    kernel (not the B200TAG_TEST_SMALL_CHUNKS variant) takes that tile in two rounds."""
    w, h = 256, 64
    img = np.zeros((h, w), np.uint8)
    img[8:12, :] = 255          # the bar that joins the teeth into one component
    img[12:24, 0::2] = 255      # teeth in the even columns, tips on row 23 = last row of the tiles covering rows 16..23
    orc = oracle.detect(oracle.make_config(w, h, "gray", 1, 0.0), img)
    per_tile = np.bincount((orc.points["by"] // 8) * (w // 64) + orc.points["bx"] // 64)
    assert per_tile.max() > 1024, per_tile.max()
    for batch in (1, 3):
        det = D.GpuDetector(w, h, "gray", quad_decimate=1, keep_stages=True, max_batch=batch)
        if batch > 1:
            det.DetectBatch([img] * batch)
        else:
            det.Detect(img)
        for f in range(batch):
            compare_all(det, orc, f, "gray")
        det.close()


@pytest.mark.parametrize("flags", [1, 2, 3, 4, 7])
def test_fallback_paths_give_identical_results(D, oracle, flags):
    """The rarely taken paths -- points counted straight in the global blob-pair hash (crowded CTA-local table),
    the bitonic angle sort (crowded theta buckets) and a boundary tile's points grouped in several rounds (more points
    than its list holds) -- forced on, every stage still equal to the oracle."""
    from ros_vision_b200 import synth
    for w, h, fmt, dec, seed, ntags, side in [(640, 480, "yuyv", 2, 31, 3, (50, 140)), (1280, 800, "gray", 1, 32, 2, (300, 600))]:
        sc = synth.make_scene(w, h, seed, ntags, side_range=side, noise_sigma=4.0)
        frame = _pack(sc.gray, fmt)
        orc = oracle.detect(oracle.make_config(w, h, fmt, dec, 0.0), frame)
        det = D.GpuDetector(w, h, fmt, quad_decimate=dec, keep_stages=True, test_flags=flags)
        det.Detect(frame)
        assert len(compare_all(det, orc, 0, fmt)) >= 1
        det.close()


def test_overflow_is_reported_not_fatal(D):
    """Deliberately tiny device lists: the frame is flagged (B200TAG_E_OVERFLOW + status bits), nothing crashes,
    and the same detector still gives the exact answer on the next frame when the lists are large enough."""
    from ros_vision_b200 import synth
    w, h = 640, 480
    sc = synth.make_scene(w, h, 41, 3, side_range=(60, 140), noise_sigma=4.0)
    for kw, bit in ((dict(max_points=2000), D.ST_POINTS_OVERFLOW), (dict(max_blobs=8), D.ST_BLOBS_OVERFLOW),
                    (dict(max_detections=1), D.ST_DETS_OVERFLOW)):
        det = D.GpuDetector(w, h, "gray", **kw)
        with pytest.raises(D.B200TagError):
            det.Detect(sc.gray)
        assert det.FrameInfo().status & bit, (kw, det.FrameInfo().status)
        det.close()
    det = D.GpuDetector(w, h, "gray")
    det.Detect(sc.gray)
    assert det.FrameInfo().status == 0 and len(det.Detections()) == 3
    det.close()


def test_blob_overflow_with_huge_blobs(D, oracle):
    """Blob-list overflow on a frame that has blobs of more than 4096 points next to many small ones: the huge
    tier's work-list entries (filled from the back of the small tier's list) may then reach into the range the
    warp-per-blob tier reads; it must skip them instead of fitting 4096+ points into its 192-point buffers."""
    w, h = 1600, 1200
    img = np.full((h, w), 200, np.uint8)
    img[100:1100, 60:760] = 20      # two rectangles with a perimeter of 3400 px: about 6800 boundary points each
    img[100:1100, 840:1540] = 20
    for k in range(40):             # 40 small squares inside them: about 100 boundary points each
        rect, j = divmod(k, 20)
        y, x = 150 + 220 * (j // 5), (60 if rect == 0 else 840) + 60 + 120 * (j % 5)
        img[y:y + 14, x:x + 14] = 230
    for max_blobs in (4, 8, 16):
        det = D.GpuDetector(w, h, "gray", quad_decimate=1, max_blobs=max_blobs)
        with pytest.raises(D.B200TagError):
            det.Detect(img)
        assert det.FrameInfo().status & D.ST_BLOBS_OVERFLOW
        det.close()
    orc = oracle.detect(oracle.make_config(w, h, "gray", 1, 0.0), img)
    cnt = orc.clusters["count"]
    assert (cnt > 4096).sum() >= 2 and (cnt <= 192).sum() >= 20, "the scene must exercise both ends of the work list"
    det = D.GpuDetector(w, h, "gray", quad_decimate=1, keep_stages=True)
    det.Detect(img)
    assert det.FrameInfo().status == 0
    compare_all(det, orc, 0, "gray")
    det.close()


@pytest.mark.parametrize("sigma", [0.8, 1.2, 1.6, 2.6, -0.8, -1.3])
def test_quad_sigma_variants(D, oracle, sigma):
    """The Gaussian quad_sigma filter (upstream image_u8_gaussian_blur; negative = sharpen) for every filter length the
    kernel specialises (3, 5, 7) and the generic one, at decimate 1 and 2, on a frame whose quad image is not a
    multiple of the blur tile."""
    from ros_vision_b200 import synth
    for w, h, fmt, dec in ((360, 248, "gray", 1), (656, 496, "yuyv", 2)):
        sc = synth.make_scene(w, h, 21, 2, side_range=(60, 110), noise_sigma=4.0)
        frame = _pack(sc.gray, fmt)
        orc = oracle.detect(oracle.make_config(w, h, fmt, dec, sigma), frame)
        det = D.GpuDetector(w, h, fmt, quad_decimate=dec, quad_sigma=sigma, keep_stages=True)
        det.Detect(frame)
        compare_all(det, orc, 0, fmt)
        det.close()


def test_other_decimation_factors(D, oracle):
    """quad_decimate 3 and 4 (the reference supports only 2): every stage against the oracle."""
    from ros_vision_b200 import synth
    for dec, w, h in ((3, 960, 720), (4, 1280, 960)):
        sc = synth.make_scene(w, h, 50 + dec, 3, side_range=(150, 300), noise_sigma=3.0)
        for fmt in ("gray", "yuyv"):
            frame = _pack(sc.gray, fmt)
            orc = oracle.detect(oracle.make_config(w, h, fmt, dec, 0.0), frame)
            det = D.GpuDetector(w, h, fmt, quad_decimate=dec, keep_stages=True)
            det.Detect(frame)
            assert len(compare_all(det, orc, 0, fmt)) >= 2
            det.close()


def test_two_detectors_interleaved(D, oracle):
    """Two detectors (two CUDA streams) with batches in flight at the same time do not disturb each other."""
    from ros_vision_b200 import synth
    w, h = 640, 480
    fa = [synth.gray_to_yuyv(synth.make_scene(w, h, 60 + i, 2, side_range=(60, 140), noise_sigma=4.0).gray) for i in range(4)]
    fb = [synth.gray_to_yuyv(synth.make_scene(w, h, 70 + i, 3, side_range=(50, 120), noise_sigma=3.0).gray) for i in range(4)]
    da = D.GpuDetector(w, h, "yuyv", max_batch=4, keep_stages=True)
    db = D.GpuDetector(w, h, "yuyv", max_batch=4, keep_stages=True)
    pa = [D.PinnedBuffer(f.size) for f in fa]
    pb = [D.PinnedBuffer(f.size) for f in fb]
    for buf, f in zip(pa + pb, fa + fb):
        buf.array[:] = f.reshape(-1)
    for _ in range(3):
        da.EnqueuePointers([b.ptr for b in pa])
        db.EnqueuePointers([b.ptr for b in pb])
    da.Finish()
    db.Finish()
    for det, frames in ((da, fa), (db, fb)):
        for i, fr in enumerate(frames):
            compare_all(det, oracle.detect(oracle.make_config(w, h, "yuyv", 2, 0.0), fr), i, "yuyv")
    for b in pa + pb:
        b.close()
    da.close()
    db.close()


def test_host_block_enqueue(D, oracle):
    """b200tag_enqueue_host_block: frames of one pinned allocation, tightly packed and with a padded stride."""
    from ros_vision_b200 import synth
    w, h = 640, 480
    frames = [synth.gray_to_yuyv(synth.make_scene(w, h, 80 + i, 2, side_range=(60, 140), noise_sigma=4.0).gray).reshape(-1)
              for i in range(3)]
    fb = frames[0].size
    for stride in (fb, fb + 4096):
        buf = D.PinnedBuffer(stride * len(frames))
        for i, f in enumerate(frames):
            buf.array[i * stride:i * stride + fb] = f
        det = D.GpuDetector(w, h, "yuyv", max_batch=4, keep_stages=True)
        det.EnqueueHostBlock(buf.ptr, len(frames), 0 if stride == fb else stride)
        det.Finish()
        for i, f in enumerate(frames):
            compare_all(det, oracle.detect(oracle.make_config(w, h, "yuyv", 2, 0.0), f.reshape(h, w * 2)), i, "yuyv")
        det.close()
        buf.close()


def test_largest_coordinates(D, oracle):
    """A 3840x2160 quad image (decimate 1): base coordinates up to 3839 exercise the 12-bit fields of the point,
    segment and sort-key records."""
    from ros_vision_b200 import synth
    w, h = 3840, 2160
    sc = synth.make_scene(w, h, 91, 6, side_range=(200, 500), noise_sigma=2.0)
    det = D.GpuDetector(w, h, "gray", quad_decimate=1, keep_stages=True)
    det.Detect(sc.gray)
    assert det.FrameInfo().status == 0
    orc = oracle.detect(oracle.make_config(w, h, "gray", 1, 0.0), sc.gray)
    compare_front_end(det, orc, 0, "gray")
    got = compare_detections(det, orc, 0)
    assert len(got) >= 5
    quads = det.FitQuads()
    assert np.array_equal(quads["corners"], orc.corners["corners"]), "QuadCorners (bit-exact float)"
    assert float(quads["corners"][:, :, 0].max()) > 2048  # the far side of the frame is really used
    det.close()


def test_random_scenes(D, oracle):
    """Forty random frames -- sizes, formats, decimation, blur, tag family, noise level, tag count and size all drawn
    at random -- every stage against the oracle.  Rare paths (equal angles, crowded buckets, blobs on tile corners) get their
    chance here."""
    from ros_vision_b200 import synth
    import os
    seed, count = int(os.environ.get("B200TAG_SWEEP_SEED", "20261018")), int(os.environ.get("B200TAG_SWEEP_COUNT", "40"))
    rng = np.random.default_rng(seed)   # B200TAG_SWEEP_SEED / _COUNT: a longer sweep with other draws
    checked = 0
    for k in range(count):
        dec = int(rng.choice([1, 2, 2, 2, 3]))
        w = int(rng.integers(12, 60)) * 8 * dec
        h = int(rng.integers(10, 44)) * 8 * dec
        w, h = (w // (4 * dec)) * 4 * dec, (h // (4 * dec)) * 4 * dec
        fmt = str(rng.choice(["gray", "yuyv", "bgr"]))
        if fmt == "yuyv":
            w = (w // 8) * 8
            if (w // dec) % 4:
                continue
        ntags = int(rng.integers(0, 5))
        smax = max(24.0, min(w, h) / 2.5)
        sigma = float(rng.choice([0.0, 0.0, 0.0, 0.8, 1.4, 2.2, -0.9]))
        fam = str(rng.choice(["tag36h11", "tag36h11", "tag25h9", "tag16h5"]))
        sc = synth.make_scene(w, h, 7000 + k + (seed % 100000) * 1000 * (seed != 20261018), ntags, side_range=(min(20.0 * dec, smax), smax),
                              noise_sigma=float(rng.uniform(0.0, 8.0)), clutter=bool(rng.integers(0, 2)),
                              salt_pepper=float(rng.choice([0.0, 0.0, 0.01])), family=fam)
        frame = _pack(sc.gray, fmt, np.random.default_rng(k))
        orc = oracle.detect(oracle.make_config(w, h, fmt, dec, sigma, families=[fam]), frame)
        det = D.GpuDetector(w, h, fmt, quad_decimate=dec, quad_sigma=sigma, keep_stages=True, families=[fam])
        det.Detect(frame)
        compare_all(det, orc, 0, fmt)
        det.close()
        checked += 1
    assert checked >= 0.7 * count


def test_production_variant_matches_oracle(D, oracle):
    """keep_stages = 0 (what the node and bench.py run: the kernel variants without any debug-stage code, no cluster
    extents pass): quads and detections against the oracle, single frames and a batch, with and without CUDA graphs."""
    import os
    from ros_vision_b200 import synth
    w, h = 1280, 800
    frames = [synth.config_frame(2, i)[0] for i in range(6)]
    orcs = [oracle.detect(oracle.make_config(w, h, "yuyv", 2, 0.0), f) for f in frames]
    for no_graph in ("0", "1"):
        os.environ["B200TAG_NO_GRAPH"] = no_graph
        try:
            det = D.GpuDetector(w, h, "yuyv", max_batch=8)
        finally:
            os.environ.pop("B200TAG_NO_GRAPH", None)
        for rounds in range(2):  # the second round replays the captured graph
            det.DetectBatch(frames)
            for i, orc in enumerate(orcs):
                compare_detections(det, orc, i)
                quads = det.FitQuads(i)
                assert np.array_equal(quads["corners"], orc.corners["corners"]), "QuadCorners (bit-exact float)"
                info = det.FrameInfo(i)
                assert info.num_points == len(orc.points) and info.num_blobs == int(orc.clusters["selected"].sum())
                assert info.num_selected_points == len(orc.spoints) and info.num_fit_quads == len(orc.fitquads)
        det.Detect(frames[3])
        compare_detections(det, orcs[3], 0)
        det.close()


def test_detector_driven_from_other_threads(D, oracle):
    """A detector created on one thread and driven from others (a ROS executor does that): every entry point selects
    the detector's own device and restores the caller's (csrc/detector.cu DeviceGuard)."""
    import threading
    from ros_vision_b200 import synth
    sc = synth.make_scene(640, 480, 77, 3, side_range=(60, 130), noise_sigma=3.0)
    want = [int(d["id"]) for d in oracle.detect(oracle.make_config(640, 480, "gray", 2, 0.0), sc.gray).detections]
    det = D.GpuDetector(640, 480, "gray", device=0)
    got, errs = {}, []

    def work(k):
        try:
            for _ in range(5):
                with lock:
                    det.Detect(sc.gray)
                    got[k] = [int(i) for i in det.Detections()["id"]]
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    lock = threading.Lock()   # a detector is not re-entrant (same contract as the reference): one call at a time
    ts = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs
    assert all(v == want for v in got.values()) and len(got) == 4
    det.close()


def test_invalid_configurations_are_rejected(D):
    with pytest.raises(D.B200TagError):
        D.GpuDetector(642, 480, "gray")  # quad image width not a multiple of 4
    with pytest.raises(D.B200TagError):
        D.GpuDetector(640, 480, "gray", max_nmaxima=8)
    det = D.GpuDetector(64, 48, "gray")
    with pytest.raises(ValueError):
        det.Detect(np.zeros((10, 10), np.uint8))
    det.close()


def test_full_size_properties(D, oracle):
    """BASELINE configs 3, 4 and 5 at full size: size-independent properties + oracle on the final answer."""
    from ros_vision_b200 import synth
    for cfg_id in (3, 4, 5):
        frame, fmt, w, h, dec, sigma, sc = synth.config_frame(cfg_id)
        det = D.GpuDetector(w, h, fmt, quad_decimate=dec, quad_sigma=sigma, keep_stages=True)
        det.Detect(frame)
        info = det.FrameInfo()
        assert info.status == 0
        th = det.CopyThresholdedTo()
        labels = det.CopyUnionMarkersTo()
        sizes = det.CopyUnionMarkersSizeTo()
        mask = th.reshape(-1) != 127
        # labels are fixed points of the label map, roots carry the whole pixel count
        assert np.array_equal(labels[labels[mask]], labels[mask])
        assert int(sizes.sum()) == int(mask.sum())
        roots, counts = np.unique(labels[mask], return_counts=True)
        assert np.array_equal(sizes[roots], counts.astype(np.uint32))
        # every pixel has the colour of its root
        assert np.array_equal(th.reshape(-1)[labels[mask]], th.reshape(-1)[mask])
        # idempotence: a second run on the same detector gives the same answer
        first = det.Detections().copy()
        det.Detect(frame)
        assert np.array_equal(first, det.Detections())
        # parity proper, at full size: every stage against the oracle (thresholded image, labels, boundary points,
        # blob extents, sorted points, prefix moments, errors, FitQuads, QuadCorners bit-exact; detections within
        # 0.05 px / 1e-4)
        orc = oracle.detect(oracle.make_config(w, h, fmt, dec, sigma), frame)
        got = compare_all(det, orc, 0, fmt)
        # (detector quality, not parity: most of the rendered tags are found)
        truth_ids = sorted(t.tag_id for t in sc.tags)
        assert len(got) >= 0.9 * len(truth_ids)
        det.close()
