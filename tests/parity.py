"""Stage-by-stage comparison of the CUDA engine (through the C ABI) with the CPU oracle.

Bars (BASELINE.json north_star): thresholded image, component labels, boundary points, blob
extents, point order, moments, fit quads and quad corners bit-exact; tag ids and hamming
bit-exact; detection corners within 0.05 px and homographies within 1e-4 relative (the refine
stage calls single-precision sin/cos/atan2 from the host libm in the reference and from the
device libm here).
"""
from __future__ import annotations

import numpy as np

from ros_vision_b200 import detector as D

CORNER_TOL_PX = 0.05
H_REL_TOL = 1e-4


def _key64(rep0, rep1):
    return (np.asarray(rep0, dtype=np.uint64) << np.uint64(32)) | np.asarray(rep1, dtype=np.uint64)


def compare_front_end(det: D.GpuDetector, orc, frame=0, fmt="yuyv"):
    if fmt != "gray":
        assert np.array_equal(det.CopyGrayTo(frame), orc.gray), "gray image"
    assert np.array_equal(det.CopyDecimatedTo(frame), orc.quad_im), "quad (decimated/blurred) image"
    assert np.array_equal(det.CopyThresholdedTo(frame), orc.thresh), "thresholded image"
    if det.cfg.keep_stages:
        assert np.array_equal(det.CopyStage(D.STAGE_MINMAX, frame).reshape(orc.minmax.shape), orc.minmax), "tile min/max"
    labels = det.CopyUnionMarkersTo(frame)
    sizes = det.CopyUnionMarkersSizeTo(frame)
    mask = orc.thresh.reshape(-1) != 127
    assert np.array_equal(labels[mask], orc.labels[mask]), "component labels"
    assert np.array_equal(sizes, orc.sizes), "component sizes"
    return labels, sizes


def compare_points_and_blobs(det: D.GpuDetector, orc, frame=0):
    info = det.FrameInfo(frame)
    assert info.status == 0, f"device buffer overflow, status={info.status}"
    assert info.num_points == len(orc.points), (info.num_points, len(orc.points))
    assert info.num_clusters == len(orc.clusters)
    assert info.num_blobs == int(orc.clusters["selected"].sum())
    assert info.num_selected_points == len(orc.spoints)
    clusters = det.CopyStage(D.STAGE_CLUSTERS, frame)
    assert len(clusters) == len(orc.clusters)
    ckey = _key64(clusters["rep0"], clusters["rep1"])
    order = np.argsort(ckey, kind="stable")
    clusters = clusters[order]
    okey = _key64(orc.clusters["rep0"], orc.clusters["rep1"])
    assert np.array_equal(ckey[order], okey), "blob pair keys"
    for a, b in (("min_x", "min_x"), ("min_y", "min_y"), ("max_x", "max_x"), ("max_y", "max_y"), ("count", "count"),
                 ("gx_sum", "gx_sum"), ("gy_sum", "gy_sum"), ("pxgx_plus_pygy_sum", "pxgx_plus_pygy_sum")):
        assert np.array_equal(clusters[a].astype(np.int64), orc.clusters[b].astype(np.int64)), f"extents.{a}"
    assert np.array_equal(clusters["selected"] != 0, orc.clusters["selected"] != 0), "SelectBlobs"
    # boundary points as a set keyed by blob pair
    slot_to_key = {int(s): int(k) for s, k in zip(clusters["slot"], okey)}
    pts = det.CopyStage(D.STAGE_POINTS, frame)
    pkey = np.array([slot_to_key[int(s)] for s in pts["slot"]], dtype=np.uint64) if len(pts) else np.zeros(0, np.uint64)
    got = np.stack([pkey, pts["x"].astype(np.uint64), pts["y"].astype(np.uint64), pts["dir"].astype(np.uint64),
                    pts["black_to_white"].astype(np.uint64)], axis=1) if len(pts) else np.zeros((0, 5), np.uint64)
    ref = np.stack([_key64(orc.points["rep0"], orc.points["rep1"]), orc.points["x"].astype(np.uint64),
                    orc.points["y"].astype(np.uint64), orc.points["dir"].astype(np.uint64),
                    orc.points["b2w"].astype(np.uint64)], axis=1) if len(orc.points) else np.zeros((0, 5), np.uint64)
    got = got[np.lexsort(got.T[::-1])]
    ref = ref[np.lexsort(ref.T[::-1])]
    assert np.array_equal(got, ref), "boundary point set"
    return clusters


def compare_line_fit(det: D.GpuDetector, orc, frame=0):
    """Sorted points, prefix moments, errors, filtered errors -- per blob, bit-exact."""
    blobs = det.CopyStage(D.STAGE_BLOBS, frame)
    keys = det.CopyStage(D.STAGE_SORTED_POINTS, frame)
    lfp = det.CopyStage(D.STAGE_LINE_FIT_POINTS, frame)
    errs = det.CopyStage(D.STAGE_ERRORS, frame)
    filt = det.CopyStage(D.STAGE_FILTERED_ERRORS, frame)
    sel = orc.clusters[orc.clusters["selected"] != 0]
    okey = _key64(sel["rep0"], sel["rep1"])
    # STAGE_BLOBS lists every blob pair within the point-count limits; .selected = passed SelectBlobs
    cand = orc.clusters[(orc.clusters["count"] >= 24) & (orc.clusters["count"] <= 4 * (orc.w + orc.h))]
    assert np.array_equal(np.sort(_key64(blobs["rep0"], blobs["rep1"])), _key64(cand["rep0"], cand["rep1"])), "candidate blob set"
    chosen = blobs[blobs["selected"] != 0]
    bkey = _key64(chosen["rep0"], chosen["rep1"])
    assert np.array_equal(np.sort(bkey), okey), "selected blob set"
    lut = {int(k): i for i, k in enumerate(okey)}
    for b in chosen:
        o = sel[lut[int(_key64(b["rep0"], b["rep1"]))]]
        cnt = int(b["count"])
        assert cnt == int(o["count"])
        for fld in ("min_x", "min_y", "max_x", "max_y", "gx_sum", "gy_sum", "pxgx_plus_pygy_sum"):
            assert int(b[fld]) == int(o[fld]), f"blob extents .{fld}"
        g0, o0 = int(b["offset"]), int(o["sel_start"])
        sp = orc.spoints[o0:o0 + cnt]
        ref_key = ((sp["theta"].astype(np.uint64) << np.uint64(26)) | (sp["dir"].astype(np.uint64) << np.uint64(24)) |
                   (sp["by"].astype(np.uint64) << np.uint64(12)) | sp["bx"].astype(np.uint64))
        assert np.array_equal(keys[g0:g0 + cnt], ref_key), f"angle-sorted points of blob {b['rep0']},{b['rep1']}"
        for fld in ("Mxx", "Myy", "Mxy", "Mx", "My", "W"):
            assert np.array_equal(lfp[fld][g0:g0 + cnt], orc.lfps[fld][o0:o0 + cnt]), f"prefix moment {fld}"
        assert np.array_equal(errs[g0:g0 + cnt].astype(np.float64), orc.errs[o0:o0 + cnt]), "line-fit errors"
        assert np.array_equal(filt[g0:g0 + cnt], orc.filtered_errs[o0:o0 + cnt]), "filtered errors"
    return blobs


def compare_quads(det: D.GpuDetector, orc, blobs, frame=0):
    fq = det.CopyStage(D.STAGE_FIT_QUADS, frame)
    assert len(fq) == len(orc.fitquads), "number of FitQuads"
    gk = _key64(blobs["rep0"][fq["blob_index"]], blobs["rep1"][fq["blob_index"]])
    fq = fq[np.argsort(gk, kind="stable")]
    ok = _key64(orc.fitquads["rep0"], orc.fitquads["rep1"])
    assert np.array_equal(np.sort(gk), ok)
    assert np.array_equal(fq["valid"] != 0, orc.fitquads["valid"] != 0), "FitQuad.valid"
    assert np.array_equal(fq["num_peaks"], orc.fitquads["npeaks"]), "peak counts"
    v = fq["valid"] != 0
    assert np.array_equal(fq["indices"][v], orc.fitquads["indices"][v]), "FitQuad.indices"
    assert np.array_equal(fq["err"][v], orc.fitquads["err"][v]), "FitQuad error"
    for fld in ("Mx", "My", "W", "Mxx", "Myy", "Mxy", "N"):
        assert np.array_equal(fq["moments"][fld][v], orc.fitquads["moments"][fld][v]), f"FitQuad.moments.{fld}"
    quads = det.FitQuads(frame)
    assert len(quads) == len(orc.corners), "number of QuadCorners"
    assert np.array_equal(_key64(quads["rep0"], quads["rep1"]), _key64(orc.corners["rep0"], orc.corners["rep1"]))
    assert np.array_equal(quads["corners"], orc.corners["corners"]), "QuadCorners (bit-exact float)"
    return quads


def compare_detections(det: D.GpuDetector, orc, frame=0):
    got = det.Detections(frame)
    ref = orc.detections
    assert [int(x) for x in got["id"]] == [int(x) for x in ref["id"]], "tag ids"
    assert np.array_equal(got["hamming"], ref["hamming"]), "hamming"
    if len(ref):
        assert np.abs(got["p"] - ref["p"]).max() <= CORNER_TOL_PX, "corners"
        assert np.abs(got["c"] - ref["c"]).max() <= CORNER_TOL_PX, "centres"
        scale = np.abs(ref["H"]).max(axis=1, keepdims=True)
        assert (np.abs(got["H"] - ref["H"]) / scale).max() <= H_REL_TOL, "homography"
        assert np.abs(got["decision_margin"] - ref["decision_margin"]).max() <= 0.05, "decision margin"
    return got


def compare_all(det: D.GpuDetector, orc, frame=0, fmt="yuyv"):
    compare_front_end(det, orc, frame, fmt)
    compare_points_and_blobs(det, orc, frame)
    blobs = compare_line_fit(det, orc, frame)
    compare_quads(det, orc, blobs, frame)
    return compare_detections(det, orc, frame)
