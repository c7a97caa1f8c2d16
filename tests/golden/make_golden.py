#!/usr/bin/env python3
"""Generates the committed golden fixtures under tests/golden/.

Run in the build container (needs /root/reference and cv2); the outputs travel
with the repo, /root/reference does not.

Known answers pinned here:
  * src/apriltags_cuda/test/gpu_detector_test.cu:84-157 -- colorimage.jpg holds
    exactly one tag36h11 detection, colorimage_notags.jpg none, CPU and GPU
    detectors agree on id and on centre/corners within 0.5 px.
  * an independent cross-check of ids/corners with cv2.aruco
    (DICT_APRILTAG_36h11, CORNER_REFINE_APRILTAG, decimate 2): id 554 / none / 585.

Each fixture is a lossless crop of the luma plane (the detector discards chroma,
threshold.cu:21) around the tag, with dimensions divisible by 8, plus a JSON with
the known answer, the cv2 cross-check and the oracle's full-pipeline outputs on
exactly those bytes.
"""
import hashlib
import json
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import pyoracle as po  # noqa: E402
from ros_vision_b200 import synth  # noqa: E402

REF = "/root/reference/src/apriltags_cuda/test/data/"
CAM = (905.495617, 609.916016, 907.909470, 352.682645)  # fx, cx, fy, cy  gpu_detector_test.cu:63-66
DIST = (0.059238, -0.075154, -0.003801, 0.001113, 0.0)  # gpu_detector_test.cu:69-73


def aruco(gray):
    det = cv2.aruco.ArucoDetector(cv2.aruco.getPredefinedDictionary(cv2.aruco.DICT_APRILTAG_36h11))
    p = det.getDetectorParameters()
    p.cornerRefinementMethod = cv2.aruco.CORNER_REFINE_APRILTAG
    p.aprilTagQuadDecimate = 2.0
    det.setDetectorParameters(p)
    c, ids, _ = det.detectMarkers(gray)
    if ids is None:
        return []
    return [{"id": int(i[0]), "corners": np.asarray(cc[0], dtype=float).round(3).tolist()} for i, cc in zip(ids, c)]


def summarize(r):
    return {
        "num_points": int(len(r.points)),
        "num_clusters": int(len(r.clusters)),
        "num_selected_clusters": int(r.clusters["selected"].sum()),
        "num_selected_points": int(len(r.spoints)),
        "num_fitquads": int(len(r.fitquads)),
        "num_valid_fitquads": int((r.fitquads["valid"] != 0).sum()),
        "num_corners": int(len(r.corners)),
        "thresh_sha256": hashlib.sha256(r.thresh.tobytes()).hexdigest(),
        "labels_sha256": hashlib.sha256(r.labels.tobytes()).hexdigest(),
        "num_components": int((r.sizes > 0).sum()),
        "detections": [
            {"id": int(d["id"]), "hamming": int(d["hamming"]), "decision_margin": float(d["decision_margin"]),
             "c": d["c"].tolist(), "p": d["p"].tolist(), "H": d["H"].tolist()} for d in r.detections
        ],
    }


def crop_fixture(name, luma, x0, y0, w, h, known, camera=None, dist=None):
    crop = np.ascontiguousarray(luma[y0:y0 + h, x0:x0 + w])
    assert w % 8 == 0 and h % 8 == 0
    png = f"{name}.png"
    cv2.imwrite(os.path.join(HERE, png), crop, [cv2.IMWRITE_PNG_COMPRESSION, 9])
    out = {"image": png, "width": w, "height": h, "format": "gray", "crop_origin": [x0, y0], "known_answer": known,
           "image_sha256": hashlib.sha256(crop.tobytes()).hexdigest(), "cases": {}}
    out["cv2_aruco"] = aruco(crop)
    # case A: identity camera (un/redistort are the identity)
    r = po.detect(po.make_config(w, h, "gray", 2, 0.0), crop)
    out["cases"]["identity_camera"] = {"camera": None, "dist": None, "oracle": summarize(r)}
    # case B: the test's intrinsics, principal point shifted into the crop
    if camera is not None:
        cam = (camera[0], camera[1] - x0, camera[2], camera[3] - y0)
        r = po.detect(po.make_config(w, h, "gray", 2, 0.0, camera=cam, dist=dist), crop)
        out["cases"]["test_camera"] = {"camera": list(cam), "dist": list(dist), "oracle": summarize(r)}
    # the same bytes as YUYV must give the same answer
    r2 = po.detect(po.make_config(w, h, "yuyv", 2, 0.0), synth.gray_to_yuyv(crop))
    assert summarize(r2) == out["cases"]["identity_camera"]["oracle"]
    with open(os.path.join(HERE, f"{name}.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(name, out["cv2_aruco"], [(d["id"], d["hamming"]) for d in out["cases"]["identity_camera"]["oracle"]["detections"]])


def full_frame_answers():
    """Full 1920x1080 frames cannot be committed small; record the oracle's answers on them so the
    known answers (1 tag / 0 tags) are documented as having been checked at full size."""
    res = {}
    for name in ["colorimage.jpg", "colorimage_notags.jpg"]:
        bgr = cv2.imread(REF + name)
        yuyv = cv2.cvtColor(bgr, cv2.COLOR_BGR2YUV_YUYV)
        assert np.array_equal(synth.bgr_to_luma(bgr), yuyv[:, :, 0]), "A.1 luma formula vs cv2"
        H, W = bgr.shape[:2]
        r = po.detect(po.make_config(W, H, "yuyv", 2, 0.0, camera=CAM, dist=DIST), yuyv)
        rb = po.detect(po.make_config(W, H, "bgr", 2, 0.0, camera=CAM, dist=DIST), bgr)
        assert summarize(r) == summarize(rb)
        res[name] = {"width": W, "height": H, "oracle": summarize(r),
                     "cv2_aruco": aruco(cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY))}
    g = cv2.imread(REF + "grayimage.jpg", cv2.IMREAD_UNCHANGED)
    r = po.detect(po.make_config(1280, 800, "gray", 2, 0.0), g)
    res["grayimage.jpg"] = {"width": 1280, "height": 800, "oracle": summarize(r), "cv2_aruco": aruco(g)}
    with open(os.path.join(HERE, "reference_full_frames.json"), "w") as f:
        json.dump(res, f, indent=1)
    for k, v in res.items():
        print(k, [d["id"] for d in v["oracle"]["detections"]], [d["id"] for d in v["cv2_aruco"]])


def synthetic_fixture():
    """BASELINE config 1 (640x480 gray, 4 tags, decimate 2): frame + oracle outputs."""
    fr, fmt, W, H, dec, sig, sc = synth.config_frame(1)
    cv2.imwrite(os.path.join(HERE, "synthetic_cfg1.png"), fr, [cv2.IMWRITE_PNG_COMPRESSION, 9])
    r = po.detect(po.make_config(W, H, fmt, dec, sig), fr)
    out = {"image": "synthetic_cfg1.png", "width": W, "height": H, "format": fmt, "quad_decimate": dec,
           "image_sha256": hashlib.sha256(fr.tobytes()).hexdigest(),
           "truth": [{"id": t.tag_id, "corners": t.corners.tolist()} for t in sc.tags], "oracle": summarize(r)}
    with open(os.path.join(HERE, "synthetic_cfg1.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("synthetic_cfg1", [d["id"] for d in out["oracle"]["detections"]])


if __name__ == "__main__":
    bgr = cv2.imread(REF + "colorimage.jpg")
    luma = synth.bgr_to_luma(bgr)
    crop_fixture("ref_colorimage_crop", luma, 480, 232, 640, 480, {"num_detections": 1, "source": "gpu_detector_test.cu:91,138"},
                 CAM, DIST)
    bgr = cv2.imread(REF + "colorimage_notags.jpg")
    luma = synth.bgr_to_luma(bgr)
    crop_fixture("ref_colorimage_notags_crop", luma, 480, 232, 640, 480,
                 {"num_detections": 0, "source": "gpu_detector_test.cu:101,119"}, CAM, DIST)
    g = cv2.imread(REF + "grayimage.jpg", cv2.IMREAD_UNCHANGED)
    crop_fixture("ref_grayimage_crop", g, 320, 96, 640, 480, {"num_detections": 1, "source": "cv2.aruco cross-check only"})
    full_frame_answers()
    synthetic_fixture()
