"""Drop-in for the pip `apriltag` module surface that the reference's offline tools use
(src/extrinsic_calibration/extrinsic_calibration/solver.py:181-200: `apriltag.DetectorOptions(families=...)`,
`apriltag.Detector(options).detect(grey)`, `detector.detection_pose(det, (fx, fy, cx, cy), tag_size)`, and the
`tag_id / corners / center / homography` fields of a detection) -- SURVEY.md section 8 row f4.

    import ros_vision_b200.apriltag_api as apriltag        # instead of `import apriltag`

Detection runs on the GPU through the same C ABI as the ROS node's GpuDetector; there is no CPU fallback.
"""
from __future__ import annotations

import collections

import numpy as np

from . import detector as _D

Detection = collections.namedtuple(
    "Detection", ["tag_family", "tag_id", "hamming", "goodness", "decision_margin", "homography", "center", "corners"])


class DetectorOptions:
    """Same keyword surface as pip apriltag's DetectorOptions; fields the engine has no use for are kept as attributes."""

    def __init__(self, families="tag36h11", border=1, nthreads=4, quad_decimate=1.0, quad_blur=0.0, refine_edges=True,
                 refine_decode=False, refine_pose=False, debug=False, quad_contours=True):
        self.families = families
        self.border = int(border)
        self.nthreads = int(nthreads)
        self.quad_decimate = float(quad_decimate)
        self.quad_sigma = float(quad_blur)
        self.refine_edges = int(refine_edges)
        self.refine_decode = int(refine_decode)
        self.refine_pose = int(refine_pose)
        self.debug = int(debug)
        self.quad_contours = quad_contours


class Detector:
    """`Detector(options).detect(gray)` -> list of Detection, like pip apriltag; one GPU detector per image size."""

    def __init__(self, options: DetectorOptions | None = None, searchpath=None):
        self.options = options or DetectorOptions()
        fams = self.options.families
        fams = fams.replace(",", " ").split() if isinstance(fams, str) else list(fams)
        if fams != ["tag36h11"]:
            raise ValueError(f"only the tag36h11 family is supported (got {fams})")
        dec = self.options.quad_decimate
        if dec < 1 or abs(dec - round(dec)) > 1e-9:
            raise ValueError("quad_decimate must be an integer >= 1")
        self._dets = {}

    def _engine(self, w, h):
        key = (w, h)
        if key not in self._dets:
            dec = int(round(self.options.quad_decimate))
            self._dets[key] = _D.GpuDetector(w, h, "gray", quad_decimate=dec, quad_sigma=self.options.quad_sigma,
                                             refine_edges=bool(self.options.refine_edges))
        return self._dets[key]

    def detect(self, img, return_image=False):
        img = np.ascontiguousarray(img)
        if img.ndim != 2 or img.dtype != np.uint8:
            raise ValueError("detect() takes a single-channel uint8 image")
        h, w = img.shape
        det = self._engine(w, h)
        det.Detect(img)
        out = []
        for d in det.Detections():
            out.append(Detection(b"tag36h11", int(d["id"]), int(d["hamming"]), 0.0, float(d["decision_margin"]),
                                 np.array(d["H"], dtype=np.float64).reshape(3, 3), np.array(d["c"], dtype=np.float64),
                                 np.array(d["p"], dtype=np.float64).reshape(4, 2)))
        if return_image:
            return out, np.zeros_like(img)
        return out

    def detection_pose(self, detection: Detection, camera_params, tag_size=1.0, z_sign=1):
        """(4x4 pose of the tag in the camera frame, initial error, final error), like pip apriltag.  The pose is the
        better of the two local minima of the object-space error (libapriltag's estimate_tag_pose)."""
        fx, fy, cx, cy = camera_params
        rec = np.zeros(1, dtype=_D.DETECTION_DT)
        rec["id"][0] = detection.tag_id
        rec["H"][0] = np.asarray(detection.homography, dtype=np.float64).reshape(9)
        rec["c"][0] = detection.center
        rec["p"][0] = detection.corners
        p = _D.estimate_poses(rec, tag_size, fx, fy, cx, cy)[0]
        pose = np.eye(4)
        pose[:3, :3] = p["R"]
        pose[:3, 3] = p["t"]
        if z_sign < 0:  # pip apriltag's option for a camera looking down -z
            pose[2, :] *= -1
            pose[:, 2] *= -1
        return pose, float(p["err"]), float(p["err"])  # no separate "initial" estimate is kept: both are the final error

    def close(self):
        for d in self._dets.values():
            d.close()
        self._dets = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
