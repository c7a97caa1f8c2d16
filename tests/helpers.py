"""Shared helpers for the parity tests."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_png_gray(path):
    import cv2
    im = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    assert im is not None and im.ndim == 2, path
    return np.ascontiguousarray(im)


def load_golden(name):
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        meta = json.load(f)
    img = load_png_gray(os.path.join(GOLDEN, meta["image"]))
    assert img.shape == (meta["height"], meta["width"])
    return meta, img


def match_corner_sets(a, b):
    """Max distance between two 4-corner sets under the best cyclic shift / reversal."""
    a = np.asarray(a, dtype=np.float64).reshape(4, 2)
    b = np.asarray(b, dtype=np.float64).reshape(4, 2)
    best = np.inf
    for rev in (False, True):
        bb = b[::-1] if rev else b
        for s in range(4):
            d = np.abs(a - np.roll(bb, s, axis=0)).max()
            best = min(best, d)
    return best


def canonical_partition(labels, mask):
    """Relabel `labels` (flat) by first occurrence over mask==True pixels; others -> -1."""
    labels = np.asarray(labels).reshape(-1)
    mask = np.asarray(mask).reshape(-1)
    out = np.full(labels.shape, -1, dtype=np.int64)
    vals = labels[mask]
    _, first_idx, inv = np.unique(vals, return_index=True, return_inverse=True)
    order = np.argsort(np.argsort(first_idx))
    out[mask] = order[inv]
    return out
