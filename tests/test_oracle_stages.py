"""Independent (numpy / scipy) checks of the oracle's stage semantics, SURVEY.md Appendix A."""
import numpy as np
import pytest
from scipy import ndimage

from helpers import canonical_partition
from ros_vision_b200 import synth


def _np_threshold(im, mwbd=5):
    """threshold.cu:60-147 restated with array ops."""
    h, w = im.shape
    t = im.reshape(h // 4, 4, w // 4, 4)
    mn = t.min(axis=(1, 3)).astype(np.int32)
    mx = t.max(axis=(1, 3)).astype(np.int32)
    mnp = np.pad(mn, 1, constant_values=255)
    mxp = np.pad(mx, 1, constant_values=0)
    fmn = np.full_like(mn, 255)
    fmx = np.zeros_like(mx)
    for dy in range(3):
        for dx in range(3):
            fmn = np.minimum(fmn, mnp[dy:dy + h // 4, dx:dx + w // 4])
            fmx = np.maximum(fmx, mxp[dy:dy + h // 4, dx:dx + w // 4])
    MN = np.repeat(np.repeat(fmn, 4, 0), 4, 1)
    MX = np.repeat(np.repeat(fmx, 4, 0), 4, 1)
    thr = MN + (MX - MN) // 2
    out = np.where(im.astype(np.int32) > thr, 255, 0).astype(np.uint8)
    out[(MX - MN) < mwbd] = 127
    return out, np.stack([fmn, fmx], axis=-1).astype(np.uint8)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_threshold_and_decimate(oracle, seed):
    rng = np.random.default_rng(seed)
    sc = synth.make_scene(320, 240, seed, 2, side_range=(40, 90), noise_sigma=3.0)
    gray = sc.gray.copy()
    gray[:40, :40] = 200  # a flat patch -> 127 region
    gray[100:104, 100:104] = rng.integers(0, 256, (4, 4))
    yuyv = synth.gray_to_yuyv(gray)
    r = oracle.detect(oracle.make_config(320, 240, "yuyv", 2, 0.0, max_stage=oracle.STAGE_THRESHOLD), yuyv)
    assert np.array_equal(r.gray, gray)
    assert np.array_equal(r.quad_im, gray[::2, ::2])
    t, mm = _np_threshold(gray[::2, ::2])
    assert np.array_equal(r.thresh, t)
    assert np.array_equal(r.minmax, mm)
    assert (r.thresh == 127).any() and (r.thresh == 0).any() and (r.thresh == 255).any()


def test_decimate_1_and_blur(oracle):
    sc = synth.make_scene(160, 120, 5, 1, side_range=(40, 60), noise_sigma=5.0)
    r = oracle.detect(oracle.make_config(160, 120, "gray", 1, 0.8, max_stage=oracle.STAGE_THRESHOLD), sc.gray)
    assert np.array_equal(r.gray, sc.gray)
    # image_u8_gaussian_blur, sigma 0.8 -> ksz 3, k = [60, 133, 60]
    k = np.array([60, 133, 60], dtype=np.uint32)
    a = sc.gray.astype(np.uint32)
    hpass = a.copy()
    hpass[:, 1:-2] = (a[:, :-3] * k[0] + a[:, 1:-2] * k[1] + a[:, 2:-1] * k[2]) >> 8
    v = hpass.copy()
    v[1:-2, :] = (hpass[:-3, :] * k[0] + hpass[1:-2, :] * k[1] + hpass[2:-1, :] * k[2]) >> 8
    assert np.array_equal(r.quad_im, v.astype(np.uint8))


@pytest.mark.parametrize("seed", [0, 3])
def test_components_match_scipy(oracle, seed):
    sc = synth.make_scene(256, 192, seed, 2, side_range=(40, 80), noise_sigma=4.0)
    r = oracle.detect(oracle.make_config(256, 192, "gray", 2, 0.0, max_stage=oracle.STAGE_LABELS), sc.gray)
    t = r.thresh
    lw, _ = ndimage.label(t == 255, structure=np.ones((3, 3)))
    lb, nb = ndimage.label(t == 0, structure=[[0, 1, 0], [1, 1, 1], [0, 1, 0]])
    ref = np.where(t == 255, lw + nb + 1, np.where(t == 0, lb, 0))
    mask = t != 127
    assert np.array_equal(canonical_partition(r.labels, mask), canonical_partition(ref, mask))
    # label == smallest pixel index of the component, size == pixel count at the root
    lab = r.labels.reshape(-1)
    idx = np.arange(lab.size)
    assert np.array_equal(lab[~mask.reshape(-1)], idx[~mask.reshape(-1)])
    roots, counts = np.unique(lab[mask.reshape(-1)], return_counts=True)
    first = {}
    for i in idx[mask.reshape(-1)]:
        first.setdefault(lab[i], i)
    assert all(first[k] == k for k in roots)
    sizes = np.zeros_like(r.sizes)
    sizes[roots] = counts
    assert np.array_equal(sizes, r.sizes)


def _comb(w=256, h=64):
    """One-pixel teeth hanging from a bar, tips on row 23: four boundary points at every tip, two per pixel along the teeth
    (the densest pattern the boundary rule produces over a whole 64x8 tile: 1056 points in 512 pixels)."""
    img = np.zeros((h, w), np.uint8)
    img[8:12, :] = 255
    img[12:24, 0::2] = 255
    return img


@pytest.mark.parametrize("case", ["scene", "comb"])
def test_boundary_points_bruteforce(oracle, case):
    if case == "scene":
        sc = synth.make_scene(128, 96, 7, 1, side_range=(50, 60), noise_sigma=3.0)
        r = oracle.detect(oracle.make_config(128, 96, "gray", 2, 0.0, max_stage=oracle.STAGE_BLOBS), sc.gray)
    else:
        r = oracle.detect(oracle.make_config(256, 64, "gray", 1, 0.0, max_stage=oracle.STAGE_BLOBS), _comb())
        per_tile = np.bincount((r.points["by"] // 8) * 4 + r.points["bx"] // 64)
        assert per_tile.max() == 1056   # 7 rows x 64 px x 2 + 32 tips x 4 + 32 gaps x 1
    t, L, S = r.thresh, r.labels.reshape(r.h, r.w), r.sizes
    pts = set()
    for y in range(1, r.h - 1):
        for x in range(1, r.w - 1):
            v0 = int(t[y, x])
            if v0 == 127 or S[L[y, x]] < 25:
                continue
            for d, (dx, dy) in enumerate([(1, 0), (1, 1), (0, 1), (-1, 1)]):
                if d == 3:
                    vl, v2 = int(t[y, x - 1]), int(t[y + 1, x])
                    if vl != 127 and v2 != 127 and vl != v2 and x != 1 and S[L[y, x - 1]] >= 25 and S[L[y + 1, x]] >= 25:
                        continue
                v1 = int(t[y + dy, x + dx])
                if v0 + v1 != 255 or S[L[y + dy, x + dx]] < 25:
                    continue
                a, b = int(L[y, x]), int(L[y + dy, x + dx])
                pts.add((min(a, b), max(a, b), 2 * x + dx, 2 * y + dy, d, int(v1 > v0)))
    got = set((int(p["rep0"]), int(p["rep1"]), int(p["x"]), int(p["y"]), int(p["dir"]), int(p["b2w"])) for p in r.points)
    assert got == pts and len(r.points) == len(pts)
    # clusters partition the sorted point list
    c = r.clusters
    assert int(c["count"].sum()) == len(r.points)
    assert np.array_equal(c["start"], np.concatenate([[0], np.cumsum(c["count"])[:-1]]))
    # selected points are sorted by (blob, theta, dir, by, bx)
    sp = r.spoints
    key = np.stack([sp["blob"], sp["theta"], sp["dir"], sp["by"], sp["bx"]], axis=1).astype(np.int64)
    order = np.lexsort(key.T[::-1])
    assert np.array_equal(order, np.arange(len(sp)))


def test_cuda_math_emulation_sanity(oracle):
    """Emulated device atan2f/hypotf stay within 2 ulp of libm (bit-exactness vs the GPU is a -m gpu test)."""
    rng = np.random.default_rng(0)
    L = oracle.lib()
    for _ in range(2000):
        y, x = (float(np.float32(v)) for v in rng.normal(0, 100, 2))
        a = L.orc_emul_atan2f(y, x)
        assert abs(a - np.arctan2(np.float32(y), np.float32(x))) <= 3e-7 * max(1.0, abs(a))
        hgot = L.orc_emul_hypotf(y, x)
        assert abs(hgot - np.hypot(np.float64(y), np.float64(x))) <= 2e-7 * hgot + 1e-30
    assert L.orc_emul_hypotf(3.0, 4.0) == 5.0 and L.orc_emul_hypotf(0.0, 7.0) == 7.0
    assert L.orc_emul_atan2f(0.0, -1.0) == np.float32(np.pi)


def test_decode_codeword_rotations(oracle):
    L = oracle.lib()
    import ctypes as C
    from ros_vision_b200.tag36h11 import CODES
    rot = lambda w: ((w << 9) | (w >> 27)) & ((1 << 36) - 1)
    for tid in (0, 7, 554, 586):
        code = CODES[tid]
        assert L.orc_tag36h11_code(tid) == code
        for k in range(4):
            h, r = C.c_int(), C.c_int()
            # a code observed rotated by k quarter turns decodes after (4-k)%4 further rotations
            w = code
            for _ in range(k):
                w = rot(w)
            assert L.orc_decode_codeword(w ^ 0b101, C.byref(h), C.byref(r)) == tid
            assert h.value == 2 and r.value == (4 - k) % 4
    h, r = C.c_int(), C.c_int()
    assert L.orc_decode_codeword(CODES[0] ^ 0b111, C.byref(h), C.byref(r)) == -1 and h.value == 255


def test_undistort_redistort_roundtrip(oracle):
    import ctypes as C
    cfg = oracle.make_config(1920, 1080, camera=(905.495617, 609.916016, 907.909470, 352.682645),
                             dist=(0.059238, -0.075154, -0.003801, 0.001113, 0.0))
    L = oracle.lib()
    for u0, v0 in [(100.0, 100.0), (960.0, 540.0), (1200.0, 700.0)]:
        u, v = C.c_double(u0), C.c_double(v0)
        assert L.orc_undistort(C.byref(u), C.byref(v), C.byref(cfg)) == 1
        L.orc_redistort(C.byref(u), C.byref(v), C.byref(cfg))
        # the reference's UnDistort (apriltag_detect.cu:372) is not the exact inverse of ReDistort
        # (its tangential term differs); they agree to within a pixel at these intrinsics
        assert abs(u.value - u0) < 1.0 and abs(v.value - v0) < 1.0
    ident = oracle.make_config(640, 480)
    u, v = C.c_double(12.25), C.c_double(99.5)
    L.orc_undistort(C.byref(u), C.byref(v), C.byref(ident))
    assert (u.value, v.value) == (12.25, 99.5)


def test_gaussian_blur_against_opencv(oracle):
    """Row x1 (quad_sigma): the blur follows upstream image_u8_gaussian_blur from memory -- an 8-bit integer kernel
    `k[i] = (uint8)(255 * g[i] / sum g)` applied along rows, then columns, each pass `>> 8`.  Its SHAPE is pinned here
    against OpenCV's float Gaussian of the same length and sigma: the integer kernel sums to a little under 256 and the
    shift divides by 256, so a pass darkens by sum(k) / 256; with that gain applied to OpenCV's result the two agree
    within 3 grey levels on every interior pixel, 1 level lower on average (half a level of truncation per pass)."""
    cv2 = pytest.importorskip("cv2")
    from ros_vision_b200 import synth
    sc = synth.make_scene(640, 480, 3, 4, side_range=(40, 100), noise_sigma=5.0)
    for sigma in (0.8, 1.2, 1.6, 2.6):
        r = oracle.detect(oracle.make_config(640, 480, "gray", 1, sigma, max_stage=oracle.STAGE_THRESHOLD), sc.gray)
        ksz = int(4 * sigma)
        ksz += (ksz % 2 == 0)
        g = np.exp(-0.5 * ((np.arange(ksz) - ksz // 2) / sigma) ** 2)
        k = np.floor(255 * g / g.sum()).astype(np.int64)
        gain = (k.sum() / 256.0) ** 2
        ref = cv2.GaussianBlur(sc.gray.astype(np.float64), (ksz, ksz), sigma, borderType=cv2.BORDER_REPLICATE) * gain
        # OpenCV's kernel is the normalised float Gaussian; the integer one differs from it by < 1/255 per tap
        d = r.quad_im.astype(np.float64)[ksz:-ksz, ksz:-ksz] - ref[ksz:-ksz, ksz:-ksz]
        assert np.abs(d).max() <= 3.0, (sigma, np.abs(d).max())     # observed 2.1 ... 2.7
        assert -1.3 <= d.mean() <= -0.7, (sigma, d.mean())         # two truncations of half a level each: observed -0.99
        # border rule of convolve(): the first and last ksz/2 pixels of a line are copied, not filtered
        rr = ksz // 2
        assert np.array_equal(r.quad_im[:rr, :rr], sc.gray[:rr, :rr]) and np.array_equal(r.quad_im[-rr:, -rr:], sc.gray[-rr:, -rr:])


def test_point_weight_root_equals_hypotf(oracle):
    """The fit kernels compute the point weight (int)(hypotf(gx, gy) + 1) (apriltag_gpu.cu:644-657) as
    (int)(sqrt_rn((float)(gx^2 + gy^2)) + 1): equal for every pair of pixel differences, against the bit-exact emulation
    of CUDA's hypotf and against the integer square root."""
    import ctypes as C
    import math
    L = oracle.lib()
    for gx in range(0, 256):          # hypotf is even in both arguments
        for gy in range(gx, 256):     # and symmetric
            h = L.orc_emul_hypotf(C.c_float(gx), C.c_float(gy))
            s = gx * gx + gy * gy
            w_ref = int(np.float32(h) + np.float32(1))
            w_new = int(np.float32(np.sqrt(np.float32(s))) + np.float32(1))
            assert w_ref == w_new == math.isqrt(s) + 1, (gx, gy)
