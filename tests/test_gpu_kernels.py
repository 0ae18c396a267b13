"""T1: each CUDA kernel, called through the C ABI (vstab_k_*), against cv2 4.13.0 on the same
seeded inputs and against the committed golden vectors.  Bit-exact for the integer stages
(ingest, pyramid, warp), identical corner list for GFTT, bit-identical points for LK (north-star
tolerance 0.05 px), identical consensus set and <= 1e-6 px for the similarity fit."""
import cv2
import numpy as np
import pytest

import vstab_b200 as vs
from conftest import render_clip
from oracle import cv_restate as R
from oracle import stabilizer_ref as sr

pytestmark = pytest.mark.gpu

LK_TOL_PX = 0.05          # BASELINE.json north_star
H_TOL_PX = 0.1            # corner reprojection


# ------------------------------------------------------------------ K1 ingest
@pytest.mark.parametrize("H,W,wh", [(720, 1280, 360), (1080, 1920, 360), (2160, 3840, 360), (1080, 1920, 1080),
                                    (480, 854, 360), (250, 333, 100), (97, 131, 91), (1080, 1920, 720)])
def test_ingest_bit_exact(H, W, wh):
    rng = np.random.default_rng(H * 7 + wh)
    src = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    gray, sums = vs.k_ingest(src, wh)
    dw, dh, _ = R.working_size(H, W, wh)
    ref = cv2.cvtColor(cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR), cv2.COLOR_BGR2GRAY)
    assert gray.shape == ref.shape
    assert np.array_equal(gray, ref)
    assert np.array_equal(sums, src.reshape(-1, 3).astype(np.uint64).sum(axis=0))


def test_ingest_golden_and_strided_input(golden):
    f0 = golden["f0"]
    for wh in (180, 120, 100, 360):
        gray, sums = vs.k_ingest(f0, wh)
        assert np.array_equal(gray, golden[f"gray_wh{wh}"])
        assert np.array_equal(sums, golden["sums_f0"])
    # extreme content: saturated frame (largest sums), constant frame
    for val in (0, 255):
        img = np.full((360, 640, 3), val, np.uint8)
        gray, sums = vs.k_ingest(img, 120)
        assert (gray == val).all() and (sums == 360 * 640 * val).all()


# ------------------------------------------------------------------ K2 pyramid
@pytest.mark.parametrize("shape", [(360, 640), (1080, 1920), (91, 173), (100, 177), (2160, 3840)])
def test_pyramid_bit_exact(shape):
    rng = np.random.default_rng(shape[1])
    g = rng.integers(0, 256, shape, dtype=np.uint8)
    ref = g
    for o in vs.k_pyramid(g):
        ref = cv2.pyrDown(ref)
        assert np.array_equal(o, ref)


def test_pyramid_golden(golden):
    outs = vs.k_pyramid(golden["g0"])
    for l in (1, 2, 3):
        assert np.array_equal(outs[l - 1], golden[f"g0_pyr{l}"])


# ------------------------------------------------------------------ K3 GFTT
def _gray(frame, wh):
    H, W = frame.shape[:2]
    dw, dh, _ = R.working_size(H, W, wh)
    return cv2.cvtColor(cv2.resize(frame, (dw, dh), interpolation=cv2.INTER_LINEAR), cv2.COLOR_BGR2GRAY)


@pytest.mark.parametrize("W,H,wh", [(1280, 720, 360), (1920, 1080, 360), (1920, 1080, 1080), (640, 360, 100)])
def test_gftt_identical_corner_list(texture, W, H, wh):
    for f in render_clip(texture, W, H, 2, start=3):
        g = _gray(f, wh)
        md = int(10 * g.shape[0] / 720.0)
        pts, eig = vs.k_gftt(g, 1300, 0.01, md, want_eig=True)
        eref = cv2.cornerMinEigenVal(g, 3, ksize=3)
        ref = cv2.goodFeaturesToTrack(g, 1300, 0.01, md).reshape(-1, 2)
        # min-eigenvalue map: bit-identical but for the f64 running-sum residue of OpenCV's box filter
        assert (eig != eref).mean() < 2e-3
        assert np.abs(eig - eref).max() <= 1e-7 * eref.max()
        assert np.array_equal(pts, ref)       # same corners, same order


def test_gftt_golden_noise_and_flat(golden):
    pts = vs.k_gftt(golden["g0"], 1300, 0.01, int(golden["gftt_min_distance"]))
    assert np.array_equal(pts, golden["corners0"])
    rng = np.random.default_rng(5)
    g = rng.integers(0, 256, (360, 640), dtype=np.uint8)      # dense candidates: ~1/9 of the pixels
    # (30, 1300): the image saturates below 1300 points, so every candidate chunk of the fused top-k kernel is consumed
    for md, n in ((5, 1300), (1, 1300), (12, 400), (30, 1300)):
        pts = vs.k_gftt(g, n, 0.01, md)
        ref = cv2.goodFeaturesToTrack(g, n, 0.01, md).reshape(-1, 2)
        assert np.array_equal(pts, ref)
    flat = np.full((120, 200), 77, np.uint8)                   # no corners at all
    assert len(vs.k_gftt(flat, 1300, 0.01, 1)) == 0


# ------------------------------------------------------------------ K4 LK
def _cv_lk(a, b, pts):
    cur, st, _ = cv2.calcOpticalFlowPyrLK(a, b, pts.reshape(-1, 1, 2), None, winSize=(21, 21), maxLevel=3,
                                          criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 50, 0.01),
                                          flags=0, minEigThreshold=1e-4)
    return cur.reshape(-1, 2), st.reshape(-1)


def test_lk_parity(texture):
    fs = render_clip(texture, 1280, 720, 3)
    gs = [_gray(f, 360) for f in fs]
    pts = cv2.goodFeaturesToTrack(gs[0], 1300, 0.01, 5).reshape(-1, 2)
    # points on and near the border exercise the reflect / zero-derivative / out-of-bounds paths
    extra = np.float32([[2, 3], [637, 2], [5, 357], [638, 358], [0, 0], [320, 0], [639, 359], [0.5, 200.25]])
    pts = np.concatenate([pts[:1292], extra])
    for a, b in ((0, 1), (1, 2), (0, 2), (2, 2)):
        cur, st = _cv_lk(gs[a], gs[b], pts)
        mine, mst = vs.k_lk(gs[a], gs[b], pts)
        assert np.array_equal(st, mst)
        ok = st == 1
        assert np.abs(mine[ok] - cur[ok]).max() <= LK_TOL_PX   # the north-star tolerance ...
        # ... and what the kernel achieves: the same float32 bits (OpenCV's five-chain float accumulation order is replayed)
        assert np.array_equal(mine[ok].view(np.uint32), cur[ok].view(np.uint32))


def test_lk_golden_and_edge_cases(golden):
    mine, mst = vs.k_lk(golden["g0"], golden["g1"], golden["corners0"])
    assert np.array_equal(mst, golden["lk_status"])
    ok = mst == 1
    assert np.array_equal(mine[ok], golden["lk_pts"][ok])        # bit-exact
    # empty point list and a textureless image (min-eigenvalue gate => status 0)
    out, st = vs.k_lk(golden["g0"], golden["g1"], np.zeros((0, 2), np.float32))
    assert len(out) == 0 and len(st) == 0
    flat = np.full((180, 320), 100, np.uint8)
    pts = np.float32([[50, 50], [100, 90]])
    _, st = vs.k_lk(flat, flat, pts)
    _, ref = _cv_lk(flat, flat, pts)
    assert np.array_equal(st, ref) and (st == 0).all()
    # points tracked beyond the border keep their coordinates but lose their status
    g0 = golden["g0"]
    g1 = np.roll(g0, (3, 16), axis=(0, 1))
    pts = np.array([[g0.shape[1] - 6, 100], [g0.shape[1] - 3, 40], [150, g0.shape[0] - 2], [160, 90], [2, 3]], np.float32)
    out, st = vs.k_lk(g0, g1, pts)
    cur, ref = _cv_lk(g0, g1, pts)
    assert np.array_equal(st, ref) and (ref == 0).any() and (ref == 1).any()
    assert np.array_equal(out[ref == 1], cur[ref == 1])


def test_lk_small_image_drops_pyramid_levels(golden):
    """128x96: OpenCV keeps levels 0..2 only (level 3 would not exceed the 21 px window); tiny levels
    also exercise the multiply-reflected padding."""
    for (w, h) in ((128, 96), (177, 100), (120, 91)):
        gray = lambda f: cv2.cvtColor(cv2.resize(f, (w, h), interpolation=cv2.INTER_LINEAR), cv2.COLOR_BGR2GRAY)
        a, b = gray(golden["clip"][3]), gray(golden["clip"][4])
        pts = cv2.goodFeaturesToTrack(a, 1300, 0.01, 1).reshape(-1, 2)
        cur, ref = _cv_lk(a, b, pts)
        out, st = vs.k_lk(a, b, pts)
        assert np.array_equal(st, ref)
        assert np.array_equal(out[ref == 1], cur[ref == 1])


# ------------------------------------------------------------------ K5 fit
def _kill_scale(M, w, h):
    o = sr.StabilizerRef(15, 15, 360)
    o.work_size = (w, h)
    Hm = np.eye(3)
    Hm[:2] = M
    return o._kill_scale(Hm)


def test_fit_matches_opencv_ransac(golden):
    corners = np.array([[0, 0, 1], [640, 0, 1], [0, 360, 1], [640, 360, 1]], float).T
    rng = np.random.default_rng(6)
    for trial in range(30):
        n = int(rng.integers(12, 1300))
        p = np.stack([rng.uniform(0, 640, n), rng.uniform(0, 360, n)], 1).astype(np.float32)
        th = rng.uniform(-0.03, 0.03)
        s = 1.0 + rng.uniform(-0.01, 0.01)
        A = s * np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        q = (p @ A.T + rng.uniform(-8, 8, 2) + rng.normal(0, 0.3, (n, 2))).astype(np.float32)
        bad = rng.random(n) < rng.choice([0, 0.01, 0.05, 0.2, 0.5, 0.8])
        q[bad] += rng.uniform(-40, 40, (bad.sum(), 2)).astype(np.float32)
        st = (rng.random(n) > 0.05).astype(np.uint8)
        thr = 3.0 if trial % 2 == 0 else 5.0
        keep = st == 1
        Mr, inl = cv2.estimateAffinePartial2D(p[keep].reshape(-1, 1, 2), q[keep].reshape(-1, 1, 2), method=cv2.RANSAC,
                                              ransacReprojThreshold=thr)
        M, T, cnt = vs.k_fit(p, q, st, 640, 360, thresh=thr)
        assert cnt == (int(keep.sum()), int(inl.sum()))        # identical consensus size
        assert np.abs(M @ corners - Mr @ corners).max() < 1e-6
        Tr = _kill_scale(Mr, 640, 360)
        assert np.abs((T @ corners)[:2] - (Tr @ corners)[:2]).max() < 1e-6


def test_fit_golden_and_degenerate(golden):
    keep = golden["lk_status"] == 1
    M, T, cnt = vs.k_fit(golden["corners0"], golden["lk_pts"], golden["lk_status"], 320, 180)
    assert cnt == (int(keep.sum()), int(golden["inliers"].sum()))
    assert np.abs(M - golden["M"]).max() < 1e-9
    assert np.abs(T - golden["T"]).max() < 1e-9
    # fewer than 10 tracked points -> identity (src/stabilizer.cpp:215)
    p = np.float32([[i * 10, i * 7] for i in range(9)])
    _, T, cnt = vs.k_fit(p, p + 1, np.ones(9, np.uint8), 320, 180)
    assert np.array_equal(T, np.eye(3)) and cnt == (9, 0)
    # all points identical -> no valid model -> identity
    p = np.zeros((20, 2), np.float32)
    _, T, _ = vs.k_fit(p, p, np.ones(20, np.uint8), 320, 180)
    assert np.array_equal(T, np.eye(3))
    # empty input
    _, T, cnt = vs.k_fit(np.zeros((0, 2), np.float32), np.zeros((0, 2), np.float32), np.zeros(0, np.uint8), 320, 180)
    assert np.array_equal(T, np.eye(3)) and cnt == (0, 0)


# ------------------------------------------------------------------ K7 warp
def _rigid(th, tx, ty):
    return np.array([[np.cos(th), -np.sin(th), tx], [np.sin(th), np.cos(th), ty], [0, 0, 1.0]])


@pytest.mark.parametrize("shape", [(1080, 1920), (360, 642), (97, 131)])
def test_warp_bit_exact(texture, shape):
    rng = np.random.default_rng(shape[1])
    noise = rng.integers(0, 256, (shape[0], shape[1], 3), dtype=np.uint8)
    frame = render_clip(texture, shape[1], shape[0], 1)[0]
    for src in (frame, noise):
        for Hm in (_rigid(0.01, 3.3, -7.7), _rigid(-0.05, 40.2, 11.9), np.eye(3), _rigid(3.0, 500.0, 300.0),
                   _rigid(0.12, -20.5, 30.25), _rigid(-0.3, 100.0, -50.0), _rigid(0.7, 300.0, -200.0),
                   np.array([[0.98, 0.02, 5.5], [-0.015, 1.01, -3.25], [1e-5, -2e-5, 1.0]]),
                   _rigid(0.0, 1e7, 0.0)):                      # everything lands on the border colour
            bd = tuple(0.5 * v for v in cv2.mean(src))
            ref = cv2.warpPerspective(src, Hm, (src.shape[1], src.shape[0]), flags=cv2.INTER_LINEAR,
                                      borderMode=cv2.BORDER_CONSTANT, borderValue=bd)
            bv = [int(np.clip(np.rint(b), 0, 255)) for b in bd[:3]]
            assert np.array_equal(vs.k_warp(src, Hm, bv), ref)


def test_warp_golden(golden):
    f0 = golden["f0"]
    bv = [int(np.clip(np.rint(b), 0, 255)) for b in golden["warp_border"]]
    for i in range(2):
        assert np.array_equal(vs.k_warp(f0, golden[f"warp_H{i}"], bv), golden[f"warp_out{i}"])


# ------------------------------------------------------------------ K14 feathered trail (copyFeathered)
@pytest.mark.parametrize("shape", [(360, 640), (250, 333), (720, 1280)])
def test_copy_feathered_bit_exact(texture, shape):
    """K14 == Stabilizer::copyFeathered over cv2 (src/stabilizer.cpp:1051-1155): warp + 7x7 / 101x101 fixed-point Gaussians
    + fillConvexPoly mask + float blend, byte for byte -- also when the warped polygon leaves the image."""
    h, w = shape
    fs = render_clip(texture, w, h, 2, start=11)
    fg, bg = fs[0], fs[1]
    for th, tx, ty in ((0.0, 0.0, 0.0), (0.02, 7.3, -4.1), (-0.05, -30.5, 12.25), (0.11, 60.0, 45.0)):
        Hm = _rigid(th, tx, ty)
        ref = sr.copy_feathered(fg, bg, Hm)
        assert np.array_equal(vs.k_copy_feathered(fg, bg, Hm), ref)
    # a black trail background (the first call) and a projective H
    Hm = _rigid(0.01, 3.0, 2.0)
    Hm[2, 0] = 1e-5
    z = np.zeros_like(fg)
    assert np.array_equal(vs.k_copy_feathered(fg, z, Hm), sr.copy_feathered(fg, z, Hm))
