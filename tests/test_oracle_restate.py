"""T0: the numpy restatements of the OpenCV primitives (oracle/cv_restate.py) against cv2
4.13.0 and against the committed golden vectors.  CPU only."""
import cv2
import numpy as np
import pytest

from oracle import cv_restate as R


def test_gray_bit_exact():
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (97, 131, 3), dtype=np.uint8)
    assert np.array_equal(R.bgr2gray(img), cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))


@pytest.mark.parametrize("sh,sw,wh", [(720, 1280, 360), (1080, 1920, 360), (540, 960, 135), (480, 854, 360),
                                      (600, 800, 240), (360, 640, 360), (250, 333, 100)])
def test_resize_linear_bit_exact(sh, sw, wh):
    rng = np.random.default_rng(sh + wh)
    src = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
    dw, dh, _ = R.working_size(sh, sw, wh)
    assert np.array_equal(R.resize_linear_bgr(src, dw, dh), cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR))


def test_golden_gray(golden):
    f0 = golden["f0"]
    for wh in (180, 120, 100, 360):
        dw, dh, _ = R.working_size(f0.shape[0], f0.shape[1], wh)
        assert np.array_equal(R.bgr2gray(R.resize_linear_bgr(f0, dw, dh)), golden[f"gray_wh{wh}"])


@pytest.mark.parametrize("shape", [(360, 640), (180, 320), (45, 80), (91, 173)])
def test_pyrdown_bit_exact(shape):
    rng = np.random.default_rng(shape[0])
    g = rng.integers(0, 256, shape, dtype=np.uint8)
    assert np.array_equal(R.pyr_down(g), cv2.pyrDown(g))


def test_golden_pyramid(golden):
    lv = R.lk_pyramid(golden["g0"])
    for l in (1, 2, 3):
        assert np.array_equal(lv[l], golden[f"g0_pyr{l}"])


def test_warp_bit_exact(golden):
    f0 = golden["f0"]
    bd = tuple(golden["warp_border"])
    assert np.allclose(R.border_value(f0), bd, rtol=0, atol=1e-9)
    for i in range(2):
        out = R.warp_perspective_bgr(f0, golden[f"warp_H{i}"], bd)
        assert np.array_equal(out, golden[f"warp_out{i}"])


def test_warp_weight_table_is_arithmetic():
    """The Q15 weights are (32-ay)(32-ax)*32 etc. except the saturated (0,0) entry."""
    tab = R.warp_bilinear_tab()
    for ay in range(32):
        for ax in range(32):
            w = [(32 - ay) * (32 - ax) * 32, (32 - ay) * ax * 32, ay * (32 - ax) * 32, ay * ax * 32]
            if ax == 0 and ay == 0:
                assert list(tab[ay, ax]) == [32767, 1, 0, 0]
            else:
                assert list(tab[ay, ax]) == w


def test_invert3_matches_cv():
    rng = np.random.default_rng(2)
    for _ in range(20):
        H = np.eye(3) + rng.normal(0, 0.05, (3, 3))
        assert np.array_equal(R.invert3x3(H), cv2.invert(H)[1])


def test_gftt_identical_list(golden):
    g0 = golden["g0"]
    md = int(golden["gftt_min_distance"])
    e = R.corner_min_eigen_val(g0)
    eref = golden["eig0"]
    # f64 running column sums inside OpenCV's box filter leave 1-ulp residues on a few pixels
    assert (e != eref).mean() < 2e-3
    assert np.abs(e - eref).max() <= 1e-7 * eref.max()
    pts = R.good_features_to_track(g0, 1300, 0.01, md, eig=e)
    assert np.array_equal(pts, golden["corners0"])


def test_lk_matches_golden(golden):
    pts = golden["corners0"][::3]
    out, st = R.calc_optical_flow_pyr_lk(golden["g0"], golden["g1"], pts)
    assert np.array_equal(st, golden["lk_status"][::3])
    ok = st == 1
    # north-star tolerance 0.05 px; the restatement replays OpenCV's float accumulation order: same bits
    assert np.array_equal(out[ok].view(np.uint32), golden["lk_pts"][::3][ok].view(np.uint32))


def test_lk_bit_identical_to_cv2(texture_small):
    """The oracle's LK (five-chain float accumulation, `_lk_chain_sum`) against cv2 itself on rendered frames with
    several pixels of motion: identical status and identical float32 bits, lost points included."""
    import cv2
    from conftest import render_clip
    fs = render_clip(texture_small, 320, 240, 3, start=20)
    gs = [cv2.cvtColor(f, cv2.COLOR_BGR2GRAY) for f in fs]
    pts = cv2.goodFeaturesToTrack(gs[0], 1300, 0.01, 5).reshape(-1, 2)[::4].copy()
    for a, b in ((0, 1), (0, 2)):
        ref, st, _ = cv2.calcOpticalFlowPyrLK(gs[a], gs[b], pts.reshape(-1, 1, 2), None, winSize=(21, 21), maxLevel=3,
                                              criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 50, 0.01),
                                              flags=0, minEigThreshold=1e-4)
        out, got = R.calc_optical_flow_pyr_lk(gs[a], gs[b], pts)
        assert np.array_equal(got, st.reshape(-1))
        assert np.array_equal(out.view(np.uint32), ref.reshape(-1, 2).view(np.uint32))


def test_lk_point_leaving_the_image_loses_status(golden):
    """A point tracked beyond the right/bottom border keeps its coordinates but gets status 0
    (the `err` block of OpenCV's LK at level 0; seen on the 720p clip at frame 38)."""
    import cv2
    g0 = golden["g0"]
    g1 = np.roll(g0, (3, 16), axis=(0, 1))               # content moves +16 px in x, +3 in y
    pts = np.array([[g0.shape[1] - 6, 100], [g0.shape[1] - 3, 40], [150, g0.shape[0] - 2], [160, 90]], np.float32)
    ref, st, _ = cv2.calcOpticalFlowPyrLK(g0, g1, pts.reshape(-1, 1, 2), None, winSize=(21, 21), maxLevel=3,
                                          criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 50, 0.01),
                                          flags=0, minEigThreshold=1e-4)
    out, got = R.calc_optical_flow_pyr_lk(g0, g1, pts)
    assert np.array_equal(got, st.reshape(-1))
    assert (st.reshape(-1) == 0).any() and (st.reshape(-1) == 1).any()
    assert np.array_equal(out, ref.reshape(-1, 2))       # lost points keep OpenCV's coordinates too


def test_lk_small_image_drops_pyramid_levels(golden):
    """On 128x96 OpenCV keeps pyramid levels 0..2 only (level 3 would be 16x12 <= the 21 px window)."""
    import cv2
    gray = lambda f: cv2.cvtColor(cv2.resize(f, (128, 96), interpolation=cv2.INTER_LINEAR), cv2.COLOR_BGR2GRAY)
    a, b = gray(golden["clip"][3]), gray(golden["clip"][4])
    pts = cv2.goodFeaturesToTrack(a, 1300, 0.01, 1).reshape(-1, 2)[:60]
    ref, st, _ = cv2.calcOpticalFlowPyrLK(a, b, pts.reshape(-1, 1, 2), None, winSize=(21, 21), maxLevel=3,
                                          criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 50, 0.01),
                                          flags=0, minEigThreshold=1e-4)
    out, got = R.calc_optical_flow_pyr_lk(a, b, pts)
    assert np.array_equal(got, st.reshape(-1))
    assert np.array_equal(out, ref.reshape(-1, 2))


def test_ransac_exact_consensus(golden):
    p, q = golden["ransac_p"], golden["ransac_q"]
    corners = np.array([[0, 0, 1], [320, 0, 1], [0, 180, 1], [320, 180, 1]], float).T
    for thr in (3, 5):
        M, mask = R.estimate_affine_partial_2d(p, q, float(thr))
        assert np.array_equal(mask, golden[f"ransac_inl_thr{thr}"])
        assert np.abs(M @ corners - golden[f"ransac_M_thr{thr}"] @ corners).max() < 1e-9


def test_ransac_random_vs_cv2():
    rng = np.random.default_rng(6)
    for trial in range(25):
        n = int(rng.integers(12, 900))
        p = np.stack([rng.uniform(0, 640, n), rng.uniform(0, 360, n)], 1).astype(np.float32)
        th = rng.uniform(-0.03, 0.03)
        A = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        q = (p @ A.T + rng.uniform(-8, 8, 2) + rng.normal(0, 0.3, (n, 2))).astype(np.float32)
        bad = rng.random(n) < rng.choice([0, 0.05, 0.3, 0.7])
        q[bad] += rng.uniform(-40, 40, (bad.sum(), 2)).astype(np.float32)
        Mr, inl = cv2.estimateAffinePartial2D(p.reshape(-1, 1, 2), q.reshape(-1, 1, 2), method=cv2.RANSAC)
        M, mask = R.estimate_affine_partial_2d(p, q, 3.0)
        assert np.array_equal(mask, inl.reshape(-1))
        assert np.abs(M - Mr).max() < 1e-8


# ---------------------------------------------------------------- copyFeathered (stabilizer.cpp:1051-1155)
def test_gaussian_blur_u8_fixed_point_tables():
    """cv::GaussianBlur on u8 is 8.8 / 16.16 fixed point with these tap tables (7x7 and 101x101, sigma 0)."""
    import cv2
    rng = np.random.default_rng(0)
    for shape in ((90, 120), (270, 480), (37, 53)):
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        assert np.array_equal(R.gaussian_blur_u8(img, R.GAUSS7_Q8), cv2.GaussianBlur(img, (7, 7), 0))
        if min(shape) > 50:
            assert np.array_equal(R.gaussian_blur_u8(img, R.GAUSS101_Q8), cv2.GaussianBlur(img, (101, 101), 0, 0))
    assert R.GAUSS7_Q8.sum() == 256 and R.GAUSS101_Q8.sum() == 256


def test_fill_convex_poly_rows_equal_opencv():
    """clipLine + left-to-right Bresenham outline + 16.16 scan-line edges == cv2.fillConvexPoly on warped inset rectangles,
    including vertices outside the image."""
    import cv2
    rng = np.random.default_rng(1)
    for _ in range(120):
        w, h = int(rng.integers(60, 400)), int(rng.integers(40, 300))
        ang, s = rng.uniform(-0.25, 0.25), rng.uniform(0.85, 1.15)
        Hm = np.array([[s * np.cos(ang), -s * np.sin(ang), rng.uniform(-40, 40)], [s * np.sin(ang), s * np.cos(ang), rng.uniform(-40, 40)],
                       [0, 0, 1.0]])
        c = np.float32([[10, 10], [w - 10, 10], [w - 10, h - 10], [10, h - 10]])
        poly = R.perspective_points(c, Hm)
        t = cv2.perspectiveTransform(c.reshape(-1, 1, 2), Hm).reshape(-1, 2)
        assert np.array_equal(poly, np.array([[int(np.rint(x)), int(np.rint(y))] for x, y in t]))
        ref = np.zeros((h, w), np.uint8)
        cv2.fillConvexPoly(ref, poly.astype(np.int32), 255)
        assert np.array_equal(R.mask_from_row_spans(R.fill_convex_poly_rows(w, h, poly), w), ref)


def test_copy_feathered_restatement_equals_opencv():
    """The integer restatement of Stabilizer::copyFeathered == the reference's cv2 call sequence, byte for byte."""
    import cv2
    from oracle import stabilizer_ref as sr
    rng = np.random.default_rng(4)
    for _ in range(4):
        w, h = int(rng.integers(120, 360)), int(rng.integers(110, 260))
        fg = cv2.GaussianBlur(rng.integers(0, 256, (h, w, 3), dtype=np.uint8), (0, 0), 1.2)
        bg = cv2.GaussianBlur(rng.integers(0, 256, (h, w, 3), dtype=np.uint8), (0, 0), 2.0)
        ang = rng.uniform(-0.08, 0.08)
        Hm = np.array([[np.cos(ang), -np.sin(ang), rng.uniform(-25, 25)], [np.sin(ang), np.cos(ang), rng.uniform(-25, 25)], [0, 0, 1.0]])
        assert np.array_equal(R.copy_feathered(fg, bg, Hm), sr.copy_feathered(fg, bg, Hm))
    with pytest.raises(ValueError):
        sr.copy_feathered(fg, bg[:-1], Hm)


def test_oracle_partial_lock_fix_feeds_the_reference_formulas():
    """StabilizerRef(partial_lock_fix=True): :1246-1260 fed with the accumulated lock; off (the reference) they see the identity."""
    from oracle import stabilizer_ref as sr, synth as osynth, camera_engine_ref as ce
    tex = osynth.make_texture(512)
    path = osynth.camera_path(16)
    frames = [ce.render_frame(tex, path[i], 320, 180, osynth.focal_for_width(320)) for i in range(16)]
    out = {}
    for fix in (False, True):
        for mode in (sr.TRANSLATION_LOCK, sr.ROTATION_LOCK, sr.ACCUMULATED_FULL_LOCK):
            ref = sr.StabilizerRef(4, 3, 180, faithful_waste=False, partial_lock_fix=fix)
            for i, f in enumerate(frames):
                if i == 6:
                    ref.set_stabilization_mode(mode)
                ref.stabilize_frame(f)
            out[(fix, mode)] = ref.taps.H_stabilize.copy()
    eye = np.eye(3)
    assert np.array_equal(out[(False, sr.TRANSLATION_LOCK)], eye) and np.array_equal(out[(False, sr.ROTATION_LOCK)], eye)
    full = out[(True, sr.ACCUMULATED_FULL_LOCK)]
    assert np.array_equal(full, out[(False, sr.ACCUMULATED_FULL_LOCK)])
    tl, rl = out[(True, sr.TRANSLATION_LOCK)], out[(True, sr.ROTATION_LOCK)]
    th = np.arctan2(full[1, 0], full[0, 0])
    assert abs(th) > 1e-5
    assert abs(np.arctan2(tl[1, 0], tl[0, 0])) < 1e-12
    assert abs(np.arctan2(rl[1, 0], rl[0, 0]) - th) < 1e-12
    c = np.array([160.0, 90.0, 1.0])
    assert np.allclose(rl @ c, c, atol=1e-9)
    # translation lock = R * H_lock with R a rotation about the centre: the centre is displaced as far as under the full lock
    assert abs(np.linalg.norm((tl @ c - c)[:2]) - np.linalg.norm((full @ c - c)[:2])) < 1e-9
