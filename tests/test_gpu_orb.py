"""K8-K10: the ORB registration path kernels against cv2 4.13.0 (bit-exact integer work):
preprocessing chain, ORB keypoint set / angles / descriptors over all pyramid levels, Hamming
2-NN match indices and the ratio test."""
import cv2
import numpy as np
import pytest

import vstab_b200 as vs
from conftest import render_clip
from oracle import stabilizer_ref as sr

pytestmark = pytest.mark.gpu


def _cv_prep(frame, wh):
    o = sr.StabilizerRef(15, 15, wh)
    o._initialize_frame(frame)
    return o._preprocess_for_features(frame)


@pytest.mark.parametrize("W,H,wh", [(1280, 720, 360), (1920, 1080, 1080), (1920, 1080, 360), (640, 360, 100), (333, 250, 97)])
def test_featprep_bit_exact(texture, W, H, wh):
    frame = render_clip(texture, W, H, 1, start=4)[0]
    got = vs.k_featprep(frame, wh)
    want = _cv_prep(frame, wh)
    assert got.shape == want.shape
    assert np.array_equal(got, want)


def test_featprep_noise_and_flat():
    rng = np.random.default_rng(5)
    for img in (rng.integers(0, 256, (360, 640, 3), dtype=np.uint8), np.full((360, 640, 3), 77, np.uint8),
                np.zeros((360, 640, 3), np.uint8)):
        assert np.array_equal(vs.k_featprep(img, 180), _cv_prep(img, 180))


def _cv_orb(gray):
    det = cv2.ORB_create(2500, 1.2, 12, 31, 0, 2, cv2.ORB_FAST_SCORE, 31, 20)
    kps, desc = det.detectAndCompute(gray, None)
    return kps, desc


def _as_dict(kps, desc):
    """keypoints keyed by (octave, rounded level coordinates): order independent comparison"""
    return {(int(k[5]), round(float(k[0]), 3), round(float(k[1]), 3)): (float(k[2]), float(k[3]), float(k[4]), bytes(d))
            for k, d in zip(kps, desc)}


@pytest.mark.parametrize("W,H,wh", [(1280, 720, 360), (1920, 1080, 1080)])
def test_orb_identical_keypoints_and_descriptors(texture, W, H, wh):
    frame = render_clip(texture, W, H, 1, start=7)[0]
    gray = _cv_prep(frame, wh)
    kps, desc = vs.k_orb(gray)
    ckps, cdesc = _cv_orb(gray)
    want = {(k.octave, round(k.pt[0], 3), round(k.pt[1], 3)): (k.size, k.angle, k.response, bytes(d)) for k, d in zip(ckps, cdesc)}
    got = _as_dict(kps, desc)
    assert len(got) == len(kps)
    assert set(got) == set(want)                               # identical FAST corner sets on every level
    for key, (size, angle, resp, d) in want.items():
        g = got[key]
        assert g[0] == np.float32(size) and g[2] == resp
        assert g[1] == np.float32(angle), (key, g[1], angle)   # fastAtan2 polynomial, bit-identical
        assert g[3] == d, key                                  # 256-bit rBRIEF descriptor
    # level-major, row-major order
    order = [(int(k[5]), float(k[1]), float(k[0])) for k in kps]
    assert order == sorted(order)


@pytest.mark.parametrize("W,H,wh", [(1280, 720, 360), (1920, 1080, 1080), (1920, 1080, 720)])
def test_orb_reference_order_equals_opencv(texture, W, H, wh):
    """reference_order replays retainBest's std::nth_element + std::partition permutation: keypoints and
    descriptors come out in exactly cv::ORB's order (which fixes the RANSAC sample sequence downstream)."""
    frame = render_clip(texture, W, H, 1, start=11)[0]
    gray = _cv_prep(frame, wh)
    kps, desc = vs.k_orb(gray, reference_order=True)
    ckps, cdesc = _cv_orb(gray)
    assert len(kps) == len(ckps)
    want = np.array([[k.pt[0], k.pt[1], k.octave] for k in ckps], np.float32)
    assert np.array_equal(kps[:, [0, 1, 5]], want)
    assert np.array_equal(desc, cdesc)


def test_orb_size_filter_and_empty(texture):
    frame = render_clip(texture, 1280, 720, 1, start=7)[0]
    gray = _cv_prep(frame, 360)
    kps, desc = vs.k_orb(gray, size_ratio=0.10)
    ckps, cdesc = _cv_orb(gray)
    ckps, cdesc = sr.filter_keypoints_by_relative_size(gray.shape[0], list(ckps), cdesc, 0.10)
    assert len(kps) == len(ckps) and len(kps) > 100
    assert set(map(bytes, desc)) == set(map(bytes, cdesc))
    flat = np.full((360, 640), 90, np.uint8)
    kps, desc = vs.k_orb(flat)
    assert len(kps) == 0


def test_hamming_knn_and_ratio(texture):
    f0, f1 = render_clip(texture, 1280, 720, 2, start=9)
    d0 = _cv_orb(_cv_prep(f0, 360))[1]
    d1 = _cv_orb(_cv_prep(f1, 360))[1]
    bi, bd, sd, good = vs.k_hamming(d0, d1, 0.6)
    knn = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(d0, d1, 2)
    assert len(knn) == len(d0)
    for q, pair in enumerate(knn):
        assert pair[0].trainIdx == bi[q] and int(pair[0].distance) == bd[q] and int(pair[1].distance) == sd[q]
        assert bool(good[q]) == bool(pair[0].distance < np.float32(0.6) * np.float32(pair[1].distance))
    assert good.sum() > 50
    # degenerate train sets
    bi, bd, sd, good = vs.k_hamming(d0[:5], d1[:1], 0.6)
    assert (good == 0).all() and (bi == 0).all()


def test_l2_match_tcgen05_equals_bfmatcher(texture):
    """K12: exact L2 nearest neighbour of real SIFT descriptors on the tensor cores == cv2.BFMatcher(NORM_L2)
    (indices and distances), plus the reference's distance filter d <= max(0.5 * mean(d), 0.02)."""
    f0, f1 = render_clip(texture, 1280, 720, 2, start=9)
    sift = cv2.SIFT_create(2500, 3, 0.04, 5, 1.2)
    d0 = sift.detectAndCompute(_cv_prep(f0, 720), None)[1]
    d1 = sift.detectAndCompute(_cv_prep(f1, 720), None)[1]
    assert len(d0) > 2000 and len(d1) > 2000
    bi, bd2, good = vs.k_l2match(d0, d1)
    ms = cv2.BFMatcher(cv2.NORM_L2).match(d0, d1)
    assert len(ms) == len(d0)
    want_idx = np.array([m.trainIdx for m in ms])
    want_d = np.array([m.distance for m in ms], np.float32)
    # exact squared distances from the integer descriptors (numpy int64)
    d2 = ((d0.astype(np.int64)[:, None, :] - d1.astype(np.int64)[None, :, :]) ** 2).sum(-1)
    assert np.array_equal(bd2, d2.min(1))
    assert np.array_equal(bi, d2.argmin(1))                      # ties -> lowest index
    assert np.array_equal(bi, want_idx)
    assert np.array_equal(np.sqrt(bd2.astype(np.float32)), want_d)
    avg = float(np.sum(want_d.astype(np.float64))) / len(d0)
    assert np.array_equal(good.astype(bool), want_d.astype(np.float64) <= max(avg * 0.5, 0.02))
    # a batch of current sets against the one reference set in ONE launch (frames along the grid's z axis)
    sets = [d1, d1[:1000], d0[::-1].copy(), d1[300:301]]
    bis, bds, _ = vs.k_l2match_batch(d0, sets)
    for k, c in enumerate(sets):
        dd = ((d0.astype(np.int64)[:, None, :] - c.astype(np.int64)[None, :, :]) ** 2).sum(-1)
        assert np.array_equal(bds[k], dd.min(1)) and np.array_equal(bis[k], dd.argmin(1))
    # ragged sizes: a non-multiple of the 128-row tiles on both sides, and a single train row
    for nr, nc in ((300, 77), (129, 1), (5, 2500)):
        bi, bd2, _ = vs.k_l2match(d0[:nr], d1[:nc])
        assert np.array_equal(bi, d2[:nr, :nc].argmin(1)) and np.array_equal(bd2, d2[:nr, :nc].min(1))


def _sift_overlap(gray):
    kps, desc = vs.k_sift(gray)
    ckps, cdesc = cv2.SIFT_create(2500, 3, 0.04, 5, 1.2).detectAndCompute(gray, None)
    mine = {}
    for i, k in enumerate(kps):
        mine.setdefault((int(k[5]) & 0xffff, round(float(k[0]), 1), round(float(k[1]), 1)), []).append(i)
    hit, pos_err, ang_err, dsc_err = 0, [], [], []
    for k, d in zip(ckps, cdesc):
        best = None
        for dx in (-0.1, 0.0, 0.1):
            for dy in (-0.1, 0.0, 0.1):
                for i in mine.get((k.octave & 0xffff, round(k.pt[0] + dx, 1), round(k.pt[1] + dy, 1)), []):
                    e = max(abs(kps[i][0] - k.pt[0]), abs(kps[i][1] - k.pt[1]))
                    a = abs((kps[i][3] - k.angle + 180.0) % 360.0 - 180.0)
                    if e <= 0.05 and a <= 2.0 and (best is None or e + a < best[0]):
                        best = (e + a, e, a, i)
        if best is not None:
            hit += 1
            pos_err.append(best[1]); ang_err.append(best[2])
            dsc_err.append(np.abs(desc[best[3]].astype(float) - d).mean())
    return len(kps), len(ckps), hit, np.array(pos_err), np.array(ang_err), np.array(dsc_err)


@pytest.mark.parametrize("W,H,wh", [(1280, 720, 360), (1920, 1080, 1080), (3840, 2160, 2160)])
def test_sift_overlaps_opencv(texture, W, H, wh):
    """K11 is a float pipeline and not bit-pinned: the keypoint set must overlap cv2's
    (same octave/layer, position <= 0.05 px, orientation <= 2 deg) and matched descriptors must agree."""
    frame = render_clip(texture, W, H, 1, start=7)[0]
    gray = _cv_prep(frame, wh)
    n, nc, hit, pe, ae, de = _sift_overlap(gray)
    print(f"SIFT {W}x{H} wh{wh}: ours {n}, cv2 {nc}, matched {hit}, pos p99 {np.percentile(pe, 99):.4f}px, "
          f"angle p99 {np.percentile(ae, 99):.3f}deg, mean |desc diff| {de.mean():.3f}")
    assert abs(n - nc) <= 0.03 * nc
    assert hit >= 0.95 * nc
    assert np.percentile(pe, 99) <= 0.02 and np.percentile(ae, 99) <= 0.5
    assert de.mean() <= 1.0                     # descriptor entries are 0..255
