"""T5: property tests.  hypothesis draws the inputs; the properties are the ones the domain offers --
the RANSAC restatement returns OpenCV's consensus set for any point cloud, the warp restatement equals cv2 for any
near-identity homography and is the identity for H = I, the decompose/compose pair round-trips, the LK chain sum is
exact below 2^24, the frame checksum is linear and order-free.  CPU only (the oracle is the thing under test here);
the GPU twins of the kernel-level properties live in tests/test_gpu_properties.py."""
import cv2
import numpy as np
import pytest

hyp = pytest.importorskip("hypothesis")
from hypothesis import given, settings, strategies as st  # noqa: E402

from oracle import cv_restate as R  # noqa: E402
from oracle import stabilizer_ref as sr  # noqa: E402
import vstab_b200 as vs  # noqa: E402

SET = dict(max_examples=25, deadline=None, derandomize=True)


@settings(**SET)
@given(seed=st.integers(0, 2**31 - 1), n=st.integers(10, 400), outliers=st.floats(0.0, 0.6), noise=st.floats(0.0, 1.5),
       ang=st.floats(-0.2, 0.2), scale=st.floats(0.8, 1.25), thr=st.sampled_from([3.0, 5.0]))
def test_ransac_restatement_equals_opencv_for_any_cloud(seed, n, outliers, noise, ang, scale, thr):
    rng = np.random.default_rng(seed)
    p = rng.uniform(0, 640, (n, 2)).astype(np.float32)
    c, s = np.cos(ang) * scale, np.sin(ang) * scale
    q = (p @ np.array([[c, s], [-s, c]], np.float32)) + rng.uniform(-30, 30, 2).astype(np.float32)
    q = (q + rng.normal(0, noise, q.shape)).astype(np.float32)
    bad = rng.random(n) < outliers
    q[bad] = rng.uniform(0, 640, (int(bad.sum()), 2)).astype(np.float32)
    Mref, inl = cv2.estimateAffinePartial2D(p, q, method=cv2.RANSAC, ransacReprojThreshold=thr)
    M, mask = R.estimate_affine_partial_2d(p, q, thr)
    if Mref is None:
        assert M is None or not np.isfinite(M).all()
        return
    assert np.array_equal(mask.astype(bool).ravel(), inl.astype(bool).ravel())       # identical consensus set
    corners = np.array([[0, 0, 1], [640, 0, 1], [0, 360, 1], [640, 360, 1]], float).T
    assert np.abs(M @ corners - Mref @ corners).max() < 1e-6                        # closed-form LS == OpenCV's LM refinement


@settings(**SET)
@given(seed=st.integers(0, 2**31 - 1), w=st.integers(17, 90), h=st.integers(17, 70), ang=st.floats(-0.15, 0.15),
       tx=st.floats(-12, 12), ty=st.floats(-12, 12), p0=st.floats(-2e-4, 2e-4), p1=st.floats(-2e-4, 2e-4))
def test_warp_restatement_equals_opencv_for_any_homography(seed, w, h, ang, tx, ty, p0, p1):
    rng = np.random.default_rng(seed)
    src = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    c, s = np.cos(ang), np.sin(ang)
    Hm = np.array([[c, -s, tx], [s, c, ty], [p0, p1, 1.0]])
    bd = tuple(float(v) for v in rng.integers(0, 256, 3))
    ref = cv2.warpPerspective(src, Hm, (w, h), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=bd)
    assert np.array_equal(R.warp_perspective_bgr(src, Hm, bd), ref)


@settings(**SET)
@given(seed=st.integers(0, 2**31 - 1), w=st.integers(12, 64), h=st.integers(12, 64))
def test_warp_identity_is_identity(seed, w, h):
    src = np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)
    assert np.array_equal(R.warp_perspective_bgr(src, np.eye(3), (1.0, 2.0, 3.0)), src)


@settings(**SET)
@given(s=st.floats(0.5, 2.0), theta=st.floats(-3.0, 3.0), k=st.floats(0.5, 2.0), delta=st.floats(-0.5, 0.5),
       tx=st.floats(-100, 100), ty=st.floats(-100, 100), v0=st.floats(-1e-4, 1e-4), v1=st.floats(-1e-4, 1e-4),
       cx=st.floats(0, 640), cy=st.floats(0, 360))
def test_decompose_compose_round_trip(s, theta, k, delta, tx, ty, v0, v1, cx, cy):
    """composeHomography(decomposeHomography(H)) == H (src/stabilizer.cpp:1435-1566) for any valid parameter set."""
    p = sr.HomographyParameters(s, theta, k, delta, (tx, ty), (v0, v1))
    Hm = sr.compose_homography(p, (cx, cy))
    q = sr.decompose_homography(Hm, (cx, cy))
    assert q is not None
    H2 = sr.compose_homography(q, (cx, cy))
    assert np.abs(H2 / H2[2, 2] - Hm / Hm[2, 2]).max() < 1e-8 * max(1.0, np.abs(Hm).max())


@settings(**SET)
@given(seed=st.integers(0, 2**31 - 1), mag=st.integers(1, 38000))
def test_lk_chain_sum_is_the_exact_sum_below_2_pow_24(seed, mag):
    """OpenCV's five-chain float accumulation is order-free (= the integer sum) whenever the absolute values of
    the terms sum to less than 2^24 -- the rule csrc/lk.cu's exactness tiers rest on."""
    rng = np.random.default_rng(seed)
    t = rng.integers(-mag, mag + 1, (3, 21, 21)).astype(np.int64)
    want = t.sum(axis=(1, 2))
    assert np.abs(t).sum(axis=(1, 2)).max() < 2**24
    got = R._lk_chain_sum(t.astype(np.float32))
    assert np.array_equal(got.astype(np.int64), want)


@settings(**SET)
@given(seed=st.integers(0, 2**31 - 1), w=st.integers(1, 40), h=st.integers(1, 30))
def test_frame_checksum_is_linear_in_the_pixels(seed, w, h):
    """checksum(a + b) == checksum(a) + checksum(b) mod 2^64 when a + b does not overflow a byte: every byte enters with a
    fixed positional weight, so a sum of per-shard checksums is shard-split independent."""
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 128, (h, w, 3), dtype=np.uint8)
    b = rng.integers(0, 128, (h, w, 3), dtype=np.uint8)
    m = (1 << 64) - 1
    assert vs.frame_checksum(a + b) == (vs.frame_checksum(a) + vs.frame_checksum(b)) & m
    assert vs.frame_checksum(np.zeros_like(a)) == 0
