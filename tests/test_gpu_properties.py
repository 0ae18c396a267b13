"""T5 on the device: hypothesis draws the inputs, the CUDA kernels (through the C ABI) must agree with cv2
on every draw -- RANSAC consensus for any point cloud, bit-exact warp for any near-identity homography and any frame
size, bit-exact conditioning chain (median / sharpen / CLAHE) for any image, bit-identical LK for any shift, checksum of
the warp output == checksum of the bytes."""
import cv2
import numpy as np
import pytest

hyp = pytest.importorskip("hypothesis")
from hypothesis import given, settings, strategies as st  # noqa: E402

import vstab_b200 as vs  # noqa: E402
from oracle import stabilizer_ref as sr  # noqa: E402

pytestmark = pytest.mark.gpu
SET = dict(max_examples=20, deadline=None, derandomize=True)


@settings(**SET)
@given(seed=st.integers(0, 2**31 - 1), n=st.integers(10, 1300), outliers=st.floats(0.0, 0.7), noise=st.floats(0.0, 1.0),
       ang=st.floats(-0.05, 0.05), thr=st.sampled_from([3.0, 5.0]))
def test_fit_kernel_equals_opencv_for_any_cloud(seed, n, outliers, noise, ang, thr):
    rng = np.random.default_rng(seed)
    p = np.stack([rng.uniform(0, 640, n), rng.uniform(0, 360, n)], 1).astype(np.float32)
    A = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]])
    q = (p @ A.T + rng.uniform(-8, 8, 2) + rng.normal(0, noise, (n, 2))).astype(np.float32)
    bad = rng.random(n) < outliers
    q[bad] += rng.uniform(-40, 40, (int(bad.sum()), 2)).astype(np.float32)
    Mr, inl = cv2.estimateAffinePartial2D(p.reshape(-1, 1, 2), q.reshape(-1, 1, 2), method=cv2.RANSAC, ransacReprojThreshold=thr)
    M, T, cnt = vs.k_fit(p, q, np.ones(n, np.uint8), 640, 360, thresh=thr)
    if Mr is None:
        assert np.array_equal(T, np.eye(3))
        return
    assert cnt == (n, int(inl.sum()))
    corners = np.array([[0, 0, 1], [640, 0, 1], [0, 360, 1], [640, 360, 1]], float).T
    assert np.abs(M @ corners - Mr @ corners).max() < 1e-6


@settings(**SET)
@given(seed=st.integers(0, 2**31 - 1), w=st.integers(16, 700), h=st.integers(16, 400), ang=st.floats(-0.12, 0.12),
       tx=st.floats(-40, 40), ty=st.floats(-40, 40), scale=st.floats(0.9, 1.1), p0=st.floats(-1e-5, 1e-5))
def test_warp_kernel_bit_exact_for_any_homography_and_size(seed, w, h, ang, tx, ty, scale, p0):
    rng = np.random.default_rng(seed)
    src = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    c, s = np.cos(ang) * scale, np.sin(ang) * scale
    Hm = np.array([[c, -s, tx], [s, c, ty], [p0, 0.0, 1.0]])
    bd = [int(v) for v in rng.integers(0, 256, 3)]
    ref = cv2.warpPerspective(src, Hm, (w, h), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT,
                              borderValue=tuple(float(v) for v in bd))
    assert np.array_equal(vs.k_warp(src, Hm, bd), ref)


@settings(**SET)
@given(seed=st.integers(0, 2**31 - 1), w=st.integers(64, 500), h=st.integers(100, 300), kind=st.sampled_from(["noise", "smooth", "flat"]))
def test_conditioning_chain_bit_exact_for_any_image(seed, w, h, kind):
    """median5 -> sharpen -> CLAHE(2.0, 8x8) -> median5 on an NN-resized gray (src/stabilizer.cpp:448-477)."""
    rng = np.random.default_rng(seed)
    if kind == "noise":
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    elif kind == "flat":
        img = np.full((h, w, 3), int(rng.integers(0, 256)), np.uint8)
    else:
        img = cv2.GaussianBlur(rng.integers(0, 256, (h, w, 3), dtype=np.uint8), (0, 0), 3.0)
    wh = max(91, h // 2)
    o = sr.StabilizerRef(15, 15, wh)
    o._initialize_frame(img)
    assert np.array_equal(vs.k_featprep(img, wh), o._preprocess_for_features(img))


@settings(max_examples=10, deadline=None, derandomize=True)
@given(seed=st.integers(0, 2**31 - 1), dx=st.floats(-6, 6), dy=st.floats(-6, 6), contrast=st.floats(0.2, 1.0))
def test_lk_kernel_bit_identical_for_any_shift(seed, dx, dy, contrast):
    """Sub-pixel shifted random blobs: every tier of the tracker's exactness ladder shows up (low contrast: exact sums;
    full contrast and several pixels of motion: the chain replay)."""
    rng = np.random.default_rng(seed)
    base = cv2.GaussianBlur(rng.integers(0, 256, (240, 320), dtype=np.uint8), (0, 0), 1.5)
    base = cv2.normalize(base, None, 128 - 127 * contrast, 128 + 127 * contrast, cv2.NORM_MINMAX).astype(np.uint8)
    M = np.float32([[1, 0, dx], [0, 1, dy]])
    nxt = cv2.warpAffine(base, M, (320, 240), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)
    pts = cv2.goodFeaturesToTrack(base, 400, 0.01, 5)
    if pts is None:
        return
    pts = pts.reshape(-1, 2)
    ref, stt, _ = cv2.calcOpticalFlowPyrLK(base, nxt, pts.reshape(-1, 1, 2), None, winSize=(21, 21), maxLevel=3,
                                           criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 50, 0.01), flags=0,
                                           minEigThreshold=1e-4)
    out, got = vs.k_lk(base, nxt, pts)
    assert np.array_equal(got, stt.ravel())
    ok = got == 1
    assert np.array_equal(out[ok].view(np.uint32), ref.reshape(-1, 2)[ok].view(np.uint32))
