"""Oracle pipeline (oracle/stabilizer_ref.py) against the golden clip, plus the API semantics
of the reference it restates (T3 in SURVEY 7.4) and the streaming/offline index algebra of
Appendix C.  CPU only."""
import numpy as np
import pytest

from oracle import stabilizer_ref as sr


def test_ctor_argument_checks():
    for args in ((0, 0, 360), (5, 5, 90), (5, 5, 2161)):
        with pytest.raises(ValueError):
            sr.StabilizerRef(*args)
    s = sr.StabilizerRef(60, 45, 360)
    assert s.total_frame_window_size() == 106
    assert s.mode == sr.GLOBAL_SMOOTHING
    assert (sr.ACCUMULATED_FULL_LOCK, sr.ORB_FULL_LOCK, sr.SIFT_FULL_LOCK, sr.TRANSLATION_LOCK,
            sr.ROTATION_LOCK, sr.GLOBAL_SMOOTHING) == (0, 1, 2, 3, 4, 5)


def test_golden_clip_reproduces(golden):
    clip = golden["clip"]
    for name, lock_at in (("smooth", None), ("lock", 7)):
        s = sr.StabilizerRef(4, 3, 96)
        for i, fr in enumerate(clip):
            if lock_at is not None and i == lock_at:
                s.set_stabilization_mode(sr.ACCUMULATED_FULL_LOCK)
            out = s.stabilize_frame(fr)
            assert np.array_equal(out, golden[f"clip_{name}_out"][i])
            if i:
                assert np.allclose(s.taps.T, golden[f"clip_{name}_T"][i], rtol=0, atol=1e-12)


def test_first_frame_passthrough_and_warmup(golden):
    clip = golden["clip"]
    s = sr.StabilizerRef(4, 3, 96)
    out0 = s.stabilize_frame(clip[0])
    assert out0 is clip[0] or np.array_equal(out0, clip[0])
    pres = []
    for fr in clip[1:]:
        s.stabilize_frame(fr)
        pres.append(s.taps.presentation_idx)
    # frame 0 is presented F+1 times in total (call 0 + calls 1..F), then p = n - F
    assert pres == [max(0, n - 3) for n in range(1, len(clip))]


def test_size_change_raises(golden):
    s = sr.StabilizerRef(4, 3, 96)
    s.stabilize_frame(golden["clip"][0])
    with pytest.raises(ValueError):
        s.stabilize_frame(golden["clip"][1][:100])
    with pytest.raises(ValueError):
        sr.StabilizerRef(4, 3, 96).stabilize_frame(np.zeros((10, 200, 3), np.uint8))


def test_lock_before_window_advances_asserts(golden):
    s = sr.StabilizerRef(4, 3, 96)
    s.stabilize_frame(golden["clip"][0])
    s.set_stabilization_mode(sr.ACCUMULATED_FULL_LOCK)
    s.stabilize_frame(golden["clip"][1])
    with pytest.raises(AssertionError):
        s.stabilize_frame(golden["clip"][2])      # presentation index still 0


def test_translation_rotation_lock_are_identity(golden):
    for mode in (sr.TRANSLATION_LOCK, sr.ROTATION_LOCK):
        s = sr.StabilizerRef(4, 3, 96)
        s.set_stabilization_mode(mode)
        for fr in golden["clip"][:6]:
            s.stabilize_frame(fr)
        assert np.allclose(s.taps.H_stabilize, np.eye(3), atol=1e-12)


def test_decompose_compose_roundtrip(golden):
    H = golden["decomp_H"]
    p = sr.decompose_homography(H, (320.0, 180.0))
    got = np.array([p.s, p.theta, p.k, p.delta, p.t[0], p.t[1], p.v[0], p.v[1]])
    assert np.allclose(got, golden["decomp_params"], rtol=0, atol=1e-14)
    assert np.abs(sr.compose_homography(p, (320.0, 180.0)) - H).max() < 1e-12
    assert sr.decompose_homography(np.zeros((3, 3))) is None                 # h33 ~ 0
    bad = np.eye(3)
    bad[0, 0] = -1.0
    assert sr.decompose_homography(bad) is None                              # det < 0
    with pytest.raises(ValueError):
        sr.decompose_homography(np.eye(3, dtype=np.float32))


def _window_average(T, c, P, F):
    """SURVEY Appendix C closed form for call c given transforms T[1..c]."""
    W = P + 1 + F
    lo = max(0, c - W + 1)
    p = max(0, c - F)
    acc, total, count = np.eye(3), np.zeros((3, 3)), 0
    for k in range(p, lo, -1):
        acc = np.linalg.inv(T[k]) @ acc
        total += acc
        count += 1
    acc = np.eye(3)
    for k in range(p + 1, c):
        acc = acc @ T[k]
        total += acc
        count += 1
    return total / count if count else np.eye(3)


@pytest.mark.parametrize("P,F", [(4, 3), (5, 0), (0, 7), (3, 3)])
def test_index_algebra_matches_deque_transcription(P, F):
    """The absolute-index algebra used by K6 equals the reference's deque bookkeeping."""
    rng = np.random.default_rng(P * 10 + F)
    n = 30
    T = [np.eye(3)]
    for _ in range(1, n):
        th = rng.normal(0, 0.01)
        T.append(np.array([[np.cos(th), -np.sin(th), rng.normal(0, 3)], [np.sin(th), np.cos(th), rng.normal(0, 3)], [0, 0, 1]]))
    s = sr.StabilizerRef(P, F, 360)
    from collections import deque
    frames, transforms = deque(), deque()
    for c in range(n):
        frames.append((None, c))
        while len(frames) > s.total_frame_window_size():
            frames.popleft()
        if c == 0:
            continue
        transforms.append((T[c], c - 1, c))
        while len(transforms) > s.total_frame_window_size() - 1:
            transforms.popleft()
        s.frames, s.transforms = frames, transforms
        p_local = len(frames) - F - 1 if len(frames) > F else 0
        assert frames[p_local][1] == max(0, c - F)
        ref = s._global_smoothing(p_local)
        assert np.abs(ref - _window_average(T, c, P, F)).max() < 1e-12
