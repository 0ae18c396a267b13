"""TEST INFRASTRUCTURE (oracle) -- CPU restatement of the reference hot path.

This is a line-faithful Python restatement of `Stabilizer::stabilizeFrame`
(/root/reference/src/stabilizer.cpp:1158-1325) and everything it calls.  The
reference itself cannot be compiled in this image (no OpenCV C++ SDK; SURVEY.md
§8c), and all of its heavy arithmetic lives in the un-vendored, unpinned
third-party dependency **OpenCV 4.x** (/root/reference/CMakeLists.txt:7).  The
stand-in is the same OpenCV core through `opencv-python-headless 4.13.0.92`
(`cv2`), called at exactly the reference's call sites with exactly the
reference's constants.  The reference holds no tests or golden vectors for this
path, so **parity is pinned against cv2 4.13.0 outputs, not
against fixtures of the reference: "parity unpinned" by the reference itself.**

Only tests/, __graft_entry__.smoke() and bench.py's CPU arms may import this
module; it is never on the product path.

Reference quirks reproduced on purpose:
  * goodFeaturesToTrack runs twice per frame (stabilizer.cpp:949,961);
  * the smoothing average has no identity term and excludes the newest
    transform (stabilizer.cpp:805-849);
  * the future chain is right-multiplied (stabilizer.cpp:834);
  * TRANSLATION_LOCK / ROTATION_LOCK evaluate to identity (stabilizer.cpp:313-441,
    1241-1275);
  * only the translation is rescaled to full resolution (stabilizer.cpp:1291-1296);
  * full-frame clones at :129, :160, :1319 (kept when `faithful_waste=True` so
    the CPU baseline is what the reference really costs).
"""
from __future__ import annotations

import math
import time
from collections import deque
from dataclasses import dataclass, field

import cv2
import numpy as np

# enum class StabilizationMode, include/stabilizer.hpp:31-38
ACCUMULATED_FULL_LOCK = 0
ORB_FULL_LOCK = 1
SIFT_FULL_LOCK = 2
TRANSLATION_LOCK = 3
ROTATION_LOCK = 4
GLOBAL_SMOOTHING = 5

MIN_POINTS_FOR_MOTION_ESTIMATION = 10  # stabilizer.cpp:20


@dataclass
class HomographyParameters:
    """include/stabilizer.hpp:44-57"""
    s: float = 1.0
    theta: float = 0.0
    k: float = 1.0
    delta: float = 0.0
    t: tuple = (0.0, 0.0)
    v: tuple = (0.0, 0.0)


def qr_decomposition_2x2(A: np.ndarray):
    """stabilizer.cpp:1342-1432 (Gram-Schmidt).  Raises like the reference."""
    eps = 1e-6
    det = A[0, 0] * A[1, 1] - A[0, 1] * A[1, 0]
    if abs(det) < eps:
        raise ValueError("singular")
    a1 = A[:, 0].copy()
    a2 = A[:, 1].copy()
    n1 = math.sqrt(a1[0] * a1[0] + a1[1] * a1[1])
    if n1 < eps:
        raise RuntimeError("first column zero")
    q1 = a1 / n1
    r12 = a2[0] * q1[0] + a2[1] * q1[1]
    u2 = a2 - r12 * q1
    n2 = math.sqrt(u2[0] * u2[0] + u2[1] * u2[1])
    if n2 < eps:
        raise RuntimeError("dependent columns")
    q2 = u2 / n2
    Q = np.array([[q1[0], q2[0]], [q1[1], q2[1]]])
    Rm = np.array([[n1, r12], [0.0, n2]])
    if np.abs(A - Q @ Rm).sum(axis=1).max() > eps:
        raise RuntimeError("QR failed")
    if np.abs(Q.T @ Q - np.eye(2)).sum(axis=1).max() > eps:
        raise RuntimeError("Q not orthogonal")
    return Q, Rm


def decompose_homography(H: np.ndarray, center=(0.0, 0.0)):
    """stabilizer.cpp:1435-1533.  Returns HomographyParameters or None (== false)."""
    H = np.asarray(H)
    if H.shape != (3, 3) or H.dtype != np.float64:
        raise ValueError("H must be 3x3 float64")
    eps = 1e-6
    if not np.all(np.isfinite(H)):
        return None
    h33 = H[2, 2]
    if abs(h33) < eps:
        return None
    Hn = H / h33
    t = Hn[0:2, 2].copy()
    v = Hn[2, 0:2].copy()
    A = Hn[0:2, 0:2]
    sRK = A - np.outer(t, v)
    if not np.all(np.isfinite(sRK)):
        return None
    det = sRK[0, 0] * sRK[1, 1] - sRK[0, 1] * sRK[1, 0]
    if math.isnan(det) or math.isinf(det) or det < 0 or abs(det) < eps:
        return None
    s = math.sqrt(det)
    RK = sRK / s
    try:
        R, K = qr_decomposition_2x2(RK)
    except (ValueError, RuntimeError):
        # the reference lets this exception escape (Appendix B.14); the
        # replacement returns false instead, and so does the oracle.
        return None
    if not (np.all(np.isfinite(R)) and np.all(np.isfinite(K))):
        return None
    detR = R[0, 0] * R[1, 1] - R[0, 1] * R[1, 0]
    if abs(detR - 1.0) > eps:
        return None
    cos_t = (R[0, 0] + R[1, 1]) / 2
    sin_t = (R[1, 0] - R[0, 1]) / 2
    theta = math.atan2(sin_t, cos_t)
    k1 = K[0, 0]
    delta = K[0, 1]
    c = np.array(center, dtype=np.float64)
    t_shift = (np.eye(2) - s * R) @ c
    t_shifted = t - t_shift
    return HomographyParameters(s, theta, k1, delta,
                                (float(t_shifted[0]), float(t_shifted[1])),
                                (float(v[0]), float(v[1])))


def compose_homography(p: HomographyParameters, center=(0.0, 0.0)) -> np.ndarray:
    """stabilizer.cpp:1535-1566."""
    R = np.array([[math.cos(p.theta), -math.sin(p.theta)],
                  [math.sin(p.theta), math.cos(p.theta)]])
    K = np.array([[p.k, p.delta], [0.0, 1 / p.k]])
    c = np.array(center, dtype=np.float64)
    t_shift = (np.eye(2) - p.s * R) @ c
    t_shifted = np.array(p.t) + t_shift
    A = p.s * R @ K + np.outer(t_shifted, np.array(p.v))
    H = np.eye(3)
    H[0:2, 0:2] = A
    H[0, 2] = t_shifted[0]
    H[1, 2] = t_shifted[1]
    H[2, 0] = p.v[0]
    H[2, 1] = p.v[1]
    return H


def copy_feathered(foreground: np.ndarray, background_image: np.ndarray, H: np.ndarray) -> np.ndarray:
    """Stabilizer::copyFeathered, stabilizer.cpp:1051-1155, call for call over cv2 (the reference keeps it behind
    `#if 0` at :1304; `StabilizerRef(trail=True)` takes that branch)."""
    if foreground.shape != background_image.shape:
        raise ValueError("Stabilizer: copyFeathered: foreground and background_image must have the same size")
    H = np.asarray(H)
    if H.shape != (3, 3) or H.dtype != np.float64 or not np.all(np.isfinite(H)):
        raise ValueError("Stabilizer: copyFeathered: Bad homography matrix H.")
    h, w = foreground.shape[:2]
    warped = cv2.warpPerspective(foreground, H, (w, h)).astype(np.float32)
    bg = cv2.cvtColor(background_image, cv2.COLOR_BGR2GRAY)
    bg = cv2.GaussianBlur(bg, (7, 7), 0)
    bg = cv2.multiply(bg, 0.99)                                   # background_image_changed *= 0.99
    bgf = cv2.cvtColor(bg, cv2.COLOR_GRAY2BGR).astype(np.float32)
    B = 10
    corners = np.float32([[B, B], [w - B, B], [w - B, h - B], [B, h - B]]).reshape(-1, 1, 2)
    tc = cv2.perspectiveTransform(corners, H).reshape(-1, 2)
    poly = np.array([[int(np.rint(x)), int(np.rint(y))] for x, y in tc], np.int32)     # cv::Point(Point2f): cvRound
    mask = np.zeros((h, w), np.uint8)
    cv2.fillConvexPoly(mask, poly, 255)
    mask = cv2.warpPerspective(mask, H, (w, h))
    alpha = cv2.GaussianBlur(mask, (101, 101), 0, 0).astype(np.float32) * np.float32(1.0 / 255.0)
    alpha3 = cv2.cvtColor(alpha, cv2.COLOR_GRAY2BGR)
    fgc = cv2.multiply(alpha3, warped)
    bgc = cv2.multiply(1.0 - alpha3, bgf)
    return np.clip(np.rint(cv2.add(fgc, bgc)), 0, 255).astype(np.uint8)


def filter_keypoints_by_relative_size(image_height, kps, desc, max_rel=0.05):
    """stabilizer.cpp:290-309 (float compare: size < float(image_height*ratio))."""
    max_allowed = np.float32(np.float32(image_height) * np.float32(max_rel))
    keep = [i for i, k in enumerate(kps) if np.float32(k.size) < max_allowed]
    kps2 = [kps[i] for i in keep]
    desc2 = desc[keep] if desc is not None and len(keep) else (
        None if desc is None else desc[:0])
    return kps2, desc2


@dataclass
class Taps:
    """Per-call intermediate results exposed for parity tests."""
    gray: np.ndarray | None = None
    prev_pts: np.ndarray | None = None       # corners detected on the previous gray
    lk_pts: np.ndarray | None = None         # raw LK output for prev_pts
    lk_status: np.ndarray | None = None
    M: np.ndarray | None = None              # 2x3 from estimateAffinePartial2D
    inliers: np.ndarray | None = None
    T: np.ndarray | None = None              # 3x3 pushed into the window
    H_smooth: np.ndarray | None = None
    H_lock: np.ndarray | None = None
    H_stabilize: np.ndarray | None = None
    H_scaled: np.ndarray | None = None
    border: tuple | None = None
    presentation_idx: int = 0                # absolute frame index presented
    new_pts: np.ndarray | None = None
    n_matches: int = 0


class StabilizerRef:
    """Mirror of class Stabilizer (include/stabilizer.hpp:106-475)."""

    def __init__(self, past_frames: int = 15, future_frames: int = 15,
                 working_height: int = 360, faithful_waste: bool = True,
                 exact_sift_matcher: bool = False, trail: bool = False, partial_lock_fix: bool = False):
        # stabilizer.cpp:36-53
        if past_frames == 0 and future_frames == 0:
            raise ValueError("Stabilizer: pastFrames and futureFrames cannot both be 0")
        if working_height <= 90.0:
            raise ValueError("Stabilizer: workingHeight must be greater than 90")
        if working_height > 2160:
            raise ValueError("Stabilizer: workingHeight must be no more than 2160")
        self.P, self.F, self.wh = past_frames, future_frames, working_height
        self.faithful_waste = faithful_waste
        self.exact_sift_matcher = exact_sift_matcher
        self.trail = trail                      # take the `#if 0` copyFeathered branch of stabilizeFrame (:1303-1307)
        # TRANSLATION_/ROTATION_LOCK: feed :1246-1260 with the accumulated lock instead of the identity :311-441 returns
        # in those modes (the hpp:23 @todo; vstab_set_partial_lock_fix).  Not reference behaviour: off by default.
        self.partial_lock_fix = partial_lock_fix
        self.scale = 1.0
        self.orig_size = (0, 0)      # (w, h)
        self.work_size = (0, 0)
        self.frames: deque = deque()           # (image, frame_idx)
        self.transforms: deque = deque()       # (H, from, to)
        self.prev_gray = None
        self.prev_pts = None
        self.trail_background = None
        self.acc_H = None
        self.acc_to = None
        self.mode = GLOBAL_SMOOTHING           # include/stabilizer.hpp:460
        self.reference_gray = None
        self.ref_kps = None
        self.ref_desc = None
        self.detector = None
        # function-static in the reference (stabilizer.cpp:446); per instance here
        self.previously_returned_H = np.eye(3)
        self.timers = {k: 0.0 for k in ("resize", "gray", "gftt", "lk", "ransac",
                                        "window", "lock", "mean", "warp", "clone")}
        self.taps = Taps()
        self.collect_taps = True

    # ---- API -------------------------------------------------------------
    def total_frame_window_size(self) -> int:   # include/stabilizer.hpp:196
        return self.P + 1 + self.F

    def set_stabilization_mode(self, mode: int) -> None:   # stabilizer.cpp:55-96
        self.reference_gray = None
        self.ref_kps = None
        self.ref_desc = None
        self.detector = None
        self.acc_H = None
        self.acc_to = None
        self.mode = mode

    # ---- helpers ---------------------------------------------------------
    def _initialize_frame(self, frame):          # stabilizer.cpp:98-126
        rows, cols = frame.shape[:2]
        if rows <= 10 or cols <= 10:
            raise ValueError("Stabilizer: Frame has invalid size")
        already = self.orig_size[0] > 0 and self.orig_size[1] > 0
        changed = self.orig_size != (cols, rows)
        if changed:
            if already:
                raise ValueError("Stabilizer: Frame size has changed. This is not supported.")
            self.orig_size = (cols, rows)
            self.scale = float(self.wh) / rows
            self.work_size = (int(cols * self.scale), self.wh)
        if (self.faithful_waste or self.trail) and (self.trail_background is None):
            self.trail_background = np.zeros_like(frame)             # :128-130

    def _detect_new_features(self, gray):        # stabilizer.cpp:931-980
        ratio = float(gray.shape[0]) / 720.0
        min_dist = int(10 * ratio)
        t0 = time.perf_counter()
        if self.faithful_waste:
            cv2.goodFeaturesToTrack(gray, 1300, 0.01, min_dist, mask=None, blockSize=3,
                                    gradientSize=3, useHarrisDetector=False, k=0.04)
        pts = cv2.goodFeaturesToTrack(gray, 1300, 0.01, min_dist)
        self.timers["gftt"] += time.perf_counter() - t0
        if pts is None:
            return np.zeros((0, 2), np.float32)
        return pts.reshape(-1, 2)

    def _track_features(self, prev_gray, gray, prev_pts):   # stabilizer.cpp:170-209
        if len(prev_pts) == 0:
            return prev_pts[:0], prev_pts[:0]
        t0 = time.perf_counter()
        cur, status, _err = cv2.calcOpticalFlowPyrLK(
            prev_gray, gray, prev_pts.reshape(-1, 1, 2), None, winSize=(21, 21), maxLevel=3,
            criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 50, 0.01),
            flags=0, minEigThreshold=1e-4)
        self.timers["lk"] += time.perf_counter() - t0
        cur = cur.reshape(-1, 2)
        status = status.reshape(-1)
        if self.collect_taps:
            self.taps.lk_pts = cur.copy()
            self.taps.lk_status = status.copy()
        keep = status == 1
        return prev_pts[keep], cur[keep]

    def _kill_scale(self, H):                    # stabilizer.cpp:261-272 / 752-758
        c = (self.work_size[0] / 2.0, self.work_size[1] / 2.0)
        p = decompose_homography(H, c)
        if p is None:
            return None
        p.s = 1.0
        return compose_homography(p, c)

    def _estimate_motion(self, prev, cur):       # stabilizer.cpp:211-275
        H = np.eye(3)
        self.taps.M = None
        self.taps.inliers = None
        if len(cur) < MIN_POINTS_FOR_MOTION_ESTIMATION:
            return H
        t0 = time.perf_counter()
        M, inl = cv2.estimateAffinePartial2D(prev.reshape(-1, 1, 2), cur.reshape(-1, 1, 2),
                                             method=cv2.RANSAC)
        self.timers["ransac"] += time.perf_counter() - t0
        if M is None or not np.all(np.isfinite(M)):
            return H
        if self.collect_taps:
            self.taps.M = M.copy()
            self.taps.inliers = None if inl is None else inl.reshape(-1).copy()
        H[0:2, :] = M
        H2 = self._kill_scale(H)
        return np.eye(3) if H2 is None else H2

    def _global_smoothing(self, p):              # stabilizer.cpp:793-852
        Hs = np.eye(3)
        avg = np.zeros((3, 3))
        count = 0
        acc = np.eye(3)
        for i in range(p, 0, -1):
            ok, inv = cv2.invert(self.transforms[i - 1][0])
            acc = inv @ acc
            avg += acc
            count += 1
        acc = np.eye(3)
        for i in range(p, len(self.transforms) - 1):
            acc = acc @ self.transforms[i][0]
            avg += acc
            count += 1
        if count > 0:
            avg = avg / count
            if np.all(np.isfinite(avg)):
                Hs = avg
        return Hs

    def _preprocess_for_features(self, frame):   # stabilizer.cpp:448-477
        r = cv2.resize(frame, self.work_size, interpolation=cv2.INTER_NEAREST)
        g = cv2.cvtColor(r, cv2.COLOR_BGR2GRAY)
        g = cv2.medianBlur(g, 5)
        k = np.array([[0, -1, 0], [-1, 5, -1], [0, -1, 0]], np.float32)
        g = cv2.filter2D(g, -1, k)
        clahe = cv2.createCLAHE()
        clahe.setClipLimit(2.0)
        clahe.setTilesGridSize((8, 8))
        g = clahe.apply(g)
        g = cv2.medianBlur(g, 5)
        return g

    def _full_lock(self, p):                     # stabilizer.cpp:311-791
        if self.mode == GLOBAL_SMOOTHING:
            return np.eye(3)
        if self.mode == ACCUMULATED_FULL_LOCK or (self.partial_lock_fix and self.mode in (TRANSLATION_LOCK, ROTATION_LOCK)):
            fidx = self.frames[p][1]
            if self.acc_H is None:
                self.acc_H = np.eye(3)
                self.acc_to = fidx
            else:
                assert p > 0, "ACCUMULATED_FULL_LOCK before the window can advance (Appendix B.6)"
                H, frm, to = self.transforms[p - 1]
                assert frm == self.acc_to
                self.acc_H = H @ self.acc_H
                self.acc_to = to
            return cv2.invert(self.acc_H)[1]                     # :438 Mat::inv() == LU
        if self.mode in (ORB_FULL_LOCK, SIFT_FULL_LOCK):
            return self._feature_lock(p)
        return np.eye(3)

    def _feature_lock(self, p):                  # stabilizer.cpp:440-787
        cur_gray = self._preprocess_for_features(self.frames[p][0])
        is_orb = self.mode == ORB_FULL_LOCK
        ratio = 0.10 if is_orb else 0.05
        if self.reference_gray is None:
            self.reference_gray = cur_gray.copy()
            self.previously_returned_H = np.eye(3)
            if is_orb:
                self.detector = cv2.ORB_create(2500, 1.2, 12, 31, 0, 2, cv2.ORB_FAST_SCORE, 31, 20)
            else:
                self.detector = cv2.SIFT_create(2500, 3, 0.04, 5, 1.2)
            kps, desc = self.detector.detectAndCompute(self.reference_gray, None)
            self.ref_kps, self.ref_desc = filter_keypoints_by_relative_size(
                self.reference_gray.shape[0], list(kps), desc, ratio)
            return np.eye(3)
        kps, desc = self.detector.detectAndCompute(cur_gray, None)
        kps, desc = filter_keypoints_by_relative_size(cur_gray.shape[0], list(kps), desc, ratio)
        if len(kps) < 10 or len(self.ref_kps) < 10:
            return self.previously_returned_H
        good = []
        if is_orb:
            m = cv2.BFMatcher(cv2.NORM_HAMMING)
            knn = m.knnMatch(self.ref_desc, desc, 2)
            for pair in knn:
                if len(pair) == 2 and pair[0].distance < np.float32(0.6) * np.float32(pair[1].distance):
                    good.append(pair[0])
        else:
            if self.exact_sift_matcher:
                matches = cv2.BFMatcher(cv2.NORM_L2).match(self.ref_desc, desc)
            else:
                matches = cv2.FlannBasedMatcher().match(self.ref_desc, desc)
            avg = sum(m.distance for m in matches) / len(self.ref_desc)
            thr = max(avg * 0.5, 0.02)
            good = [m for m in matches if m.distance <= thr]
        self.taps.n_matches = len(good)
        if len(good) < MIN_POINTS_FOR_MOTION_ESTIMATION:
            return self.previously_returned_H
        ref_pts = np.float32([self.ref_kps[m.queryIdx].pt for m in good])
        cur_pts = np.float32([kps[m.trainIdx].pt for m in good])
        M, _ = cv2.estimateAffinePartial2D(ref_pts.reshape(-1, 1, 2), cur_pts.reshape(-1, 1, 2),
                                           method=cv2.RANSAC, ransacReprojThreshold=5.0)
        if M is None or not np.all(np.isfinite(M)):
            return self.previously_returned_H
        H = np.eye(3)
        H[0:2, :] = M
        H2 = self._kill_scale(H)
        if H2 is not None:
            self.previously_returned_H = cv2.invert(H2)[1]
        return self.previously_returned_H

    # ---- the hot path ----------------------------------------------------
    def stabilize_frame(self, frame: np.ndarray) -> np.ndarray:   # stabilizer.cpp:1158-1325
        self._initialize_frame(frame)
        t0 = time.perf_counter()
        if self.faithful_waste and self.trail_background is not None:
            _ = self.trail_background.copy()                     # :1163 -> :129
        idx = self.frames[-1][1] + 1 if self.frames else 0       # :152-168
        self.frames.append((frame.copy(), idx))
        while len(self.frames) > self.total_frame_window_size():
            self.frames.popleft()
        self.timers["clone"] += time.perf_counter() - t0

        t0 = time.perf_counter()
        resized = cv2.resize(frame, self.work_size, interpolation=cv2.INTER_LINEAR)
        self.timers["resize"] += time.perf_counter() - t0
        t0 = time.perf_counter()
        gray = cv2.cvtColor(resized, cv2.COLOR_BGR2GRAY)
        self.timers["gray"] += time.perf_counter() - t0
        if self.collect_taps:
            self.taps = Taps(gray=gray)

        if self.prev_gray is None:                               # :1178-1182
            self.prev_pts = self._detect_new_features(gray)
            self.prev_gray = gray.copy()
            self.taps.new_pts = self.prev_pts
            return frame

        self.taps.prev_pts = self.prev_pts
        fprev, fcur = self._track_features(self.prev_gray, gray, self.prev_pts)
        T = self._estimate_motion(fprev, fcur)
        self.taps.T = T
        cur_idx = self.frames[-1][1]
        self.transforms.append((T.copy(), cur_idx - 1, cur_idx))  # :277-288
        while len(self.transforms) > self.total_frame_window_size() - 1:
            self.transforms.popleft()
        assert len(self.frames) == len(self.transforms) + 1
        assert self.frames[0][1] == self.transforms[0][1]

        p = 0                                                    # :1226-1229
        if len(self.frames) > self.F:
            p = len(self.frames) - self.F - 1

        t0 = time.perf_counter()
        H_smooth = self._global_smoothing(p)
        self.timers["window"] += time.perf_counter() - t0
        t0 = time.perf_counter()
        H_lock = self._full_lock(p)
        self.timers["lock"] += time.perf_counter() - t0

        # :1237-1260 -- T/R lock derivation (always evaluated; identity in practice)
        params_lock = decompose_homography(H_lock)
        if params_lock is None:
            H_lock = np.eye(3)
            params_lock = HomographyParameters()
        c = (self.work_size[0] / 2.0, self.work_size[1] / 2.0)
        R = cv2.getRotationMatrix2D(c, params_lock.theta * 180.0 / math.pi, 1.0)
        R_aug = np.eye(3)
        R_aug[0:2, :] = R
        H_translation_lock = R_aug @ H_lock
        H_rotation_lock = cv2.invert(R_aug)[1]

        if self.mode in (ACCUMULATED_FULL_LOCK, ORB_FULL_LOCK, SIFT_FULL_LOCK):
            H_stab = H_lock
        elif self.mode == TRANSLATION_LOCK:
            H_stab = H_translation_lock
        elif self.mode == ROTATION_LOCK:
            H_stab = H_rotation_lock
        elif self.mode == GLOBAL_SMOOTHING:
            H_stab = H_smooth
        else:
            raise ValueError("Stabilizer: Invalid stabilization mode")

        H_scaled = H_stab.copy()                                 # :1291-1296
        if abs(self.scale - 1.0) > 1e-6:
            H_scaled[0, 2] /= self.scale
            H_scaled[1, 2] /= self.scale

        pres = self.frames[p][0]
        t0 = time.perf_counter()
        m = cv2.mean(pres)                                       # :1309
        avg = tuple(0.5 * v for v in m)
        self.timers["mean"] += time.perf_counter() - t0
        t0 = time.perf_counter()
        if self.trail:                                           # the `#if 0` branch, :1303-1307
            out = copy_feathered(pres, self.trail_background, H_scaled)
            self.trail_background = out.copy()
        else:
            out = cv2.warpPerspective(pres, H_scaled, (frame.shape[1], frame.shape[0]),
                                      flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT,
                                      borderValue=avg)
        self.timers["warp"] += time.perf_counter() - t0

        self.prev_pts = self._detect_new_features(gray)          # :1318-1319
        t0 = time.perf_counter()
        self.prev_gray = gray.copy() if self.faithful_waste else gray
        self.timers["clone"] += time.perf_counter() - t0

        if self.collect_taps:
            self.taps.H_smooth = H_smooth
            self.taps.H_lock = H_lock
            self.taps.H_stabilize = H_stab
            self.taps.H_scaled = H_scaled
            self.taps.border = avg
            self.taps.presentation_idx = self.frames[p][1]
            self.taps.new_pts = self.prev_pts
        return out
