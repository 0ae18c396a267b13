"""Generates tests/golden/golden_v1.npz: known-answer vectors for the hot path.

The reference has no tests or fixtures and cannot be built offline, so the
vectors are outputs of the reference's own OpenCV call sites executed through cv2 4.13.0
(`oracle/stabilizer_ref.py`, which restates /root/reference/src/stabilizer.cpp line by
line).  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import camera_engine_ref as ce  # noqa: E402
from oracle import stabilizer_ref as sr  # noqa: E402
from oracle import synth  # noqa: E402


def main():
    out = {}
    tex = synth.make_texture(512, cache=False)
    out["cv2_version"] = np.array(cv2.__version__)
    # ---- a frame pair at 640x360 ------------------------------------------------------
    path = synth.camera_path(12)
    W, H = 640, 360
    f = [ce.render_frame(tex, path[i], W, H, synth.focal_for_width(W)) for i in range(2)]
    out["f0"], out["f1"] = f
    for wh in (180, 120, 100, 360):
        ww = int(W * (float(wh) / H))
        out[f"gray_wh{wh}"] = cv2.cvtColor(cv2.resize(f[0], (ww, wh), interpolation=cv2.INTER_LINEAR), cv2.COLOR_BGR2GRAY)
    out["sums_f0"] = f[0].reshape(-1, 3).astype(np.uint64).sum(axis=0)
    wh = 180
    g = [cv2.cvtColor(cv2.resize(x, (320, 180), interpolation=cv2.INTER_LINEAR), cv2.COLOR_BGR2GRAY) for x in f]
    out["g0"], out["g1"] = g
    lv = g[0]
    for l in (1, 2, 3):
        lv = cv2.pyrDown(lv)
        out[f"g0_pyr{l}"] = lv
    md = int(10 * (wh / 720.0))
    out["gftt_min_distance"] = np.array(md)
    out["eig0"] = cv2.cornerMinEigenVal(g[0], 3, ksize=3)
    pts = cv2.goodFeaturesToTrack(g[0], 1300, 0.01, md).reshape(-1, 2)
    out["corners0"] = pts
    cur, st, _ = cv2.calcOpticalFlowPyrLK(g[0], g[1], pts.reshape(-1, 1, 2), None, winSize=(21, 21), maxLevel=3,
                                          criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 50, 0.01),
                                          flags=0, minEigThreshold=1e-4)
    out["lk_pts"], out["lk_status"] = cur.reshape(-1, 2), st.reshape(-1)
    keep = st.reshape(-1) == 1
    M, inl = cv2.estimateAffinePartial2D(pts[keep].reshape(-1, 1, 2), cur.reshape(-1, 2)[keep].reshape(-1, 1, 2), method=cv2.RANSAC)
    out["M"], out["inliers"] = M, inl.reshape(-1)
    o = sr.StabilizerRef(4, 3, wh)
    o.work_size = (320, 180)
    Hm = np.eye(3)
    Hm[:2] = M
    out["T"] = o._kill_scale(Hm)
    # RANSAC with outliers (thr 3 and 5)
    rng = np.random.default_rng(11)
    p = np.stack([rng.uniform(0, 320, 700), rng.uniform(0, 180, 700)], 1).astype(np.float32)
    th = 0.02
    A = 1.01 * np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    q = (p @ A.T + np.array([3.0, -2.0]) + rng.normal(0, 0.3, (700, 2))).astype(np.float32)
    bad = rng.random(700) < 0.3
    q[bad] += rng.uniform(-30, 30, (bad.sum(), 2)).astype(np.float32)
    out["ransac_p"], out["ransac_q"] = p, q
    for thr in (3.0, 5.0):
        M2, inl2 = cv2.estimateAffinePartial2D(p.reshape(-1, 1, 2), q.reshape(-1, 1, 2), method=cv2.RANSAC, ransacReprojThreshold=thr)
        out[f"ransac_M_thr{int(thr)}"], out[f"ransac_inl_thr{int(thr)}"] = M2, inl2.reshape(-1)
    # ---- warp known answers -----------------------------------------------------------------
    Hs = [np.array([[np.cos(0.02), -np.sin(0.02), 7.3], [np.sin(0.02), np.cos(0.02), -4.6], [0, 0, 1.0]]),
          np.array([[0.98, 0.02, 5.5], [-0.015, 1.01, -3.25], [1e-5, -2e-5, 1.0]])]
    m = cv2.mean(f[0])
    bd = tuple(0.5 * v for v in m)
    out["warp_border"] = np.array(bd[:3])
    for i, Hw in enumerate(Hs):
        out[f"warp_H{i}"] = Hw
        out[f"warp_out{i}"] = cv2.warpPerspective(f[0], Hw, (W, H), flags=cv2.INTER_LINEAR,
                                                   borderMode=cv2.BORDER_CONSTANT, borderValue=bd)
    # ---- a short streaming run: 256x192 frames, wh 96, window 4/3, lock at call 7 -----------
    W2, H2 = 256, 192
    frames = [ce.render_frame(tex, path[i], W2, H2, synth.focal_for_width(W2)) for i in range(12)]
    out["clip"] = np.stack(frames)
    for name, lock_at in (("smooth", None), ("lock", 7)):
        s = sr.StabilizerRef(4, 3, 96)
        outs, Ts, Hs_ = [], [], []
        for i, fr in enumerate(frames):
            if lock_at is not None and i == lock_at:
                s.set_stabilization_mode(sr.ACCUMULATED_FULL_LOCK)
            outs.append(s.stabilize_frame(fr))
            Ts.append(np.eye(3) if i == 0 else s.taps.T)
            Hs_.append(np.eye(3) if i == 0 else s.taps.H_scaled)
        out[f"clip_{name}_out"] = np.stack(outs)
        out[f"clip_{name}_T"] = np.stack(Ts)
        out[f"clip_{name}_H"] = np.stack(Hs_)
    # ---- homography decomposition known answers -------------------------------------------------
    Hd = np.array([[0.99, -0.02, 3.0], [0.02, 0.99, -4.0], [1e-5, 2e-5, 1.0]])
    pr = sr.decompose_homography(Hd, (320.0, 180.0))
    out["decomp_H"] = Hd
    out["decomp_params"] = np.array([pr.s, pr.theta, pr.k, pr.delta, pr.t[0], pr.t[1], pr.v[0], pr.v[1]])
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst) // 1024, "KiB")


if __name__ == "__main__":
    main()
