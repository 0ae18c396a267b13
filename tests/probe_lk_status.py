"""Diagnostic (not a pytest): prints every LK status mismatch against cv2 on the config-1 clip."""
import os, sys
import numpy as np, cv2
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "video-stabilization_b200", "python"))
import vstab_b200 as vs
from oracle import camera_engine_ref as ce, synth

tex = synth.make_texture(2048)
for (W, H, n) in ((1280, 720, 50), (1920, 1080, 36)):
    path = synth.camera_path(n)
    prev = None
    for i in range(n):
        f = ce.render_frame(tex, path[i], W, H, synth.focal_for_width(W))
        g = cv2.cvtColor(cv2.resize(f, (640, 360), interpolation=cv2.INTER_LINEAR), cv2.COLOR_BGR2GRAY)
        if prev is not None:
            pts = cv2.goodFeaturesToTrack(prev, 1300, 0.01, 5).reshape(-1, 2)
            ref, st, _ = cv2.calcOpticalFlowPyrLK(prev, g, pts.reshape(-1, 1, 2), None, winSize=(21, 21), maxLevel=3,
                                                  criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 50, 0.01), flags=0, minEigThreshold=1e-4)
            got, gst = vs.k_lk(prev, g, pts)
            st = st.reshape(-1); ref = ref.reshape(-1, 2)
            bad = np.where(st != gst)[0]
            ok = (st == 1) & (gst == 1)
            d = np.abs(got[ok] - ref[ok]).max()
            for b in bad:
                print(f"{W}x{H} frame {i} pt {b} prev {pts[b]} cv2 st {st[b]} -> {ref[b]}  gpu st {gst[b]} -> {got[b]}")
            print(f"{W}x{H} frame {i}: n {len(pts)} mismatches {len(bad)} maxdiff {d:.2e}", flush=True)
        prev = g
