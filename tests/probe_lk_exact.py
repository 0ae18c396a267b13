# GPU probe: LK kernel vs cv2 on several frame pairs, count bit-equal points
import sys, numpy as np, cv2
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))); sys.path.insert(0, __import__('os').path.join(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))), 'video-stabilization_b200', 'python'))
import vstab_b200 as vs
from oracle import camera_engine_ref as ce, synth
tex = synth.make_texture(2048); path = synth.camera_path(64)
for (W, H) in ((1280, 720), (1920, 1080)):
    for (i, j) in ((2, 3), (20, 21), (40, 43), (10, 10)):
        fr = [ce.render_frame(tex, path[k], W, H, synth.focal_for_width(W)) for k in (i, j)]
        gs = [cv2.cvtColor(cv2.resize(f, (640, 360), interpolation=cv2.INTER_LINEAR), cv2.COLOR_BGR2GRAY) for f in fr]
        pts = cv2.goodFeaturesToTrack(gs[0], 1300, 0.01, 5).reshape(-1, 2).copy()
        cur, st, _ = cv2.calcOpticalFlowPyrLK(gs[0], gs[1], pts.reshape(-1, 1, 2), None, winSize=(21, 21), maxLevel=3,
            criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 50, 0.01), flags=0, minEigThreshold=1e-4)
        cur = cur.reshape(-1, 2); st = st.ravel()
        mine, mst = vs.k_lk(gs[0], gs[1], pts)
        ok = st == 1
        eq = (mine[ok] == cur[ok]).all(axis=1)
        print(W, H, i, j, 'n', len(pts), 'status eq', np.array_equal(st, mst), 'bit-equal', int(eq.sum()), '/', int(ok.sum()),
              'max diff', float(np.abs(mine[ok] - cur[ok]).max()))
