"""debug: where does the TMA warp schedule differ from cv2?  (run on the GPU box with VSTAB_WARP_VARIANT=1)"""
import os, sys
import numpy as np, cv2
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "video-stabilization_b200", "python")]
import vstab_b200 as vs

def rigid(th, tx, ty):
    return np.array([[np.cos(th), -np.sin(th), tx], [np.sin(th), np.cos(th), ty], [0, 0, 1.0]])

for shape in [(1080, 1920), (360, 642)]:
    rng = np.random.default_rng(1)
    src = rng.integers(0, 256, (shape[0], shape[1], 3), dtype=np.uint8)
    for Hm in (rigid(0.01, 3.3, -7.7), np.eye(3), rigid(-0.05, 40.2, 11.9)):
        bd = tuple(0.5 * v for v in cv2.mean(src))
        ref = cv2.warpPerspective(src, Hm, (shape[1], shape[0]), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=bd)
        bv = [int(np.clip(np.rint(b), 0, 255)) for b in bd[:3]]
        out = vs.k_warp(src, Hm, bv)
        bad = np.any(out != ref, axis=2)
        print(shape, "bad px", int(bad.sum()))
        if bad.any():
            ys, xs = np.nonzero(bad)
            tiles = sorted(set(zip((ys // 32).tolist(), (xs // 128).tolist())))
            print("  bad tiles (ty,tx):", tiles[:40], "n", len(tiles))
            y, x = ys[0], xs[0]
            print("  first bad", y, x, "out", out[y, x], "ref", ref[y, x], "zero?", bool((out[bad] == 0).all()))
            # per bad tile: fraction bad
            t = tiles[len(tiles) // 2]
            sub = bad[t[0] * 32:(t[0] + 1) * 32, t[1] * 128:(t[1] + 1) * 128]
            print("  tile", t, "bad frac", sub.mean(), "rows bad", np.nonzero(sub.any(axis=1))[0][:40], "cols bad", np.nonzero(sub.any(axis=0))[0][:20])
