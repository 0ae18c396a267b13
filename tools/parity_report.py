"""Full-length parity of the streaming Stabilizer (C ABI) against the oracle on BASELINE's configurations.
TEST INFRASTRUCTURE (same standing as tests/): oracle/ is imported here as the checker, never as the thing measured.
Runs on the GPU box.  Frames come from the K13 device renderer (bit-exact with the simulator restatement,
tests/test_gpu_offline.py::test_render_matches_camera_engine; every `--check-render`-th frame is re-rendered
by the numpy oracle here and compared), so 2000-frame 1080p clips are affordable.

  python tools/parity_report.py c1 [--frames 300]      1280x720 -> 360, GLOBAL_SMOOTHING, window 60/45
  python tools/parity_report.py c2 [--frames 2000]     1920x1080 -> 360, ACCUMULATED_FULL_LOCK at call 46, window 60/45
  python tools/parity_report.py c3 [--frames 120]      1920x1080 -> 1080, ORB_FULL_LOCK at call 46, window 60/45
  python tools/parity_report.py c4 [--frames 24]       3840x2160 -> 2160, SIFT_FULL_LOCK at call 9, window 8/4

Prints one JSON line per config: {config, frames, h_px (max corner reprojection difference of the scaled
stabilizing homography), h_px_last, t_px, t_bit_equal (calls whose 3x3 T has the oracle's bits), lk_bit_equal
(calls whose tracked points have the oracle's bits), max_lsb, px_gt1 (pixels off by more than 1 LSB, whole clip),
frames_gt1, frac_gt1 (worst frame), px_total}."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "video-stabilization_b200", "python"))

CONFIGS = {
    "c1": dict(W=1280, H=720, wh=360, P=60, F=45, mode="GLOBAL_SMOOTHING", lock_at=None, frames=300, drift=0.0015),
    "c2": dict(W=1920, H=1080, wh=360, P=60, F=45, mode="ACCUMULATED_FULL_LOCK", lock_at=46, frames=2000, drift=0.0),
    "c3": dict(W=1920, H=1080, wh=1080, P=60, F=45, mode="ORB_FULL_LOCK", lock_at=46, frames=120, drift=0.0),
    "c4": dict(W=3840, H=2160, wh=2160, P=8, F=4, mode="SIFT_FULL_LOCK", lock_at=9, frames=24, drift=0.0),
}


def corner_diff(Ha, Hb, W, H):
    c = np.array([[0, 0, 1], [W, 0, 1], [0, H, 1], [W, H, 1]], float).T
    a, b = Ha @ c, Hb @ c
    return float(np.abs(a[:2] / a[2] - b[:2] / b[2]).max())


def run(name, n_frames=None, check_render=97, chunk=64, verbose=False):
    import torch
    import vstab_b200 as vs
    from vstab_b200 import offline, synth
    from oracle import camera_engine_ref as ce, stabilizer_ref as sr

    cfg = dict(CONFIGS[name])
    n = n_frames or cfg["frames"]
    W, H, wh = cfg["W"], cfg["H"], cfg["wh"]
    tex = synth.make_texture(2048)
    poses = synth.camera_path(n, drift=cfg["drift"])
    focal = synth.focal_for_width(W)
    tex_d = torch.from_numpy(tex).cuda()
    buf = torch.empty((chunk, H, W, 3), dtype=torch.uint8, device="cuda")
    # SIFT: the oracle matches exactly (the reference's FLANN KD-trees are approximate and not reproducible)
    ref = sr.StabilizerRef(cfg["P"], cfg["F"], wh, exact_sift_matcher=cfg["mode"] == "SIFT_FULL_LOCK")
    st = vs.Stabilizer(cfg["P"], cfg["F"], wh)
    mode = getattr(sr, cfg["mode"])
    s = dict(config=name, frames=n, h_px=0.0, h_px_last=0.0, t_px=0.0, t_bit_equal=0, lk_bit_equal=0, calls=0, max_lsb=0,
             px_gt1=0, frames_gt1=0, frac_gt1=0.0, px_differ=0, px_total=0, render_checked=0)
    t0 = time.time()
    for base in range(0, n, chunk):
        m = min(chunk, n - base)
        offline.render_frames(tex_d, poses[base:base + m], H, W, focal, buf[:m])
        frames = buf[:m].cpu().numpy()
        for j in range(m):
            i = base + j
            f = frames[j]
            if check_render and i % check_render == 0:
                assert np.array_equal(f, ce.render_frame(tex, poses[i], W, H, focal)), f"device render != oracle render at frame {i}"
                s["render_checked"] += 1
            if cfg["lock_at"] is not None and i == cfg["lock_at"]:
                ref.set_stabilization_mode(mode)
                st.set_stabilization_mode(mode)
            want = ref.stabilize_frame(f)
            got = st.stabilize_frame(f)
            d = np.abs(got.astype(np.int16) - want.astype(np.int16))
            mx = int(d.max())
            g1 = int((d > 1).sum())
            s["max_lsb"] = max(s["max_lsb"], mx)
            s["px_gt1"] += g1
            s["frames_gt1"] += int(g1 > 0)
            s["frac_gt1"] = max(s["frac_gt1"], g1 / d.size)
            s["px_differ"] += int((d > 0).sum())
            s["px_total"] += d.size
            if i == 0:
                continue
            tp = ref.taps
            s["calls"] += 1
            Tm = st.tap(vs.TAP_T)
            s["t_px"] = max(s["t_px"], corner_diff(Tm, tp.T, st.working_size()[0], wh))
            s["t_bit_equal"] += int(np.array_equal(Tm, tp.T))
            hd = corner_diff(st.tap(vs.TAP_H_SCALED), tp.H_scaled, W, H)
            s["h_px"] = max(s["h_px"], hd)
            s["h_px_last"] = hd
            if np.array_equal(st.tap(vs.TAP_PREV_PTS), tp.prev_pts):
                ok = tp.lk_status == 1
                same = np.array_equal(st.tap(vs.TAP_LK_STATUS), tp.lk_status) and \
                    np.array_equal(st.tap(vs.TAP_LK_PTS)[ok], tp.lk_pts[ok])
                s["lk_bit_equal"] += int(same)
            if verbose and (g1 or i % 100 == 0):
                print(f"  frame {i}: max {mx} LSB, {g1} px > 1 LSB, h {hd:.2e} px", file=sys.stderr)
    st.close()
    s["seconds"] = round(time.time() - t0, 1)
    return s


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="+", choices=sorted(CONFIGS))
    ap.add_argument("--frames", type=int, default=None)
    ap.add_argument("--check-render", type=int, default=97)
    ap.add_argument("-v", action="store_true")
    a = ap.parse_args()
    for c in a.configs:
        print(json.dumps(run(c, a.frames, a.check_render, verbose=a.v)), flush=True)
