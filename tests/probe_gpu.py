"""Diagnostic script (not a pytest): runs every kernel entry point and the streaming
pipeline against the oracle on a B200 and prints difference statistics.  Used through
`gpurun -- python tests/probe_gpu.py` while developing; the pass/fail gates live in
tests/test_gpu_*.py."""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "video-stabilization_b200", "python"))

import cv2  # noqa: E402
import vstab_b200 as vs  # noqa: E402
from oracle import camera_engine_ref as ce  # noqa: E402
from oracle import cv_restate as R  # noqa: E402
from oracle import stabilizer_ref as sr  # noqa: E402
from oracle import synth  # noqa: E402


def stage(name):
    def deco(fn):
        def run(*a, **k):
            t = time.time()
            try:
                fn(*a, **k)
                print(f"[{name}] done in {time.time() - t:.2f}s", flush=True)
            except Exception:
                print(f"[{name}] EXCEPTION", flush=True)
                traceback.print_exc()
        return run
    return deco


def frames_for(W, H, n, tex):
    path = synth.camera_path(n)
    return [ce.render_frame(tex, path[i], W, H, synth.focal_for_width(W)) for i in range(n)]


@stage("ingest")
def probe_ingest(tex):
    rng = np.random.default_rng(3)
    for (H, W, wh) in [(720, 1280, 360), (1080, 1920, 360), (2160, 3840, 360), (1080, 1920, 1080), (480, 854, 360), (250, 333, 100)]:
        src = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        g, sums = vs.k_ingest(src, wh)
        dw, dh, _ = R.working_size(H, W, wh)
        ref = cv2.cvtColor(cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR), cv2.COLOR_BGR2GRAY)
        rs = src.reshape(-1, 3).astype(np.uint64).sum(axis=0)
        print(f"  {(H, W, wh)} gray ndiff {(g != ref).sum()} max {np.abs(g.astype(int) - ref).max()} sums ok {bool((sums == rs).all())}")


@stage("pyramid")
def probe_pyramid():
    rng = np.random.default_rng(4)
    for shp in [(360, 640), (1080, 1920), (91, 173)]:
        g = rng.integers(0, 256, shp, dtype=np.uint8)
        outs = vs.k_pyramid(g)
        ref = g
        for l, o in enumerate(outs):
            ref = cv2.pyrDown(ref)
            print(f"  {shp} L{l + 1} shape {o.shape} ndiff {(o != ref).sum()}")


@stage("gftt")
def probe_gftt(tex):
    for (W, H, wh) in [(1280, 720, 360), (1920, 1080, 1080)]:
        f = frames_for(W, H, 1, tex)[0]
        dw, dh, _ = R.working_size(H, W, wh)
        g = cv2.cvtColor(cv2.resize(f, (dw, dh), interpolation=cv2.INTER_LINEAR), cv2.COLOR_BGR2GRAY)
        md = int(10 * dh / 720.0)
        t = time.time()
        pts, eig = vs.k_gftt(g, 1300, 0.01, md, want_eig=True)
        dt = time.time() - t
        eref = cv2.cornerMinEigenVal(g, 3, ksize=3)
        ref = cv2.goodFeaturesToTrack(g, 1300, 0.01, md).reshape(-1, 2)
        same = pts.shape == ref.shape and bool((pts == ref).all())
        common = len(set(map(tuple, pts.tolist())) & set(map(tuple, ref.tolist())))
        print(f"  {(W, H, wh)} eig ndiff {(eig != eref).sum()} maxrel {np.abs(eig - eref).max() / eref.max():.2e} "
              f"n {len(pts)}/{len(ref)} identical-list {same} common {common} ({dt * 1e3:.1f} ms incl. copies)")
        if not same and len(pts) == len(ref):
            bad = np.nonzero((pts != ref).any(axis=1))[0]
            print("   first mismatches", bad[:5], pts[bad[:3]], ref[bad[:3]])
    rng = np.random.default_rng(5)
    g = rng.integers(0, 256, (360, 640), dtype=np.uint8)
    pts = vs.k_gftt(g, 1300, 0.01, 5)
    ref = cv2.goodFeaturesToTrack(g, 1300, 0.01, 5).reshape(-1, 2)
    print(f"  noise n {len(pts)}/{len(ref)} identical {pts.shape == ref.shape and bool((pts == ref).all())}")


@stage("lk")
def probe_lk(tex):
    fs = frames_for(1280, 720, 3, tex)
    gs = [cv2.cvtColor(cv2.resize(f, (640, 360), interpolation=cv2.INTER_LINEAR), cv2.COLOR_BGR2GRAY) for f in fs]
    pts = cv2.goodFeaturesToTrack(gs[0], 1300, 0.01, 5).reshape(-1, 2)
    extra = np.float32([[2, 3], [637, 2], [5, 357], [638, 358], [0, 0], [320, 0]])
    pts = np.concatenate([pts[:1290], extra])
    for a, b in ((0, 1), (1, 2), (0, 2)):
        cur, st, _ = cv2.calcOpticalFlowPyrLK(gs[a], gs[b], pts.reshape(-1, 1, 2), None, winSize=(21, 21), maxLevel=3,
                                              criteria=(3, 50, 0.01), flags=0, minEigThreshold=1e-4)
        cur = cur.reshape(-1, 2)
        st = st.reshape(-1)
        mine, mst = vs.k_lk(gs[a], gs[b], pts)
        ok = (st == 1) & (mst == 1)
        d = np.abs(mine[ok] - cur[ok]).max(axis=1)
        print(f"  pair {(a, b)} status mismatches {(st != mst).sum()} tracked {st.sum()} maxdiff {d.max():.2e} "
              f"median {np.median(d):.2e} n>0.05 {(d > 0.05).sum()} n>1e-3 {(d > 1e-3).sum()}")


@stage("fit")
def probe_fit():
    rng = np.random.default_rng(6)
    for trial in range(4):
        n = 1200
        p = np.stack([rng.uniform(0, 640, n), rng.uniform(0, 360, n)], 1).astype(np.float32)
        th = rng.uniform(-0.03, 0.03)
        s = 1.0 + rng.uniform(-0.01, 0.01)
        A = s * np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        t = rng.uniform(-8, 8, 2)
        q = (p @ A.T + t + rng.normal(0, 0.05, (n, 2))).astype(np.float32)
        out = rng.random(n) < 0.08 * trial
        q[out] += rng.uniform(-40, 40, (out.sum(), 2)).astype(np.float32)
        st = np.ones(n, np.uint8)
        st[::17] = 0
        M, T, cnt = vs.k_fit(p, q, st, 640, 360)
        keep = st == 1
        Mr, inl = cv2.estimateAffinePartial2D(p[keep].reshape(-1, 1, 2), q[keep].reshape(-1, 1, 2), method=cv2.RANSAC)
        corners = np.array([[0, 0, 1], [640, 0, 1], [0, 360, 1], [640, 360, 1]], float).T
        dM = np.abs(M @ corners - Mr @ corners).max()
        o = sr.StabilizerRef(15, 15, 360)
        o.work_size = (640, 360)
        H = np.eye(3)
        H[:2] = Mr
        Tr = o._kill_scale(H)
        dT = np.abs((T @ corners)[:2] - (Tr @ corners)[:2]).max()
        print(f"  trial {trial} outliers {out.sum()} counts {cnt} cv inliers {int(inl.sum())} corner diff M {dM:.2e} T {dT:.2e}")


@stage("warp")
def probe_warp(tex):
    f = frames_for(1920, 1080, 1, tex)[0]
    rng = np.random.default_rng(7)
    noise = rng.integers(0, 256, (360, 642, 3), dtype=np.uint8)

    def rigid(th, tx, ty):
        return np.array([[np.cos(th), -np.sin(th), tx], [np.sin(th), np.cos(th), ty], [0, 0, 1.0]])
    for name, src in (("frame", f), ("noise", noise)):
        for H in (rigid(0.01, 3.3, -7.7), rigid(-0.05, 40.2, 11.9), np.eye(3),
                  np.array([[0.98, 0.02, 5.5], [-0.015, 1.01, -3.25], [1e-5, -2e-5, 1.0]])):
            m = cv2.mean(src)
            bd = tuple(0.5 * v for v in m)
            ref = cv2.warpPerspective(src, H, (src.shape[1], src.shape[0]), flags=cv2.INTER_LINEAR,
                                      borderMode=cv2.BORDER_CONSTANT, borderValue=bd)
            bv = [int(np.clip(np.rint(b), 0, 255)) for b in bd[:3]]
            mine = vs.k_warp(src, H, bv)
            d = np.abs(mine.astype(int) - ref)
            print(f"  {name} warp ndiff {(d > 0).sum()} max {d.max()}")


@stage("stream")
def probe_stream(tex, W=1280, H=720, wh=360, n=70, P=20, F=10, lock_at=None):
    fs = frames_for(W, H, n, tex)
    ref = sr.StabilizerRef(P, F, wh)
    st = vs.Stabilizer(P, F, wh)
    corners = np.array([[0, 0, 1], [W, 0, 1], [0, H, 1], [W, H, 1]], float).T
    worst = {"T": 0, "Hs": 0, "pix_max": 0, "pix_nd": 0, "lk": 0, "gftt_diff": 0, "status": 0}
    t_gpu = 0.0
    for i, f in enumerate(fs):
        if lock_at is not None and i == lock_at:
            ref.set_stabilization_mode(sr.ACCUMULATED_FULL_LOCK)
            st.set_stabilization_mode(vs.ACCUMULATED_FULL_LOCK)
        o_ref = ref.stabilize_frame(f)
        t = time.time()
        o = st.stabilize_frame(f)
        t_gpu += time.time() - t
        d = np.abs(o.astype(int) - o_ref)
        worst["pix_max"] = max(worst["pix_max"], int(d.max()))
        worst["pix_nd"] = max(worst["pix_nd"], int((d > 0).sum()))
        if i == 0:
            g = st.tap(vs.TAP_GRAY)
            print(f"  call0 gray ndiff {(g != ref.taps.gray).sum()} newpts same {np.array_equal(st.tap(vs.TAP_NEW_PTS), ref.taps.new_pts)}")
            continue
        tp = ref.taps
        newp = st.tap(vs.TAP_NEW_PTS)
        same_new = np.array_equal(newp, tp.new_pts)
        prevp = st.tap(vs.TAP_PREV_PTS)
        same_prev = np.array_equal(prevp, tp.prev_pts)
        T = st.tap(vs.TAP_T)
        Hs = st.tap(vs.TAP_H_SCALED)
        dT = np.abs((T @ corners * (wh / H))[:2] - (tp.T @ corners * (wh / H))[:2]).max()
        a = Hs @ corners
        b = tp.H_scaled @ corners
        dH = np.abs(a[:2] / a[2] - b[:2] / b[2]).max()
        worst["T"] = max(worst["T"], dT)
        worst["Hs"] = max(worst["Hs"], dH)
        if same_prev:
            lk = st.tap(vs.TAP_LK_PTS)
            ls = st.tap(vs.TAP_LK_STATUS)
            ok = (ls == 1) & (tp.lk_status == 1)
            worst["lk"] = max(worst["lk"], float(np.abs(lk[ok] - tp.lk_pts[ok]).max()))
            worst["status"] = max(worst["status"], int((ls != tp.lk_status).sum()))
        if not same_new:
            worst["gftt_diff"] += 1
        if i in (1, 2, n // 2, n - 1) or d.max() > 1 or dH > 0.1:
            print(f"  call {i} pres {st.presentation_index()}/{tp.presentation_idx} gray ndiff {(st.tap(vs.TAP_GRAY) != tp.gray).sum()} "
                  f"prev_same {same_prev} new_same {same_new} dT {dT:.2e} dH {dH:.2e} pix max {d.max()} nd {(d > 0).sum()} "
                  f"border {st.tap(vs.TAP_BORDER)} vs {[round(b, 2) for b in tp.border[:3]]} inl {st.tap(vs.TAP_INLIERS)}")
    print(f"  {W}x{H} wh{wh} n{n} lock_at {lock_at}: worst {worst}; gpu {t_gpu / n * 1e3:.2f} ms/frame (pageable host buffers)")


def main():
    print("lib", vs.LIB_PATH, "abi", vs.load_library().vstab_abi_version(), flush=True)
    tex = synth.make_texture()
    probe_ingest(tex)
    probe_pyramid()
    probe_gftt(tex)
    probe_lk(tex)
    probe_fit()
    probe_warp(tex)
    probe_stream(tex)
    probe_stream(tex, 1920, 1080, 360, 40, 12, 8, lock_at=20)


if __name__ == "__main__":
    main()
