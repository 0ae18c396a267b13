"""N > 1 host logic on CPU: the frame-sharded offline driver (vstab_b200.offline) with the oracle
as the per-pair estimator and warper, world_size 2 over gloo, must reproduce the streaming
reference restatement for every call index -- including the shard boundary (halo frame), the
warm-up calls owned by rank 0 and the ACCUMULATED_FULL_LOCK anchor."""
import os
import socket

import cv2
import numpy as np
import pytest

from vstab_b200 import offline
from oracle import stabilizer_ref as sr

P, F, WH = 4, 3, 96


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_plan_covers_every_call():
    for n_total, world, fut in [(100, 8, 45), (1000, 8, 45), (37, 4, 5), (16, 2, 0), (9, 8, 3), (100000, 8, 45)]:
        shards = offline.plan_shards(n_total, world)
        assert shards[0][0] == 0 and shards[-1][1] == n_total
        assert all(a[1] == b[0] for a, b in zip(shards, shards[1:]))
        assert max(b - a for a, b in shards) - min(b - a for a, b in shards) <= 1
        assert max(b - a for a, b in shards) == offline.padded_shard_len(n_total, world)
        calls = []
        for (a, b) in shards:
            c0, c1 = offline.calls_of_shard(a, b, n_total, fut)
            calls += list(range(c0, c1))
            for c in range(c0, c1):
                assert a <= max(0, c - fut) < b          # the owner of the presented frame renders the call
        assert calls == list(range(n_total))


def test_c_abi_shard_plan_equals_python_plan():
    """vstab_offline_plan (what vstab_offline_run shards by, include/vstab.h) == the Python index algebra above."""
    import ctypes as C
    import vstab_b200 as vs
    lib = vs.load_library()
    for n_total, world, fut in [(100, 8, 45), (1000, 8, 45), (37, 4, 5), (16, 2, 0), (9, 8, 3), (100000, 8, 45), (5, 8, 2)]:
        shards = offline.plan_shards(n_total, world)
        for r, (a, b) in enumerate(shards):
            pl = vs.ShardPlan()
            assert lib.vstab_offline_plan(n_total, world, r, fut, C.byref(pl)) == 0
            assert (pl.first, pl.last) == (a, b)
            assert (pl.call_first, pl.call_last) == offline.calls_of_shard(a, b, n_total, fut)
    pl = vs.ShardPlan()
    assert lib.vstab_offline_plan(10, 2, 2, 3, C.byref(pl)) != 0 and lib.vstab_offline_plan(0, 1, 0, 3, C.byref(pl)) != 0


def test_c_abi_fused_plan_covers_every_call_once_with_complete_windows():
    """vstab_offline_fused_plan (the one-pass GLOBAL_SMOOTHING schedule of vstab_offline_run): on every rank the fused calls
    read only transforms the rank estimates itself, present frames that are still in the ring when their window completes,
    and together with the deferred head / tail calls make up the rank's calls exactly once."""
    import ctypes as C
    import vstab_b200 as vs
    lib = vs.load_library()
    cases = [(100, 1, 6, 4, 7), (100, 3, 6, 4, 7), (1000, 8, 60, 45, 32), (1000, 8, 60, 45, 128), (391, 2, 60, 45, 100),
             (41, 1, 7, 9, 2), (41, 2, 7, 9, 3), (64, 8, 5, 3, 1), (30, 4, 0, 5, 4), (30, 4, 5, 0, 4), (9, 8, 3, 2, 4),
             (100000, 8, 60, 45, 128)]
    for n_total, world, P, F, B in cases:
        seen = []
        for r in range(world):
            pl, fp = vs.ShardPlan(), vs.FusedPlan()
            assert lib.vstab_offline_plan(n_total, world, r, F, C.byref(pl)) == 0
            assert lib.vstab_offline_fused_plan(n_total, world, r, P, F, B, C.byref(fp)) == 0
            assert pl.call_first <= fp.fused_first <= fp.fused_last <= pl.call_last
            lag = fp.ring_chunks - 3
            assert lag == (0 if F <= 1 else -(-(F - 1) // B))
            step = max(1, (fp.fused_last - fp.fused_first) // 300)          # sample the long ranges
            for c in list(range(fp.fused_first, fp.fused_last, step)) + ([fp.fused_last - 1] if fp.fused_last > fp.fused_first else []):
                lo_t, hi_t = max(1, c - P - F + 1), c - 1                    # T indices the window of call c reads
                if hi_t >= lo_t:
                    assert lo_t >= pl.first and hi_t <= pl.last - 1, (n_total, world, r, c)
                p = max(0, c - F)
                assert pl.first <= p < pl.last
                # the call is issued after the chunk holding frame c - 1 (or chunk 0 for c = 0); its frame is <= lag chunks older
                k_issue = (max(c - 1, pl.first) - pl.first) // B
                assert k_issue - (p - pl.first) // B <= lag, (n_total, world, r, c)
            deferred = (fp.fused_first - pl.call_first) + (pl.call_last - fp.fused_last)
            if r > 0 and pl.call_last > pl.call_first:
                assert deferred <= max(P - 1, 0) + max(F - 1, 0)
            if world == 1:
                assert deferred == 0
            seen += [(pl.call_first, fp.fused_first), (fp.fused_first, fp.fused_last), (fp.fused_last, pl.call_last)]
        covered = sorted((a, b) for a, b in seen if b > a)
        assert covered[0][0] == 0 and covered[-1][1] == n_total
        assert all(covered[i][1] == covered[i + 1][0] for i in range(len(covered) - 1))
    fp = vs.FusedPlan()
    assert lib.vstab_offline_fused_plan(10, 2, 0, 3, 2, 0, C.byref(fp)) != 0


def test_frame_checksum_c_equals_numpy_and_is_position_sensitive():
    import ctypes as C
    import vstab_b200 as vs
    lib = vs.load_library()
    rng = np.random.default_rng(3)
    for (h, w) in [(37, 53), (16, 64), (9, 5), (1, 1)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        a = vs.frame_checksum(img)
        assert a == lib.vstab_frame_checksum(img.ctypes.data_as(C.c_void_p), h, w, w * 3)
        if h > 1:
            sw = img.copy(); sw[[0, 1]] = sw[[1, 0]]
            assert np.array_equal(sw, img) or vs.frame_checksum(sw) != a      # swapping two rows changes it
        if w > 4:
            sw = img.copy(); sw[:, [0, 4]] = sw[:, [4, 0]]
            assert np.array_equal(sw, img) or vs.frame_checksum(sw) != a      # and two columns of different groups


# ---- oracle-backed estimate / render callbacks ---------------------------------------------------
def _estimate(frames_local, first, halo):
    """T[first+i] from the pair (frame first+i-1, frame first+i): stabilizer.cpp:1169-1209."""
    s = sr.StabilizerRef(P, F, WH)
    s._initialize_frame(frames_local[0])
    gray = lambda f: cv2.cvtColor(cv2.resize(f, s.work_size, interpolation=cv2.INTER_LINEAR), cv2.COLOR_BGR2GRAY)
    T = np.tile(np.eye(3).reshape(1, 9), (len(frames_local), 1))
    sums = np.stack([f.reshape(-1, 3).astype(np.int64).sum(0) for f in frames_local])
    prev = gray(halo) if halo is not None else None
    for i, f in enumerate(frames_local):
        g = gray(f)
        if prev is not None:
            pts = s._detect_new_features(prev)
            a, b = s._track_features(prev, g, pts)
            T[i] = s._estimate_motion(a, b).reshape(9)
        prev = g
    return T, sums


def _h_for_call(T, c, mode, lock_call):
    W = P + 1 + F
    lo, p = max(0, c - W + 1), max(0, c - F)
    if mode == sr.ACCUMULATED_FULL_LOCK and c >= lock_call:
        acc = np.eye(3)
        for k in range(lock_call - F + 1, p + 1):
            acc = T[k] @ acc
        return cv2.invert(acc)[1], p
    acc, total, count = np.eye(3), np.zeros((3, 3)), 0
    for k in range(p, lo, -1):
        acc = cv2.invert(T[k])[1] @ acc
        total += acc
        count += 1
    acc = np.eye(3)
    for k in range(p + 1, c):
        acc = acc @ T[k]
        total += acc
        count += 1
    return (total / count if count else np.eye(3)), p


def _render(frames_local, frame_base, call_first, ncalls, T_all, sums_local, mode, lock_call):
    T = T_all.reshape(-1, 3, 3)
    rows, cols = frames_local[0].shape[:2]
    scale = WH / rows
    outs = []
    for c in range(call_first, call_first + ncalls):
        H, p = _h_for_call(T, c, mode, lock_call)
        H = H.copy()
        H[0, 2] /= scale
        H[1, 2] /= scale
        fr = frames_local[p - frame_base]
        border = tuple(0.5 * v / (rows * cols) for v in sums_local[p - frame_base]) + (0.0,)
        outs.append(fr if c == 0 else cv2.warpPerspective(fr, H, (cols, rows), flags=cv2.INTER_LINEAR,
                                                          borderMode=cv2.BORDER_CONSTANT, borderValue=border))
    return outs


def _worker(rank, world, port, clip, mode, lock_call, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cv2.setNumThreads(1)
    n_total = len(clip)
    first, last = offline.plan_shards(n_total, world)[rank]
    pad = offline.padded_shard_len(n_total, world)

    def gather(T_local):
        Tp = torch.zeros((pad, 9), dtype=torch.float64)
        Tp[: T_local.shape[0]] = torch.from_numpy(T_local)
        return offline.gather_transforms(Tp, n_total, world).numpy()

    c0, outs = offline.stabilize_clip_sharded(
        clip[first:last], clip[first - 1] if first else None, n_total, rank, world, P, F,
        _estimate, lambda fr, fb, cf, nc, T, s: _render(fr, fb, cf, nc, T, s, mode, lock_call), gather,
        mode=mode, lock_call=lock_call)
    q.put((rank, c0, [np.asarray(o) for o in outs]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name,lock_at", [("smooth", None), ("lock", 7)])
def test_two_rank_gloo_equals_streaming(golden, name, lock_at):
    import torch.multiprocessing as mp
    clip = golden["clip"]
    mode = sr.GLOBAL_SMOOTHING if lock_at is None else sr.ACCUMULATED_FULL_LOCK
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, clip, mode, lock_at or 0, q)) for r in range(2)]
    for p_ in procs:
        p_.start()
    got = {}
    for _ in procs:
        rank, c0, outs = q.get(timeout=180)
        for j, o in enumerate(outs):
            assert c0 + j not in got
            got[c0 + j] = o
    for p_ in procs:
        p_.join(timeout=60)
        assert p_.exitcode == 0
    assert sorted(got) == list(range(len(clip)))
    for c in range(len(clip)):
        assert np.array_equal(got[c], golden[f"clip_{name}_out"][c]), c
