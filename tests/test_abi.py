"""The C ABI boundary (include/vstab.h): the library loads, exports every declared symbol,
the host-side logic behaves like the reference's argument checks, and -- without a GPU --
compute entry points fail loudly instead of falling back.  CPU only (no compute calls)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import __graft_entry__ as entry
import vstab_b200 as vs
from oracle import stabilizer_ref as sr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(vs.LIB_PATH):
        entry.build()
    return vs.load_library()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "vstab.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vstab_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/vstab.h but not exported"
        assert n in vs.SYMBOLS, f"{n} has no ctypes prototype"
    assert lib.vstab_abi_version() == 1


def test_header_is_plain_c(tmp_path):
    """include/vstab.h compiles as C99 with gcc: no C++/torch/OpenCV types at the boundary."""
    src = tmp_path / "t.c"
    src.write_text('#include "vstab.h"\nint main(void){return vstab_abi_version()*0;}\n')
    import subprocess
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-c", str(src), "-I", os.path.join(ROOT, "include"),
                        "-o", str(tmp_path / "t.o")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_cpp_wrapper_compiles(tmp_path):
    """include/stabilizer.hpp (the class-shaped mirror) compiles against the C ABI."""
    src = tmp_path / "t.cpp"
    src.write_text('#include "stabilizer.hpp"\nint main(){ try { Stabilizer s(0,0,360); } catch (const std::invalid_argument&) { return 0; } return 1; }\n')
    import subprocess
    exe = tmp_path / "t"
    r = subprocess.run(["g++", "-std=c++17", "-Wall", str(src), "-I", os.path.join(ROOT, "include"),
                        "-L", os.path.dirname(vs.LIB_PATH), "-lvstab", "-Wl,-rpath," + os.path.dirname(vs.LIB_PATH),
                        "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)])
    assert r.returncode == 0


def test_ctor_argument_checks_match_reference(lib):
    # src/stabilizer.cpp:40-49 -- these are rejected before any device is touched
    for args in ((0, 0, 360), (5, 5, 90), (5, 5, 2161)):
        with pytest.raises(ValueError):
            vs.Stabilizer(*args)


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(vs.VstabError, match="no CPU fallback"):
        vs.Stabilizer(15, 15, 360)
    with pytest.raises(vs.VstabError):
        vs.k_ingest(np.zeros((64, 64, 3), np.uint8), 32 + 64)


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(vs.VstabError, match="no CPU fallback"):
        vs.load_library(str(tmp_path / "nope.so"))


def test_decompose_compose_host_functions(lib, golden):
    H = golden["decomp_H"]
    p = vs.Stabilizer.decompose_homography(H, (320.0, 180.0))
    got = np.array([p.s, p.theta, p.k, p.delta, p.t[0], p.t[1], p.v[0], p.v[1]])
    assert np.allclose(got, golden["decomp_params"], rtol=0, atol=1e-13)
    assert np.abs(vs.Stabilizer.compose_homography(p, (320.0, 180.0)) - H).max() < 1e-12
    # degenerate inputs return false (None) like the reference, bad types raise
    assert vs.Stabilizer.decompose_homography(np.zeros((3, 3))) is None
    bad = np.eye(3)
    bad[0, 0] = -1.0
    assert vs.Stabilizer.decompose_homography(bad) is None
    with pytest.raises(ValueError):
        vs.Stabilizer.decompose_homography(np.eye(3, dtype=np.float32))
    rng = np.random.default_rng(3)
    for _ in range(50):
        th = rng.uniform(-3, 3)
        Hr = np.array([[np.cos(th), -np.sin(th), rng.normal(0, 50)], [np.sin(th), np.cos(th), rng.normal(0, 50)],
                       [rng.normal(0, 1e-5), rng.normal(0, 1e-5), 1.0]]) * rng.uniform(0.5, 2.0)
        c = (rng.uniform(0, 640), rng.uniform(0, 360))
        a = vs.Stabilizer.decompose_homography(Hr, c)
        b = sr.decompose_homography(Hr, c)
        assert (a is None) == (b is None)
        if a is not None:
            assert np.allclose([a.s, a.theta, a.k, a.delta, *a.t, *a.v], [b.s, b.theta, b.k, b.delta, *b.t, *b.v], rtol=1e-12, atol=1e-12)
            assert np.abs(vs.Stabilizer.compose_homography(a, c) - Hr / Hr[2, 2]).max() < 1e-9


def test_status_strings(lib):
    assert lib.vstab_status_string(0) == b"ok"
    assert b"size" in lib.vstab_status_string(2)


def _front_end():
    exe = os.path.join(os.path.dirname(os.path.dirname(vs.LIB_PATH)), "bin", "vstab_file")
    assert os.path.exists(exe), "video-stabilization_b200/bin/vstab_file is built by make (__graft_entry__.build())"
    return exe


def test_file_front_end_usage_and_argument_errors(tmp_path):
    """examples/vstab_file.cpp (the reference's --file main loop over include/stabilizer.hpp): usage errors and the
    constructor's std::invalid_argument (src/stabilizer.cpp:40-49) surface as exit status 1 before any device is used."""
    import subprocess
    exe = _front_end()
    assert subprocess.run([exe], capture_output=True).returncode == 1
    assert subprocess.run([exe, "--file", "-", "--width", "0", "--height", "4"], capture_output=True).returncode == 1
    clip = tmp_path / "c.bgr"
    clip.write_bytes(bytes(64 * 48 * 3))
    r = subprocess.run([exe, "--file", str(clip), "--width", "64", "--height", "48", "--working-height", "90",
                        "--out", str(tmp_path / "o.bgr")], capture_output=True, text=True)
    assert r.returncode == 1 and "Error" in r.stderr          # workingHeight <= 90 is rejected
    r = subprocess.run([exe, "--file", str(clip), "--width", "64", "--height", "48", "--past-window", "0",
                        "--future-window", "0", "--out", str(tmp_path / "o.bgr")], capture_output=True, text=True)
    assert r.returncode == 1


@pytest.mark.gpu
@pytest.mark.parametrize("side_by_side", [False, True])
def test_file_front_end_equals_streaming_api(tmp_path, side_by_side):
    """The C++ front end on a raw BGR24 clip writes exactly what the per-frame C ABI returns (mode switched to the
    full lock at call 5); --side-by-side pairs every stabilized frame with the original delayed by `future` frames."""
    import subprocess
    import numpy as np
    from conftest import render_clip
    from oracle import synth
    W, H, n, P, F, wh = 320, 240, 14, 4, 3, 120
    frames = render_clip(synth.make_texture(512), W, H, n)
    clip = tmp_path / "in.bgr"
    clip.write_bytes(b"".join(f.tobytes() for f in frames))
    out = tmp_path / "out.bgr"
    cmd = [_front_end(), "--file", str(clip), "--width", str(W), "--height", str(H), "--out", str(out),
           "--past-window", str(P), "--future-window", str(F), "--working-height", str(wh), "--mode", "lock", "--mode-at", "5"]
    r = subprocess.run(cmd + (["--side-by-side"] if side_by_side else []), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    st = vs.Stabilizer(P, F, wh)
    want = []
    for i, f in enumerate(frames):
        if i == 5:
            st.set_stabilization_mode(vs.ACCUMULATED_FULL_LOCK)
        want.append(st.stabilize_frame(f))
    st.close()
    got = np.frombuffer(out.read_bytes(), np.uint8)
    if not side_by_side:
        got = got.reshape(n, H, W, 3)
        for i in range(n):
            assert np.array_equal(got[i], want[i]), i
    else:
        got = got.reshape(n - F, H, 2 * W, 3)
        for j in range(n - F):
            assert np.array_equal(got[j][:, :W], frames[j]), j            # delayed original = presentation frame
            assert np.array_equal(got[j][:, W:], want[j + F]), j


def _offline_front_end():
    exe = os.path.join(os.path.dirname(os.path.dirname(vs.LIB_PATH)), "bin", "vstab_offline")
    assert os.path.exists(exe), "video-stabilization_b200/bin/vstab_offline is built by make (__graft_entry__.build())"
    return exe


def test_offline_front_end_usage_and_argument_errors(tmp_path):
    """examples/vstab_offline.cpp (the sharded job from a C++ host, one process per GPU): usage errors and the constructor's
    argument checks (src/stabilizer.cpp:40-49) surface as exit status 1 before any device is used; I/O errors as 2."""
    import subprocess
    exe = _offline_front_end()
    assert subprocess.run([exe], capture_output=True).returncode == 1
    clip = tmp_path / "c.bgr"
    clip.write_bytes(bytes(64 * 48 * 3 * 2))
    base = [exe, "--file", str(clip), "--width", "64", "--height", "48", "--out", str(tmp_path / "o.bgr")]
    r = subprocess.run(base + ["--working-height", "90"], capture_output=True, text=True)
    assert r.returncode == 1 and "workingHeight" in r.stderr
    assert subprocess.run(base + ["--past-window", "0", "--future-window", "0"], capture_output=True).returncode == 1
    assert subprocess.run(base + ["--world", "2", "--rank", "1"], capture_output=True).returncode == 1     # no --id-file
    assert subprocess.run(base + ["--world", "2", "--rank", "2", "--id-file", str(tmp_path / "id")], capture_output=True).returncode == 1
    assert subprocess.run(base + ["--mode", "nonsense"], capture_output=True).returncode == 1
    r = subprocess.run([exe, "--file", str(tmp_path / "missing.bgr"), "--width", "64", "--height", "48", "--out", str(tmp_path / "o.bgr")],
                       capture_output=True, text=True)
    assert r.returncode == 2
    r = subprocess.run([exe, "--file", str(clip), "--width", "640", "--height", "480", "--out", str(tmp_path / "o.bgr")],
                       capture_output=True, text=True)
    assert r.returncode == 2 and "no complete" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("mode,mode_at", [("global", 0), ("lock", 5)])
def test_offline_front_end_world_of_one_equals_streaming_api(tmp_path, mode, mode_at):
    """The C++ host of the sharded job (vstab_offline_run behind examples/vstab_offline.cpp) as a world of one: the raw
    BGR24 file it writes holds the frames the per-frame streaming API returns, and its checksum file their checksums.
    `global` takes the fused one-pass schedule (chunks of 4 frames, ring of 4 chunks), `lock` the two passes."""
    import subprocess
    import numpy as np
    from conftest import render_clip
    from oracle import synth
    W, H, n, P, F, wh = 320, 240, 19, 4, 3, 120
    frames = render_clip(synth.make_texture(512), W, H, n)
    clip = tmp_path / "in.bgr"
    clip.write_bytes(b"".join(f.tobytes() for f in frames))
    out, sums = tmp_path / "out.bgr", tmp_path / "sums.txt"
    cmd = [_offline_front_end(), "--file", str(clip), "--width", str(W), "--height", str(H), "--out", str(out),
           "--past-window", str(P), "--future-window", str(F), "--working-height", str(wh), "--batch", "4",
           "--mode", mode, "--mode-at", str(mode_at), "--checksums", str(sums), "--device", "0"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    st = vs.Stabilizer(P, F, wh)
    want = []
    for i, f in enumerate(frames):
        if mode == "lock" and i == mode_at:
            st.set_stabilization_mode(vs.ACCUMULATED_FULL_LOCK)
        want.append(st.stabilize_frame(f))
    st.close()
    got = np.frombuffer(out.read_bytes(), np.uint8).reshape(n, H, W, 3)
    lines = [ln.split() for ln in sums.read_text().splitlines()]
    assert [int(a) for a, _ in lines] == list(range(n))
    for i in range(n):
        assert np.array_equal(got[i], want[i]), i
        assert int(lines[i][1], 16) == vs.frame_checksum(want[i]), i
