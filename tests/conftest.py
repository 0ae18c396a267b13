"""pytest configuration: the `gpu` marker, import paths and shared fixtures."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "video-stabilization_b200", "python")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"))


@pytest.fixture(scope="session")
def texture_small():
    from oracle import synth
    return synth.make_texture(512)


@pytest.fixture(scope="session")
def texture():
    from oracle import synth
    return synth.make_texture(2048)


def render_clip(tex, W, H, n, start=0):
    from oracle import camera_engine_ref as ce
    from oracle import synth
    path = synth.camera_path(start + n)
    return [ce.render_frame(tex, path[start + i], W, H, synth.focal_for_width(W)) for i in range(n)]
