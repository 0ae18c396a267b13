"""Synthetic inputs (seeded texture + scripted camera path) shared by the oracle and the
product side.  The generator itself lives with the package (vstab_b200/synth.py) because
bench.py's GPU arm needs it too and must not import from oracle/; it is input generation
only, no pixel of the hot path is computed there."""
import os
import sys

_PY = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "video-stabilization_b200", "python")
if _PY not in sys.path:
    sys.path.insert(0, _PY)

from vstab_b200.synth import *  # noqa: F401,F403,E402
from vstab_b200.synth import PATH_SEED, TEXTURE_SEED, camera_path, focal_for_width, make_texture  # noqa: F401,E402
