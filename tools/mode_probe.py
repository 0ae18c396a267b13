"""ORB / SIFT streaming calls (for ncu launch lists).  usage: mode_probe.py orb|sift [ncalls]"""
import os, sys, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "video-stabilization_b200", "python")]
import torch
import vstab_b200 as vs
from vstab_b200 import offline, synth

which = sys.argv[1]
ncalls = int(sys.argv[2]) if len(sys.argv) > 2 else 8
w, h, wh, mode = (1920, 1080, 1080, vs.ORB_FULL_LOCK) if which == "orb" else (3840, 2160, 2160, vs.SIFT_FULL_LOCK)
nd = 6
tex = torch.from_numpy(synth.make_texture(2048)).to("cuda:0")
frames = torch.empty((nd, h, w, 3), dtype=torch.uint8, device="cuda:0")
offline.render_frames(tex, synth.camera_path(nd, drift=0.0), h, w, synth.focal_for_width(w), frames, device=0)
host = frames.cpu().numpy()
st = vs.Stabilizer(60, 45, wh, device=0)
for i in range(ncalls):
    if i == 2:
        st.set_stabilization_mode(mode)
    out = st.stabilize_frame(host[i % nd])
st.synchronize()
print("done", st.tap(vs.TAP_ORB_COUNTS))
