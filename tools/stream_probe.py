"""Streaming calls on synthetic 1080p frames (for ncu launch lists of the per-frame kernels)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "video-stabilization_b200", "python")]
import torch
import vstab_b200 as vs

n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
W, H = 1920, 1080
dev = torch.device("cuda:0")
from oracle import synth as osynth, camera_engine_ref as ce
tex = osynth.make_texture(2048)
path = osynth.camera_path(8)
host = [ce.render_frame(tex, path[i], W, H, osynth.focal_for_width(W)) for i in range(8)]
host = host + host[-2:0:-1]          # ping-pong
st = vs.Stabilizer(60, 45, 360, device=0)
for i in range(n):
    if i == 46:
        st.set_stabilization_mode(vs.ACCUMULATED_FULL_LOCK)
    out = st.stabilize_frame(host[i % len(host)])
st.synchronize()
print("done", out.shape)
