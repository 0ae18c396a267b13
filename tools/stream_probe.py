"""Streaming calls on synthetic 1080p frames through pinned host buffers (the bench's e2e.streaming path): for ncu launch
lists of the per-frame kernels and for VSTAB_TRACE=1 (host split + device marks of a call, printed at destroy)."""
import ctypes as C
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "video-stabilization_b200", "python")]
import torch
import vstab_b200 as vs
from vstab_b200 import offline, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
lock = not (len(sys.argv) > 2 and sys.argv[2] == "smooth")     # "smooth": stay in GLOBAL_SMOOTHING
W, H = 1920, 1080
# 8 frames of the scripted path, rendered by the library's own simulator kernel (K13)
tex = torch.from_numpy(synth.make_texture(2048)).cuda()
clip = torch.empty((8, H, W, 3), dtype=torch.uint8, device="cuda")
offline.render_frames(tex, synth.camera_path(8), H, W, synth.focal_for_width(W), clip)
host = list(clip.cpu().numpy())
host = host + host[-2:0:-1]          # ping-pong
lib = vs.load_library()
lib.vstab_host_alloc.restype = C.c_void_p
nbytes = W * H * 3
hin = lib.vstab_host_alloc(len(host) * nbytes)
hout = lib.vstab_host_alloc(nbytes)
np.ctypeslib.as_array((C.c_uint8 * (len(host) * nbytes)).from_address(hin)).reshape(len(host), H, W, 3)[:] = np.stack(host)
st = vs.Stabilizer(60, 45, 360, device=0)
t0 = 0.0
for i in range(n):
    if i == 46 and lock:
        st.set_stabilization_mode(vs.ACCUMULATED_FULL_LOCK)
    if i == 50:
        st.synchronize(); t0 = time.perf_counter()
    st.stabilize_frame_ptr(hin + (i % len(host)) * nbytes, H, W, W * 3, hout, W * 3)
st.synchronize()
if n > 50:
    print("calls/s", (n - 50) / (time.perf_counter() - t0))
st.close()
print("done")
