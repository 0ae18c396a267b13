"""K12 on a stacked batch: `nframes` current descriptor sets (2500 x 128 u8 each) against one reference set in ONE launch of
l2_nn_i8_kernel (frames along grid.z).  Prints the device time per launch (CUDA events, vstab_k_l2match_batch) and the
achieved integer tensor throughput; under `ncu --set full -k regex:l2_nn_i8_kernel` it is the capture the tensor-pipe
utilisation in profiles/ comes from.  usage: l2_batch_probe.py [nframes] [rows]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "video-stabilization_b200", "python")]
import vstab_b200 as vs

nframes = int(sys.argv[1]) if len(sys.argv) > 1 else 64
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 2500
rng = np.random.default_rng(0)
# SIFT-like descriptors: non-negative integers, L2-normalised to 512 and clipped like cv::SIFT's
def descs(n):
    d = rng.gamma(0.6, 1.0, (n, 128))
    d = d / np.linalg.norm(d, axis=1, keepdims=True) * 512
    return np.clip(np.rint(d), 0, 255).astype(np.uint8)
ref = descs(rows)
curs = [descs(rows) for _ in range(nframes)]
bi, bd, ms = vs.k_l2match_batch(ref, curs, reps=1 if os.environ.get("NCU_ONE") else 20)
ops = 2.0 * rows * rows * 128 * nframes
# spot check of two frames against numpy
for k in (0, nframes - 1):
    dd = ((ref.astype(np.int64)[:, None, :] - curs[k].astype(np.int64)[None, :, :]) ** 2).sum(-1)
    assert np.array_equal(bd[k], dd.min(1)) and np.array_equal(bi[k], dd.argmin(1))
print(f"l2 batch: {nframes} frames x {rows} x {rows} x 128: {ms * 1e3:.1f} us per launch, {ops / (ms * 1e-3) / 1e12:.1f} TOP/s "
      f"(u8 x u8 -> s32 on tcgen05 kind::i8), {ms * 1e3 / nframes:.2f} us per frame")
