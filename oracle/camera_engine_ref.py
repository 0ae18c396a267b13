"""TEST INFRASTRUCTURE (oracle) -- numpy restatement of the reference simulator.

Follows /root/reference/src/camera_engine.cpp:
  * rotationMatrix            :36-61   R = Rz(roll) * Rx(tilt) * Ry(pan)
  * RenderPixelLoopBody::op() :73-155  per-pixel ray / floor-plane intersection,
                                       fmod(fmod(x,1)+1,1) wrap, int() truncation,
                                       nearest texel, sky colour (230,216,173)
  * renderFrame               :158-172 cx = W/2.0, cy = H/2.0
Every arithmetic step is a separately rounded IEEE double operation in the same
order as the C++ source (numpy never fuses multiply-add), so texel choices are
identical to an x86-64 build of the reference (no FMA contraction there).

Not on the product path: only tests/, bench.py's cpu arm and smoke() use it.
"""
from __future__ import annotations

import math
import numpy as np

SKY_BGR = np.array([230, 216, 173], dtype=np.uint8)  # camera_engine.cpp:81


def rotation_matrix(pan: float, tilt: float, roll: float) -> np.ndarray:
    """camera_engine.cpp:36-61 (degrees in, 3x3 f64 out)."""
    p = pan * math.pi / 180.0
    t = tilt * math.pi / 180.0
    r = roll * math.pi / 180.0
    ry = np.array([[math.cos(p), 0, math.sin(p)], [0, 1, 0], [-math.sin(p), 0, math.cos(p)]])
    rx = np.array([[1, 0, 0], [0, math.cos(t), -math.sin(t)], [0, math.sin(t), math.cos(t)]])
    rz = np.array([[math.cos(r), -math.sin(r), 0], [math.sin(r), math.cos(r), 0], [0, 0, 1]])
    # cv::Mat operator* evaluates (rz*rx)*ry with plain dot products; numpy matmul on
    # 3x3 doubles gives the same values to the last bit only if no FMA is used by BLAS,
    # so do the products explicitly, k ascending, like cv::gemm's generic path.
    def mm(a, b):
        out = np.zeros((3, 3))
        for i in range(3):
            for j in range(3):
                s = 0.0
                for k in range(3):
                    s = s + a[i, k] * b[k, j]
                out[i, j] = s
        return out
    return mm(mm(rz, rx), ry)


def render_frame(texture: np.ndarray, pose, width: int, height: int,
                 focal: float, rows: slice | None = None) -> np.ndarray:
    """camera_engine.cpp:73-172.  pose = (x, y, z, pan, tilt, roll)."""
    cam_x, cam_y, cam_z, pan, tilt, roll = (float(v) for v in pose)
    R = rotation_matrix(pan, tilt, roll).reshape(-1)
    tex_rows, tex_cols = texture.shape[:2]
    cx = width / 2.0
    cy = height / 2.0
    aspect = float(tex_cols) / float(tex_rows)
    tile_w = 1.0
    tile_h = tile_w / aspect

    y0, y1 = (0, height) if rows is None else (rows.start, rows.stop)
    ys = np.arange(y0, y1, dtype=np.float64)[:, None]
    xs = np.arange(width, dtype=np.float64)[None, :]
    u = xs - cx
    v = ys - cy
    mag = np.sqrt(u * u + v * v + focal * focal)
    cdx = u / mag
    cdy = v / mag
    cdz = focal / mag
    dx = R[0] * cdx + R[1] * cdy + R[2] * cdz
    dy = R[3] * cdx + R[4] * cdy + R[5] * cdz
    dz = R[6] * cdx + R[7] * cdy + R[8] * cdz
    sky = (np.abs(dz) < 1e-9) | (dz * cam_z >= 0)
    dz_safe = np.where(sky, 1.0, dz)
    t = -cam_z / dz_safe
    wx = cam_x + t * dx
    wy = cam_y + t * dy
    tx = wx / tile_w
    ty = wy / tile_h
    tu = np.fmod(np.fmod(tx, 1.0) + 1.0, 1.0)
    tv = np.fmod(np.fmod(ty, 1.0) + 1.0, 1.0)
    ix = (tu * tex_cols).astype(np.int64)   # static_cast<int>: truncation, values >= 0
    iy = (tv * tex_rows).astype(np.int64)
    ix = np.clip(ix, 0, tex_cols - 1)
    iy = np.clip(iy, 0, tex_rows - 1)
    out = texture[iy, ix]
    out[sky] = SKY_BGR
    return np.ascontiguousarray(out)


class CameraEngineRef:
    """Minimal mirror of CameraEngine (include/camera_engine.hpp:35-263): the pose
    setters used by scripted paths plus renderFrame()."""

    def __init__(self, texture: np.ndarray, width: int = 1280, height: int = 720,
                 focal: float = 1000.0):
        self.texture = texture
        self.width, self.height, self.focal = width, height, focal
        # main.cpp:29-36 defaults
        self.pose = [0.5, -0.3, 0.7, 0.0, 180.0, 180.0]

    def set_pose(self, pose) -> None:
        self.pose = [float(v) for v in pose]

    def render_frame(self) -> np.ndarray:
        return render_frame(self.texture, self.pose, self.width, self.height, self.focal)
