"""TEST INFRASTRUCTURE (oracle) -- integer/float restatements of the OpenCV
primitives on the reference hot path, in plain numpy.

The reference delegates all pixel arithmetic to OpenCV 4.x
(/root/reference/CMakeLists.txt:7, unpinned, not vendored).  These functions
restate the *published algorithms* of the exact calls the reference makes
(call sites cited per function) so that the CUDA kernels have a readable
specification; tests/test_oracle_restate.py proves each of them equal to
`cv2` 4.13.0 on seeded inputs (bit-exact unless a tolerance is stated).

Never imported by the product path.
"""
from __future__ import annotations

import math
import numpy as np

F32 = np.float32


# ----------------------------------------------------------------------------
# A.1  cvtColor(BGR2GRAY), 8U          call site: stabilizer.cpp:1175, :455
# ----------------------------------------------------------------------------
def bgr2gray(bgr: np.ndarray) -> np.ndarray:
    b = bgr[..., 0].astype(np.int32)
    g = bgr[..., 1].astype(np.int32)
    r = bgr[..., 2].astype(np.int32)
    return ((3735 * b + 19235 * g + 9798 * r + 16384) >> 15).astype(np.uint8)


# ----------------------------------------------------------------------------
# A.2  resize(INTER_LINEAR), 8UC3      call site: stabilizer.cpp:1170-1171
# ----------------------------------------------------------------------------
def working_size(rows: int, cols: int, working_height: int):
    """stabilizer.cpp:117-119: scaleFactor_ = wh/rows, (int(cols*s), wh)."""
    s = float(working_height) / rows
    return int(cols * s), working_height, s


def _linear_coeffs(src: int, dst: int):
    """Per-axis source index and Q11 coefficients of cv::resize INTER_LINEAR."""
    scale = np.float64(src) / np.float64(dst)
    d = np.arange(dst, dtype=np.float64)
    fx = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(fx).astype(np.int64)
    fx = (fx - s.astype(np.float32)).astype(np.float32)
    lo = s < 0
    s[lo] = 0
    fx[lo] = 0
    hi = s >= src - 1
    s[hi] = src - 1
    fx[hi] = 0
    c1 = np.rint(fx * F32(2048)).astype(np.int32)
    c0 = np.rint((F32(1.0) - fx) * F32(2048)).astype(np.int32)
    s1 = np.minimum(s + 1, src - 1)
    return s, s1, c0, c1


def resize_linear_bgr(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    sh, sw = src.shape[:2]
    if (dw, dh) == (sw, sh):
        return src.copy()
    if sw == 2 * dw and sh == 2 * dh:
        # exact 2x decimation silently becomes INTER_AREA: rounded 2x2 box mean
        a = src.astype(np.int32)
        out = (a[0::2, 0::2] + a[0::2, 1::2] + a[1::2, 0::2] + a[1::2, 1::2] + 2) >> 2
        return out.astype(np.uint8)
    xs0, xs1, xc0, xc1 = _linear_coeffs(sw, dw)
    ys0, ys1, yc0, yc1 = _linear_coeffs(sh, dh)
    a = src.astype(np.int32)
    # horizontal pass on the two needed source rows (x2048)
    r0 = a[ys0][:, xs0] * xc0[None, :, None] + a[ys0][:, xs1] * xc1[None, :, None]
    r1 = a[ys1][:, xs0] * xc0[None, :, None] + a[ys1][:, xs1] * xc1[None, :, None]
    b0 = yc0[:, None, None]
    b1 = yc1[:, None, None]
    out = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def resize_nearest_bgr(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """resize(INTER_NEAREST), call site stabilizer.cpp:450-451."""
    sh, sw = src.shape[:2]
    fx = np.float64(sw) / dw   # inv_scale
    fy = np.float64(sh) / dh
    xi = np.minimum(np.floor(np.arange(dw) * fx).astype(np.int64), sw - 1)
    yi = np.minimum(np.floor(np.arange(dh) * fy).astype(np.int64), sh - 1)
    return src[yi][:, xi].copy()


# ----------------------------------------------------------------------------
# A.3  pyrDown, 8U   (inside calcOpticalFlowPyrLK, call site stabilizer.cpp:192)
# ----------------------------------------------------------------------------
def _reflect101(i: np.ndarray, n: int) -> np.ndarray:
    i = np.where(i < 0, -i, i)
    i = np.where(i >= n, 2 * n - 2 - i, i)
    return i


def pyr_down(img: np.ndarray) -> np.ndarray:
    h, w = img.shape
    dh, dw = (h + 1) // 2, (w + 1) // 2
    k = np.array([1, 4, 6, 4, 1], dtype=np.int32)
    a = img.astype(np.int32)
    xs = 2 * np.arange(dw)[:, None] + np.arange(-2, 3)[None, :]
    xs = _reflect101(xs, w)
    rows = (a[:, xs] * k[None, None, :]).sum(axis=2)          # h x dw
    ys = 2 * np.arange(dh)[:, None] + np.arange(-2, 3)[None, :]
    ys = _reflect101(ys, h)
    out = (rows[ys, :] * k[None, :, None]).sum(axis=1)        # dh x dw
    return ((out + 128) >> 8).astype(np.uint8)


def lk_pyramid(img: np.ndarray, max_level: int = 3):
    levels = [img]
    for _ in range(max_level):
        levels.append(pyr_down(levels[-1]))
    return levels


# ----------------------------------------------------------------------------
# A.5  goodFeaturesToTrack            call site: stabilizer.cpp:949-963
# ----------------------------------------------------------------------------
def _fma32(a, b, c):
    """Correctly rounded float32 fma via float64 (products of two float32 are
    exact in float64; the final sum rounds once to f64 then to f32 -- double
    rounding can differ from a true fma by 1 ulp in rare cases)."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def corner_min_eigen_val(gray: np.ndarray) -> np.ndarray:
    """cornerMinEigenVal(gray, blockSize=3, ksize=3) as computed by OpenCV's
    AVX2/FMA filter engine."""
    h, w = gray.shape
    scale = 1.0 / (4.0 * 3.0 * 255.0)      # 1/(2^(ksize-1) * blockSize * 255)
    k1 = F32(scale)
    k0 = F32(2 * scale)
    yy = _reflect101(np.arange(-1, h + 1), h)
    xx = _reflect101(np.arange(-1, w + 1), w)
    p = gray[yy][:, xx].astype(np.float32)                    # (h+2) x (w+2)
    # Dx: row filter [-1 0 1] (exact), column filter [1 2 1]*scale
    S = p[:, 2:] - p[:, :-2]                                  # (h+2) x w
    Dx = _fma32(S[:-2] + S[2:], np.broadcast_to(k1, S[:-2].shape), S[1:-1] * k0)
    # Dy: row filter [1 2 1]*scale, column filter [-1 0 1]
    t = p[:, :-2] * k1
    t = _fma32(p[:, 1:-1], np.broadcast_to(k0, t.shape), t)
    R = _fma32(p[:, 2:], np.broadcast_to(k1, t.shape), t)     # (h+2) x w
    # [probe] the FMA vector body of OpenCV's row filter covers the first floor(w/32)*32 columns
    # (cv2 4.13.0, AVX-512 dispatch); the remaining columns go through the scalar tail, which is
    # not contracted: ((p[x-1]*k1) + (p[x]*k0)) + (p[x+1]*k1).  640/1280/1920/3840 have no tail.
    tail = (w // 32) * 32
    if tail < w:
        R[:, tail:] = ((p[:, tail:w] * k1) + (p[:, tail + 1:w + 1] * k0)) + (p[:, tail + 2:w + 2] * k1)
    Dy = R[2:] - R[:-2]
    dxx = Dx * Dx
    dxy = Dx * Dy
    dyy = Dy * Dy

    def box3(m):
        yy = _reflect101(np.arange(-1, h + 1), h)
        xx = _reflect101(np.arange(-1, w + 1), w)
        q = m[yy][:, xx].astype(np.float64)
        acc = np.zeros((h, w), np.float64)
        for dy in range(3):
            for dx in range(3):
                acc += q[dy:dy + h, dx:dx + w]
        return acc.astype(np.float32)

    a = box3(dxx) * F32(0.5)
    b = box3(dxy)
    c = box3(dyy) * F32(0.5)
    d = a - c
    return ((a + c) - np.sqrt(d * d + b * b)).astype(np.float32)


def good_features_to_track(gray: np.ndarray, max_corners: int = 1300,
                           quality: float = 0.01, min_distance: float = 5.0,
                           eig: np.ndarray | None = None) -> np.ndarray:
    """Selection part of goodFeaturesToTrack (threshold, 3x3 local max, sort,
    greedy grid suppression).  Returns (N,2) float32 (x,y)."""
    if eig is None:
        eig = corner_min_eigen_val(gray)
    h, w = eig.shape
    max_val = float(eig.max())
    thr = F32(max_val * quality)
    e = np.where(eig > thr, eig, F32(0))
    # dilate 3x3 (border: replicate of -inf has no effect => use edge padding with 0-safe max)
    pad = np.pad(e, 1, mode="constant", constant_values=-np.inf)
    dil = e.copy()
    for dy in range(3):
        for dx in range(3):
            dil = np.maximum(dil, pad[dy:dy + h, dx:dx + w])
    cand = (e != 0) & (e == dil)
    cand[0, :] = cand[-1, :] = False
    cand[:, 0] = cand[:, -1] = False
    ys, xs = np.nonzero(cand)
    vals = e[ys, xs]
    addr = ys.astype(np.int64) * w + xs
    order = np.lexsort((-addr, -vals.astype(np.float64)))      # value desc, address desc
    ys, xs = ys[order], xs[order]
    if min_distance < 1:
        n = min(max_corners, len(ys)) if max_corners > 0 else len(ys)
        return np.stack([xs[:n], ys[:n]], axis=1).astype(np.float32)
    cell = int(round(min_distance))
    gw = (w + cell - 1) // cell
    gh = (h + cell - 1) // cell
    grid = [[] for _ in range(gw * gh)]
    md2 = min_distance * min_distance
    out = []
    for y, x in zip(ys.tolist(), xs.tolist()):
        xc, yc = x // cell, y // cell
        x1, y1 = max(0, xc - 1), max(0, yc - 1)
        x2, y2 = min(gw - 1, xc + 1), min(gh - 1, yc + 1)
        good = True
        for gy in range(y1, y2 + 1):
            for gx in range(x1, x2 + 1):
                for (px, py) in grid[gy * gw + gx]:
                    dx, dy = x - px, y - py
                    if dx * dx + dy * dy < md2:
                        good = False
                        break
                if not good:
                    break
            if not good:
                break
        if good:
            grid[yc * gw + xc].append((x, y))
            out.append((x, y))
            if max_corners > 0 and len(out) == max_corners:
                break
    return np.array(out, dtype=np.float32).reshape(-1, 2)


# ----------------------------------------------------------------------------
# A.4  calcOpticalFlowPyrLK           call site: stabilizer.cpp:192-195
# ----------------------------------------------------------------------------
def scharr_deriv(img: np.ndarray):
    """int16 (dx, dy) with REFLECT_101 borders inside the image."""
    h, w = img.shape
    yy = _reflect101(np.arange(-1, h + 1), h)
    xx = _reflect101(np.arange(-1, w + 1), w)
    p = img[yy][:, xx].astype(np.int32)
    dx = 3 * (p[:-2, 2:] - p[:-2, :-2]) + 10 * (p[1:-1, 2:] - p[1:-1, :-2]) + 3 * (p[2:, 2:] - p[2:, :-2])
    dy = 3 * (p[2:, :-2] - p[:-2, :-2]) + 10 * (p[2:, 1:-1] - p[:-2, 1:-1]) + 3 * (p[2:, 2:] - p[:-2, 2:])
    return dx.astype(np.int16), dy.astype(np.int16)


def _pad_reflect(img, pad):
    h, w = img.shape
    yy = _reflect101(np.arange(-pad, h + pad), h)
    xx = _reflect101(np.arange(-pad, w + pad), w)
    return img[yy][:, xx]


def _lk_chain_sum(terms: np.ndarray) -> np.ndarray:
    """float32 sums of k 21x21 term windows (k, 21, 21) in the accumulation order of OpenCV's
    LKTrackerInvoker (lkpyramid.cpp, CV_SIMD128 build: SSE2 baseline of the cv2 wheels): rows top to
    bottom; in a row the columns 0..15 go four at a time into the four lanes of a v_float32x4 (lane
    k takes the columns x % 4 == k, one float add per term), the columns 16..20 one by one into a
    scalar float; the result is scalar + ((lane0 + lane2) + (lane1 + lane3)) (v_reduce_sum, SSE).
    Pinned against cv2.calcOpticalFlowPyrLK: identical float32 bits (tests/test_oracle_restate.py).
    The CUDA tracker (csrc/lk.cu) replays exactly this order."""
    t = np.asarray(terms, np.float32)
    q = np.zeros((t.shape[0], 4), np.float32)
    tail = np.zeros(t.shape[0], np.float32)
    for y in range(t.shape[1]):
        for x0 in range(0, 16, 4):
            q = q + t[:, y, x0:x0 + 4]
        for x in range(16, t.shape[2]):
            tail = tail + t[:, y, x]
    return tail + ((q[:, 0] + q[:, 2]) + (q[:, 1] + q[:, 3]))


def _lk_b_terms(prod: np.ndarray) -> np.ndarray:
    """The float terms of the mismatch sums b1, b2 from the exact integer products I_t * Ix (or Iy),
    (k, 21, 21): in each 8-column block v_dotprod adds the products of the columns x and x + 4 as
    int32 before the conversion to float; `_lk_chain_sum` then sees that pair term in column x and a
    zero in column x + 4 (adding +0.0 changes nothing).  Columns 16..20: float(int product) each."""
    p = np.asarray(prod, np.int64)
    t = np.zeros(p.shape, np.float32)
    for b in (0, 8):
        t[:, :, b:b + 4] = (p[:, :, b:b + 4] + p[:, :, b + 4:b + 8]).astype(np.float32)
    t[:, :, 16:] = p[:, :, 16:].astype(np.float32)
    return t


def calc_optical_flow_pyr_lk(prev: np.ndarray, nxt: np.ndarray, pts: np.ndarray,
                             win: int = 21, max_level: int = 3, max_iter: int = 50,
                             eps: float = 0.01, min_eig: float = 1e-4):
    """Sparse pyramidal LK, control flow, fixed-point formats and float accumulation order of
    OpenCV's LKTrackerInvoker (`_lk_chain_sum`): bit-identical to
    cv2.calcOpticalFlowPyrLK.  Returns (next_pts float32 (N,2), status u8)."""
    half = (win - 1) * 0.5
    # buildOpticalFlowPyramid drops every level that is not larger than the window in both
    # dimensions ([probe] 128x96 with win 21 keeps levels 0..2), lowering the effective maxLevel
    hh, ww = prev.shape
    for lvl in range(1, max_level + 1):
        ww, hh = (ww + 1) // 2, (hh + 1) // 2
        if ww <= win or hh <= win:
            max_level = lvl - 1
            break
    pp = lk_pyramid(prev, max_level)
    np_ = lk_pyramid(nxt, max_level)
    N = len(pts)
    out = np.zeros((N, 2), np.float32)
    status = np.ones(N, np.uint8)
    W_BITS = 14
    FLT_SCALE = F32(1.0 / (1 << 20))
    PAD = win
    for level in range(max_level, -1, -1):
        I = pp[level]
        J = np_[level]
        rows, cols = I.shape
        dIx, dIy = scharr_deriv(I)
        Ipad = _pad_reflect(I, PAD).astype(np.int32)
        Jpad = _pad_reflect(J, PAD).astype(np.int32)
        dxpad = np.pad(dIx.astype(np.int32), PAD)
        dypad = np.pad(dIy.astype(np.int32), PAD)
        scale = F32(1.0 / (1 << level))
        for i in range(N):
            prevx = F32(pts[i, 0]) * scale
            prevy = F32(pts[i, 1]) * scale
            if level == max_level:
                nx, ny = prevx, prevy
            else:
                nx, ny = F32(out[i, 0] * F32(2.0)), F32(out[i, 1] * F32(2.0))
            out[i] = (nx, ny)
            px = F32(prevx - F32(half))
            py = F32(prevy - F32(half))
            ipx = int(math.floor(px))
            ipy = int(math.floor(py))
            if ipx < -win or ipx >= cols or ipy < -win or ipy >= rows:
                if level == 0:
                    status[i] = 0
                continue
            a = F32(px - F32(ipx))
            b = F32(py - F32(ipy))
            one = F32(1.0)
            s = F32(1 << W_BITS)
            iw00 = int(np.rint((one - a) * (one - b) * s))
            iw01 = int(np.rint(a * (one - b) * s))
            iw10 = int(np.rint((one - a) * b * s))
            iw11 = (1 << W_BITS) - iw00 - iw01 - iw10
            y0, x0 = ipy + PAD, ipx + PAD
            def interp(A, rnd, sh, y0=y0, x0=x0, w=(iw00, iw01, iw10, iw11)):
                return (A[y0:y0 + win, x0:x0 + win] * w[0] + A[y0:y0 + win, x0 + 1:x0 + win + 1] * w[1]
                        + A[y0 + 1:y0 + win + 1, x0:x0 + win] * w[2]
                        + A[y0 + 1:y0 + win + 1, x0 + 1:x0 + win + 1] * w[3] + rnd) >> sh
            Iw = interp(Ipad, 1 << (W_BITS - 5 - 1), W_BITS - 5)
            Ix = interp(dxpad, 1 << (W_BITS - 1), W_BITS)
            Iy = interp(dypad, 1 << (W_BITS - 1), W_BITS)
            fx, fy = Ix.astype(np.float32), Iy.astype(np.float32)
            A11 = F32(_lk_chain_sum((fx * fx)[None])[0] * FLT_SCALE)
            A12 = F32(_lk_chain_sum((fx * fy)[None])[0] * FLT_SCALE)
            A22 = F32(_lk_chain_sum((fy * fy)[None])[0] * FLT_SCALE)
            D = F32(A11 * A22 - A12 * A12)
            mineig = F32((A22 + A11 - np.sqrt(F32((A11 - A22) * (A11 - A22) + F32(4.0) * A12 * A12))) / F32(2 * win * win))
            if float(mineig) < min_eig or D < np.finfo(np.float32).eps:   # OpenCV compares in double
                if level == 0:
                    status[i] = 0
                continue
            D = F32(one / D)
            nx = F32(nx - F32(half))
            ny = F32(ny - F32(half))
            pdx = pdy = F32(0)
            for j in range(max_iter):
                inx = int(math.floor(nx))
                iny = int(math.floor(ny))
                if inx < -win or inx >= cols or iny < -win or iny >= rows:
                    if level == 0:
                        status[i] = 0
                    break
                a = F32(nx - F32(inx))
                b = F32(ny - F32(iny))
                w00 = int(np.rint((one - a) * (one - b) * s))
                w01 = int(np.rint(a * (one - b) * s))
                w10 = int(np.rint((one - a) * b * s))
                w11 = (1 << W_BITS) - w00 - w01 - w10
                Jw = interp(Jpad, 1 << (W_BITS - 5 - 1), W_BITS - 5, iny + PAD, inx + PAD, (w00, w01, w10, w11))
                diff = (Jw - Iw).astype(np.int64)
                bb = _lk_chain_sum(_lk_b_terms(np.stack([diff * Ix, diff * Iy])))
                b1 = F32(bb[0] * FLT_SCALE)
                b2 = F32(bb[1] * FLT_SCALE)
                ddx = F32((A12 * b2 - A22 * b1) * D)
                ddy = F32((A12 * b1 - A11 * b2) * D)
                nx = F32(nx + ddx)
                ny = F32(ny + ddy)
                out[i] = (F32(nx + F32(half)), F32(ny + F32(half)))
                if float(ddx) * float(ddx) + float(ddy) * float(ddy) <= eps * eps:   # Point2f::ddot -> double
                    break
                if j > 0 and float(abs(F32(ddx + pdx))) < 0.01 and float(abs(F32(ddy + pdy))) < 0.01:
                    out[i, 0] = F32(out[i, 0] - ddx * F32(0.5))
                    out[i, 1] = F32(out[i, 1] - ddy * F32(0.5))
                    break
                pdx, pdy = ddx, ddy
            # the `err` block of LKTrackerInvoker (the reference passes an err vector,
            # stabilizer.cpp:192-195): at level 0 a still-valid point whose final window origin
            # lies outside the image gets status 0 (its coordinates are kept).
            if level == 0 and status[i]:
                fx = int(math.floor(F32(out[i, 0] - F32(half))))
                fy = int(math.floor(F32(out[i, 1] - F32(half))))
                if fx < -win or fx >= cols or fy < -win or fy >= rows:
                    status[i] = 0
    return out, status


# ----------------------------------------------------------------------------
# A.11 warpPerspective(INTER_LINEAR, BORDER_CONSTANT), 8UC3   stabilizer.cpp:1311
# ----------------------------------------------------------------------------
def invert3x3(H: np.ndarray) -> np.ndarray:
    """Adjugate inverse in f64 (cv::invert uses LU for 3x3 via the same closed
    form in its small-matrix fast path: det + cofactors)."""
    a = H.astype(np.float64)
    d = (a[0, 0] * (a[1, 1] * a[2, 2] - a[1, 2] * a[2, 1])
         - a[0, 1] * (a[1, 0] * a[2, 2] - a[1, 2] * a[2, 0])
         + a[0, 2] * (a[1, 0] * a[2, 1] - a[1, 1] * a[2, 0]))
    if d == 0:
        return np.zeros((3, 3))
    d = 1.0 / d
    t = np.empty((3, 3))
    t[0, 0] = (a[1, 1] * a[2, 2] - a[1, 2] * a[2, 1]) * d
    t[0, 1] = (a[0, 2] * a[2, 1] - a[0, 1] * a[2, 2]) * d
    t[0, 2] = (a[0, 1] * a[1, 2] - a[0, 2] * a[1, 1]) * d
    t[1, 0] = (a[1, 2] * a[2, 0] - a[1, 0] * a[2, 2]) * d
    t[1, 1] = (a[0, 0] * a[2, 2] - a[0, 2] * a[2, 0]) * d
    t[1, 2] = (a[0, 2] * a[1, 0] - a[0, 0] * a[1, 2]) * d
    t[2, 0] = (a[1, 0] * a[2, 1] - a[1, 1] * a[2, 0]) * d
    t[2, 1] = (a[0, 1] * a[2, 0] - a[0, 0] * a[2, 1]) * d
    t[2, 2] = (a[0, 0] * a[1, 1] - a[0, 1] * a[1, 0]) * d
    return t


_WARP_TAB = None


def warp_bilinear_tab() -> np.ndarray:
    """32x32x4 int16 weights (order: (y0,x0),(y0,x1),(y1,x0),(y1,x1)) summing to
    32768, as built by OpenCV's initInterTab2D(INTER_LINEAR, fixpt=true)."""
    global _WARP_TAB
    if _WARP_TAB is not None:
        return _WARP_TAB
    tab1 = np.zeros((32, 2), np.float32)
    for i in range(32):
        x = F32(i) * F32(1.0 / 32)
        tab1[i, 0] = F32(1.0) - x
        tab1[i, 1] = x
    tab = np.zeros((32, 32, 4), np.int32)
    for i in range(32):
        for j in range(32):
            isum = 0
            v = np.zeros(4, np.int32)
            for k1 in range(2):
                for k2 in range(2):
                    f = F32(tab1[i, k1] * tab1[j, k2])
                    val = int(np.clip(np.rint(f * F32(32768)), -32768, 32767))
                    v[k1 * 2 + k2] = val
                    isum += val
            if isum != 32768:
                diff = isum - 32768
                # OpenCV adjusts the largest tap (diff>0) or the smallest (diff<0)
                # within the central 2x2 (the only taps for ksize 2)
                mk = 0
                if diff < 0:
                    for k in range(4):
                        if v[k] < v[mk]:
                            mk = k
                else:
                    for k in range(4):
                        if v[k] > v[mk]:
                            mk = k
                v[mk] -= diff
            tab[i, j] = v
    _WARP_TAB = tab
    return tab


def border_value(frame: np.ndarray):
    """0.5 * cv::mean(frame) per channel (stabilizer.cpp:1309), f64."""
    n = frame.shape[0] * frame.shape[1]
    s = frame.reshape(-1, frame.shape[2]).astype(np.int64).sum(axis=0)
    return tuple(0.5 * (float(v) / n) for v in s)


def warp_perspective_bgr(src: np.ndarray, H: np.ndarray, border) -> np.ndarray:
    h, w = src.shape[:2]
    Mi = invert3x3(H)
    bv = np.array([int(np.clip(np.rint(b), 0, 255)) for b in border[:3]], np.int32)
    tab = warp_bilinear_tab()
    xs = np.arange(w, dtype=np.float64)[None, :]
    ys = np.arange(h, dtype=np.float64)[:, None]
    # OpenCV evaluates per 32x32 block: X0 = M0*bx + M1*y + M2, then X0 + M0*x1.
    bx = (np.arange(w) // 32 * 32).astype(np.float64)[None, :]
    x1 = (np.arange(w) % 32).astype(np.float64)[None, :]
    X = (Mi[0, 0] * bx + Mi[0, 1] * ys + Mi[0, 2]) + Mi[0, 0] * x1
    Y = (Mi[1, 0] * bx + Mi[1, 1] * ys + Mi[1, 2]) + Mi[1, 0] * x1
    Wd = (Mi[2, 0] * bx + Mi[2, 1] * ys + Mi[2, 2]) + Mi[2, 0] * x1
    with np.errstate(divide="ignore", invalid="ignore"):
        Wd = np.where(Wd != 0, 32.0 / Wd, 0.0)
    fX = np.clip(X * Wd, -2147483648.0, 2147483647.0)
    fY = np.clip(Y * Wd, -2147483648.0, 2147483647.0)
    iX = np.rint(fX).astype(np.int64)
    iY = np.rint(fY).astype(np.int64)
    sx = (iX >> 5)
    sy = (iY >> 5)
    ax = (iX & 31)
    ay = (iY & 31)
    # remap stores coordinates as int16 (saturate_cast<short>)
    sx = np.clip(sx, -32768, 32767)
    sy = np.clip(sy, -32768, 32767)
    wt = tab[ay, ax]                                          # h x w x 4
    a = src.astype(np.int32)

    def tap(yy, xx):
        inside = (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w)
        v = a[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)]
        return np.where(inside[..., None], v, bv[None, None, :])

    out = (tap(sy, sx) * wt[..., 0:1] + tap(sy, sx + 1) * wt[..., 1:2]
           + tap(sy + 1, sx) * wt[..., 2:3] + tap(sy + 1, sx + 1) * wt[..., 3:4] + 16384) >> 15
    return np.clip(out, 0, 255).astype(np.uint8)


# ----------------------------------------------------------------------------
# Feathered trail compositing: Stabilizer::copyFeathered   stabilizer.cpp:1051-1155
# (behind `#if 0` at :1304 in the reference).  Integer restatements of the
# OpenCV pieces it calls, each pinned against cv2 in tests/test_oracle_restate.py.
# ----------------------------------------------------------------------------
# cv::GaussianBlur on CV_8U runs in 8.8 fixed point; these are the kernels OpenCV derives for
# ksize 7 / sigma 0 (the exact small-kernel table) and for ksize 101 / sigma 0 (sigma 15.5,
# quantised with error diffusion so that the taps sum to 256), read off cv2's impulse response.
GAUSS7_Q8 = np.array([8, 28, 56, 72, 56, 28, 8], np.int64)
GAUSS101_Q8 = np.array([0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 1, 0, 1, 0, 1, 1, 1, 1, 2, 1, 2, 1, 2, 3, 2, 3, 3, 3, 3, 4,
                        4, 4, 4, 5, 5, 5, 5, 6, 5, 6, 6, 7, 6, 7, 6, 7, 6, 7, 6, 7, 6, 7, 6, 6, 5, 6, 5, 5, 5, 5, 4, 4, 4, 4,
                        3, 3, 3, 3, 2, 3, 2, 1, 2, 1, 2, 1, 1, 1, 1, 0, 1, 0, 1, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0],
                       np.int64)


def _reflect101_any(i: np.ndarray, n: int) -> np.ndarray:
    i = np.asarray(i).copy()
    while ((i < 0) | (i >= n)).any():
        i = np.abs(i)
        i = np.where(i >= n, 2 * n - 2 - i, i)
    return i


def gaussian_blur_u8(img: np.ndarray, kq8: np.ndarray) -> np.ndarray:
    """cv::GaussianBlur(u8, Size(k, k), 0) with BORDER_REFLECT_101: rows in 8.8 fixed point, columns in
    16.16, rounded once: (sum + 2^15) >> 16."""
    h, w = img.shape
    r = len(kq8) // 2
    xi = _reflect101_any(np.arange(-r, w + r), w)
    yi = _reflect101_any(np.arange(-r, h + r), h)
    a = img.astype(np.int64)[:, xi]
    hs = sum(int(kq8[j]) * a[:, j:j + w] for j in range(len(kq8)))
    b = hs[yi]
    vs = sum(int(kq8[j]) * b[j:j + h] for j in range(len(kq8)))
    return ((vs + 32768) >> 16).astype(np.uint8)


def _cdiv(a: int, b: int) -> int:
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def _clip_line(w, h, x1, y1, x2, y2):
    """cv::clipLine (Cohen-Sutherland with truncating double arithmetic)."""
    right, bottom = w - 1, h - 1
    c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8
    c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8
    if (c1 & c2) == 0 and (c1 | c2) != 0:
        if c1 & 12:
            a = 0 if c1 < 8 else bottom
            x1 += int(float(a - y1) * (x2 - x1) / (y2 - y1)); y1 = a
            c1 = (x1 < 0) + (x1 > right) * 2
        if c2 & 12:
            a = 0 if c2 < 8 else bottom
            x2 += int(float(a - y2) * (x2 - x1) / (y2 - y1)); y2 = a
            c2 = (x2 < 0) + (x2 > right) * 2
        if (c1 & c2) == 0 and (c1 | c2) != 0:
            if c1:
                a = 0 if c1 == 1 else right
                y1 += int(float(a - x1) * (y2 - y1) / (x2 - x1)); x1 = a; c1 = 0
            if c2:
                a = 0 if c2 == 1 else right
                y2 += int(float(a - x2) * (y2 - y1) / (x2 - x1)); x2 = a; c2 = 0
    return (c1 | c2) == 0, x1, y1, x2, y2


def fill_convex_poly_rows(w: int, h: int, pts) -> np.ndarray:
    """cv::fillConvexPoly(mask, pts, 255) (8-connected, shift 0) as row spans: int32 [h, n + 1, 2] = per row the
    {first, last} filled column of the interior scan line (span 0) and of the run each of the n outline segments
    leaves on that row (first > last: empty).  The outline is drawn first with cv::line (clipLine, then Bresenham
    from the LEFT end point -- a clipped segment is displaced from the true edge by up to a pixel, so it cannot be
    merged into the interior span), then the interior scan lines with 16.16 fixed-point edges."""
    v = [(int(x), int(y)) for x, y in pts]
    n = len(v)
    sp = np.empty((h, n + 1, 2), np.int64)
    sp[:, :, 0] = w
    sp[:, :, 1] = -1

    def put(k, y, xa, xb):
        if 0 <= y < h and xb >= 0 and xa < w:
            sp[y, k, 0] = min(sp[y, k, 0], max(xa, 0)); sp[y, k, 1] = max(sp[y, k, 1], min(xb, w - 1))

    p0 = v[-1]
    for k, p in enumerate(v):
        ok, x0, y0, x1, y1 = _clip_line(w, h, p0[0], p0[1], p[0], p[1])
        if ok:
            dx, dy = x1 - x0, y1 - y0
            if dx < 0:
                x0, y0, dx, dy = x1, y1, -dx, -dy
            sy = 1 if dy >= 0 else -1
            dy = abs(dy)
            steep = dy > dx
            if steep:
                dx, dy = dy, dx
            err, plus, minus = dx - 2 * dy, 2 * dx, -2 * dy
            x, y = x0, y0
            for _ in range(dx + 1):
                put(k + 1, y, x, x)
                m = err < 0
                err += minus + (plus if m else 0)
                if steep:
                    y += sy; x += 1 if m else 0
                else:
                    x += 1; y += sy if m else 0
        p0 = p
    ys = [q[1] for q in v]
    xs = [q[0] for q in v]
    ymin, ymax = min(ys), max(ys)
    if n < 3 or max(xs) < 0 or ymax < 0 or min(xs) >= w or ymin >= h:
        return sp.astype(np.int32)
    ymax = min(ymax, h - 1)
    imin = ys.index(ymin)
    one = 1 << 16
    edge = [dict(idx=imin, di=1, x=-one, dx=0, ye=ymin), dict(idx=imin, di=n - 1, x=-one, dx=0, ye=ymin)]
    y, edges = ymin, n
    while True:
        for e in edge:
            if y >= e["ye"]:
                idx0, di = e["idx"], e["di"]
                idx = (idx0 + di) % n
                while True:
                    edges -= 1
                    if edges < 0:
                        break
                    ty = v[idx][1]
                    if ty > y:
                        xs_, xe_ = v[idx0][0] << 16, v[idx][0] << 16
                        e.update(ye=ty, dx=_cdiv((xe_ - xs_) * 2 + (ty - y), 2 * (ty - y)), x=xs_, idx=idx)
                        break
                    idx0, idx = idx, (idx + di) % n
        if edges < 0:
            break
        if y >= 0:
            l, r = (0, 1) if edge[0]["x"] <= edge[1]["x"] else (1, 0)
            put(0, y, (edge[l]["x"] + (one >> 1)) >> 16, (edge[r]["x"] + (one >> 1)) >> 16)
        edge[0]["x"] += edge[0]["dx"]; edge[1]["x"] += edge[1]["dx"]
        y += 1
        if y > ymax:
            break
    return sp.astype(np.int32)


def mask_from_row_spans(sp: np.ndarray, w: int) -> np.ndarray:
    xs = np.arange(w)[None, None, :]
    return np.where(((xs >= sp[:, :, 0:1]) & (xs <= sp[:, :, 1:2])).any(axis=1), 255, 0).astype(np.uint8)


def perspective_points(pts: np.ndarray, H: np.ndarray) -> np.ndarray:
    """cv::perspectiveTransform on CV_32FC2 (double accumulators, result cast to float) followed by the
    Point2f -> Point conversion of stabilizer.cpp:1107-1110 (cvRound: half to even)."""
    out = []
    for x, y in np.asarray(pts, np.float32):
        x, y = float(x), float(y)
        w = x * H[2, 0] + y * H[2, 1] + H[2, 2]
        if abs(w) > np.finfo(np.float32).eps:
            w = 1.0 / w
            fx, fy = F32((x * H[0, 0] + y * H[0, 1] + H[0, 2]) * w), F32((x * H[1, 0] + y * H[1, 1] + H[1, 2]) * w)
        else:
            fx = fy = F32(0)
        out.append((int(np.rint(fx)), int(np.rint(fy))))
    return np.array(out, np.int64)


def warp_perspective_gray(src: np.ndarray, H: np.ndarray) -> np.ndarray:
    """warpPerspective of a 1-channel u8 image, INTER_LINEAR, constant border 0 (same Q5 / Q15 path)."""
    return warp_perspective_bgr(np.repeat(src[:, :, None], 3, axis=2), H, (0.0, 0.0, 0.0))[:, :, 0]


def copy_feathered(fg: np.ndarray, bg: np.ndarray, H: np.ndarray) -> np.ndarray:
    """Stabilizer::copyFeathered (stabilizer.cpp:1051-1155) without OpenCV."""
    h, w = fg.shape[:2]
    warped = warp_perspective_bgr(fg, H, (0.0, 0.0, 0.0)).astype(np.float32)                 # :1069-1073
    g = bgr2gray(bg)                                                                         # :1078-1079
    g = gaussian_blur_u8(g, GAUSS7_Q8)                                                       # :1081-1082
    g = np.rint(g.astype(np.float32) * F32(0.99)).astype(np.uint8)                           # :1085
    bgf = g.astype(np.float32)[:, :, None]                                                   # :1087-1090 (3 equal channels)
    B = 10                                                                                   # :1096
    corners = np.float32([[B, B], [w - B, B], [w - B, h - B], [B, h - B]])
    poly = perspective_points(corners, H)                                                    # :1104-1110
    mask = mask_from_row_spans(fill_convex_poly_rows(w, h, poly), w)                         # :1113
    mask = warp_perspective_gray(mask, H)                                                    # :1116-1118 (warped AGAIN by H)
    alpha = gaussian_blur_u8(mask, GAUSS101_Q8).astype(np.float32) * F32(1.0 / 255.0)        # :1121-1129
    alpha = alpha[:, :, None]
    out = alpha * warped + (F32(1.0) - alpha) * bgf                                          # :1137-1147
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)                                    # :1151


# ----------------------------------------------------------------------------
# A.10 estimateAffinePartial2D -> closed-form LS similarity on an inlier set
# ----------------------------------------------------------------------------
def ls_similarity(src: np.ndarray, dst: np.ndarray) -> np.ndarray:
    """argmin_{a,b,tx,ty} sum |[a -b;b a] p + t - q|^2 (f64, closed form)."""
    p = src.astype(np.float64)
    q = dst.astype(np.float64)
    n = len(p)
    mp = p.mean(axis=0)
    mq = q.mean(axis=0)
    pc = p - mp
    qc = q - mq
    den = (pc * pc).sum()
    a = (pc[:, 0] * qc[:, 0] + pc[:, 1] * qc[:, 1]).sum() / den
    b = (pc[:, 0] * qc[:, 1] - pc[:, 1] * qc[:, 0]).sum() / den
    tx = mq[0] - (a * mp[0] - b * mp[1])
    ty = mq[1] - (b * mp[0] + a * mp[1])
    return np.array([[a, -b, tx], [b, a, ty]])


class CvRNG:
    """cv::RNG (multiply-with-carry), state seeded with (uint64)-1 by
    RANSACPointSetRegistrator::run."""
    COEFF = 4164903690

    def __init__(self, state: int = 0xFFFFFFFFFFFFFFFF):
        self.state = state if state else 0xFFFFFFFF

    def next(self) -> int:
        self.state = ((self.state & 0xFFFFFFFF) * self.COEFF + (self.state >> 32)) & 0xFFFFFFFFFFFFFFFF
        return self.state & 0xFFFFFFFF

    def uniform(self, a: int, b: int) -> int:
        return a if a == b else int(self.next() % (b - a) + a)


def ransac_update_num_iters(p: float, ep: float, model_points: int, max_iters: int) -> int:
    p = min(max(p, 0.0), 1.0)
    ep = min(max(ep, 0.0), 1.0)
    num = max(1.0 - p, 2.2250738585072014e-308)
    denom = 1.0 - (1.0 - ep) ** model_points
    if denom < 2.2250738585072014e-308:
        return 0
    num = math.log(num)
    denom = math.log(denom)
    if denom >= 0 or -num >= max_iters * (-denom):
        return max_iters
    return int(np.rint(num / denom))


def similarity_from_2(p0, p1, q0, q1):
    """AffinePartial2DEstimatorCallback::runKernel (2-point similarity), f64."""
    x1, y1, x2, y2 = float(p0[0]), float(p0[1]), float(p1[0]), float(p1[1])
    X1, Y1, X2, Y2 = float(q0[0]), float(q0[1]), float(q1[0]), float(q1[1])
    with np.errstate(divide="ignore", invalid="ignore"):
        d = np.float64(1.0) / np.float64((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2))
        S0 = d * ((X1 - X2) * (x1 - x2) + (Y1 - Y2) * (y1 - y2))
        S1 = d * ((Y1 - Y2) * (x1 - x2) - (X1 - X2) * (y1 - y2))
        S2 = d * ((Y1 - Y2) * (x1 * y2 - x2 * y1) - (X1 * y2 - X2 * y1) * (y1 - y2) - (X1 * x2 - X2 * x1) * (x1 - x2))
        S3 = d * (-(X1 - X2) * (x1 * y2 - x2 * y1) - (Y1 * x2 - Y2 * x1) * (x1 - x2) - (Y1 * y2 - Y2 * y1) * (y1 - y2))
    return np.array([[S0, -S1, S2], [S1, S0, S3]], dtype=np.float64)


def similarity_errors(M, src, dst):
    """Affine2DEstimatorCallback::computeError: model cast to float, float math."""
    F = M.astype(np.float32).reshape(-1)
    fx, fy = src[:, 0].astype(np.float32), src[:, 1].astype(np.float32)
    a = ((F[0] * fx + F[1] * fy) + F[2]) - dst[:, 0].astype(np.float32)
    b = ((F[3] * fx + F[4] * fy) + F[5]) - dst[:, 1].astype(np.float32)
    return (a * a + b * b).astype(np.float32)


def ransac_similarity(src, dst, thresh=3.0, max_iters=2000, confidence=0.99):
    """RANSACPointSetRegistrator::run for the 2-point similarity model.
    Returns (best_model 2x3 f64 or None, inlier mask u8)."""
    count = len(src)
    if count < 2:
        return None, np.zeros(count, np.uint8)
    rng = CvRNG()
    niters = max(max_iters, 1)
    best_mask = np.zeros(count, np.uint8)
    best_model = None
    max_good = 0
    t = np.float32(thresh * thresh)
    if count == 2:
        return similarity_from_2(src[0], src[1], dst[0], dst[1]), np.ones(2, np.uint8)
    it = 0
    while it < niters:
        i0 = rng.uniform(0, count)
        while True:
            i1 = rng.uniform(0, count)
            if i1 != i0:
                break
        M = similarity_from_2(src[i0], src[i1], dst[i0], dst[i1])
        with np.errstate(invalid="ignore", over="ignore"):
            err = similarity_errors(M, src, dst)
            mask = (err <= t).astype(np.uint8)
        good = int(mask.sum())
        if good > max(max_good, 1):
            best_mask, best_model, max_good = mask, M, good
            niters = ransac_update_num_iters(confidence, float(count - good) / count, 2, niters)
        it += 1
    if max_good > 0:
        return best_model, best_mask
    return None, np.zeros(count, np.uint8)


def estimate_affine_partial_2d(src, dst, thresh=3.0):
    """cv::estimateAffinePartial2D(RANSAC) = RANSAC consensus + LM refinement, the
    latter restated by its fixed point: closed-form LS similarity on the inliers."""
    M, mask = ransac_similarity(src, dst, thresh)
    if M is None:
        return None, mask
    if len(src) > 2 and mask.sum() > 0:
        M = ls_similarity(src[mask == 1], dst[mask == 1])
    return M, mask
