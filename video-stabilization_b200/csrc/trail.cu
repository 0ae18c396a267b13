// K14: feathered trail compositing == Stabilizer::copyFeathered
//   /root/reference/src/stabilizer.cpp:1051-1155 (call site :1303-1307, behind `#if 0` in the reference; the author's
//   note at include/stabilizer.hpp:255-259 reserves it for "offline video processing or GPU-accelerated implementations").
// out = alpha * warp(fg, H) + (1 - alpha) * bg'   with
//   bg'   = 0.99 * GaussianBlur7x7(gray(bg))                      (:1078-1090, the trail fades and blurs a little every frame)
//   alpha = GaussianBlur101x101(warp(fillConvexPoly(H * inset corners), H)) / 255    (:1096-1129; the polygon is already in
//           output coordinates and is warped by H once more -- a quirk of the reference that is kept)
// Bit-exact with the OpenCV calls: every stage is integer arithmetic (Q15 luma, 8.8 / 16.16 fixed-point Gaussians with the
// tap tables cv::GaussianBlur derives for u8 images, the Q5 / Q15 warp of K7, cv::fillConvexPoly's clipLine + left-to-right
// Bresenham outline and 16.16 scan-line edges) up to the final float blend, whose four operations are individually rounded.
// Restated in numpy, and pinned against cv2, in oracle/cv_restate.py::copy_feathered.
#include "kernels.h"

namespace vstabk {
namespace {

__constant__ unsigned char kGauss101[101] = {
    0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 1, 0, 1, 0, 1, 1, 1, 1, 2, 1, 2, 1, 2, 3, 2, 3, 3, 3, 3, 4,
    4, 4, 4, 5, 5, 5, 5, 6, 5, 6, 6, 7, 6, 7, 6, 7, 6, 7, 6, 7, 6, 7, 6, 6, 5, 6, 5, 5, 5, 5, 4, 4, 4, 4,
    3, 3, 3, 3, 2, 3, 2, 1, 2, 1, 2, 1, 1, 1, 1, 0, 1, 0, 1, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0};

// ---- bg' : gray -> 7x7 Gaussian {8,28,56,72,56,28,8}/256 (rows 8.8, columns 16.16, one rounding) -> * 0.99 -----------------
constexpr int BT = 32;                         // output tile
__global__ void __launch_bounds__(256)
trail_bg_kernel(const uint8_t* __restrict__ bg, size_t pitch, int w, int h, uint8_t* __restrict__ out) {
    __shared__ unsigned char g[BT + 6][BT + 6];
    __shared__ unsigned short hs[BT + 6][BT];
    const int x0 = blockIdx.x * BT, y0 = blockIdx.y * BT;
    for (int i = threadIdx.x; i < (BT + 6) * (BT + 6); i += 256) {
        const int ly = i / (BT + 6), lx = i - ly * (BT + 6);
        const int sx = reflect101_multi(x0 + lx - 3, w), sy = reflect101_multi(y0 + ly - 3, h);
        const uint8_t* p = bg + (size_t)sy * pitch + (size_t)sx * 3;
        g[ly][lx] = (unsigned char)luma_q15(p[0], p[1], p[2]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (BT + 6) * BT; i += 256) {
        const int ly = i / BT, lx = i - ly * BT;
        hs[ly][lx] = (unsigned short)(8 * (g[ly][lx] + g[ly][lx + 6]) + 28 * (g[ly][lx + 1] + g[ly][lx + 5]) +
                                      56 * (g[ly][lx + 2] + g[ly][lx + 4]) + 72 * g[ly][lx + 3]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < BT * BT; i += 256) {
        const int ly = i / BT, lx = i - ly * BT;
        const int x = x0 + lx, y = y0 + ly;
        if (x >= w || y >= h) continue;
        const int v = 8 * (hs[ly][lx] + hs[ly + 6][lx]) + 28 * (hs[ly + 1][lx] + hs[ly + 5][lx]) +
                      56 * (hs[ly + 2][lx] + hs[ly + 4][lx]) + 72 * hs[ly + 3][lx];
        const int b = (v + 32768) >> 16;
        // background_image_changed *= 0.99: saturate_cast<uchar>(b * 0.99f), half to even
        out[(size_t)y * w + x] = (uint8_t)__float2int_rn(__fmul_rn((float)b, 0.99f));
    }
}

// ---- the polygon of the warped inset corners as row spans (cv::fillConvexPoly), one thread -------------------------------
// Per row 5 spans {first, last}: the interior scan line and the run each of the four outline segments leaves on that row
// (a clipped segment is displaced from the true edge by up to a pixel, so the runs cannot be merged into the interior span).
constexpr int kSpans = 5;
struct Edge { int idx, di; long long x, dx; int ye; };

__device__ void span_put(int2* rows, int k, int w, int h, int y, int xa, int xb) {
    if (y < 0 || y >= h || xb < 0 || xa >= w) return;
    xa = max(xa, 0); xb = min(xb, w - 1);
    int2& r = rows[(size_t)y * kSpans + k];
    r.x = min(r.x, xa);
    r.y = max(r.y, xb);
}
__device__ long long trunc_ll(double v) { return (long long)v; }

__global__ void trail_poly_kernel(const WarpParams* __restrict__ wp, int w, int h, int2* __restrict__ rows) {
    // rows were reset to {w, -1} (empty) by the caller
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const double* H = wp->Hs;
    const float B = 10.f;                                                   // BORDER_SIZE, :1096
    const float cxs[4] = {B, (float)w - B, (float)w - B, B}, cys[4] = {B, B, (float)h - B, (float)h - B};
    long long vx[4], vy[4];
    for (int i = 0; i < 4; ++i) {
        // cv::perspectiveTransform (double accumulation, float result), then cv::Point(Point2f) = cvRound
        const double x = cxs[i], y = cys[i];
        double ww = __dadd_rn(__dadd_rn(__dmul_rn(x, H[6]), __dmul_rn(y, H[7])), H[8]);
        float fx = 0.f, fy = 0.f;
        if (fabs(ww) > 1.1920928955078125e-07) {
            ww = __ddiv_rn(1.0, ww);
            fx = (float)__dmul_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, H[0]), __dmul_rn(y, H[1])), H[2]), ww);
            fy = (float)__dmul_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, H[3]), __dmul_rn(y, H[4])), H[5]), ww);
        }
        vx[i] = (long long)__float2int_rn(fx);
        vy[i] = (long long)__float2int_rn(fy);
    }
    // outline: cv::line = clipLine + 8-connected Bresenham from the left end point
    const long long right = w - 1, bottom = h - 1;
    for (int i = 0; i < 4; ++i) {
        long long x1 = vx[(i + 3) & 3], y1 = vy[(i + 3) & 3], x2 = vx[i], y2 = vy[i];
        int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
        int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            long long a;
            if (c1 & 12) {
                a = c1 < 8 ? 0 : bottom;
                x1 += trunc_ll(__ddiv_rn(__dmul_rn((double)(a - y1), (double)(x2 - x1)), (double)(y2 - y1)));
                y1 = a;
                c1 = (x1 < 0) + (x1 > right) * 2;
            }
            if (c2 & 12) {
                a = c2 < 8 ? 0 : bottom;
                x2 += trunc_ll(__ddiv_rn(__dmul_rn((double)(a - y2), (double)(x2 - x1)), (double)(y2 - y1)));
                y2 = a;
                c2 = (x2 < 0) + (x2 > right) * 2;
            }
            if ((c1 & c2) == 0 && (c1 | c2) != 0) {
                if (c1) {
                    a = c1 == 1 ? 0 : right;
                    y1 += trunc_ll(__ddiv_rn(__dmul_rn((double)(a - x1), (double)(y2 - y1)), (double)(x2 - x1)));
                    x1 = a; c1 = 0;
                }
                if (c2) {
                    a = c2 == 1 ? 0 : right;
                    y2 += trunc_ll(__ddiv_rn(__dmul_rn((double)(a - x2), (double)(y2 - y1)), (double)(x2 - x1)));
                    x2 = a; c2 = 0;
                }
            }
        }
        if ((c1 | c2) != 0) continue;
        long long dx = x2 - x1, dy = y2 - y1;
        long long x = x1, y = y1;
        if (dx < 0) { x = x2; y = y2; dx = -dx; dy = -dy; }
        const int sy = dy >= 0 ? 1 : -1;
        if (dy < 0) dy = -dy;
        const bool steep = dy > dx;
        if (steep) { const long long t = dx; dx = dy; dy = t; }
        long long err = dx - 2 * dy;
        const long long plus = 2 * dx, minus = -2 * dy;
        for (long long k = 0; k <= dx; ++k) {
            span_put(rows, 1 + i, w, h, (int)y, (int)x, (int)x);
            const bool m = err < 0;
            err += minus + (m ? plus : 0);
            if (steep) { y += sy; x += m ? 1 : 0; } else { x += 1; y += m ? sy : 0; }
        }
    }
    // interior: scan lines between two 16.16 fixed-point edges
    long long ymin = vy[0], ymax = vy[0], xmin = vx[0], xmax = vx[0];
    int imin = 0;
    for (int i = 0; i < 4; ++i) {
        if (vy[i] < ymin) { ymin = vy[i]; imin = i; }
        ymax = max(ymax, vy[i]); xmax = max(xmax, vx[i]); xmin = min(xmin, vx[i]);
    }
    if (xmax < 0 || ymax < 0 || xmin >= w || ymin >= h) return;
    ymax = min(ymax, (long long)h - 1);
    const long long one = 1ll << 16;
    Edge e[2] = {{imin, 1, -one, 0, (int)ymin}, {imin, 3, -one, 0, (int)ymin}};
    int edges = 4;
    long long y = ymin;
    do {
        for (int i = 0; i < 2; ++i) {
            if (y >= e[i].ye) {
                int idx0 = e[i].idx;
                const int di = e[i].di;
                int idx = (idx0 + di) & 3;
                for (; edges-- > 0;) {
                    const long long ty = vy[idx];
                    if (ty > y) {
                        const long long xs = vx[idx0] << 16, xe = vx[idx] << 16;
                        e[i].ye = (int)ty;
                        e[i].dx = ((xe - xs) * 2 + (ty - y)) / (2 * (ty - y));      // C++ division: toward zero
                        e[i].x = xs;
                        e[i].idx = idx;
                        break;
                    }
                    idx0 = idx;
                    idx = (idx + di) & 3;
                }
            }
        }
        if (edges < 0) break;
        if (y >= 0) {
            const int l = e[0].x > e[1].x ? 1 : 0, r = 1 - l;
            span_put(rows, 0, w, h, (int)y, (int)((e[l].x + (one >> 1)) >> 16), (int)((e[r].x + (one >> 1)) >> 16));
        }
        e[0].x += e[0].dx;
        e[1].x += e[1].dx;
    } while (++y <= ymax);
}

__global__ void trail_rows_reset_kernel(int2* rows, int w, int h) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < h * kSpans) rows[i] = make_int2(w, -1);
}

// ---- the mask, warped by H: warpPerspective of a 0 / 255 image given as row spans (INTER_LINEAR, constant border 0) --------
__global__ void __launch_bounds__(256)
trail_mask_kernel(const WarpParams* __restrict__ wp, const int2* __restrict__ rows, int w, int h, uint8_t* __restrict__ out) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= w || y >= h) return;
    const double* M = wp->Minv;
    // OpenCV evaluates per 32-px block: X0 = M0*bx + M1*y + M2, then X0 + M0*x1 (K7 does the same)
    const double bx = (double)(x & ~31), x1 = (double)(x & 31), yd = (double)y;
    const double X = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(M[0], bx), __dmul_rn(M[1], yd)), M[2]), __dmul_rn(M[0], x1));
    const double Y = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(M[3], bx), __dmul_rn(M[4], yd)), M[5]), __dmul_rn(M[3], x1));
    const double W = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(M[6], bx), __dmul_rn(M[7], yd)), M[8]), __dmul_rn(M[6], x1));
    const double Wd = W != 0.0 ? __ddiv_rn(32.0, W) : 0.0;
    const int iX = __double2int_rn(__dmul_rn(X, Wd)), iY = __double2int_rn(__dmul_rn(Y, Wd));
    const int sx = max(-32768, min(32767, iX >> 5)), sy = max(-32768, min(32767, iY >> 5));
    const int ax = iX & 31, ay = iY & 31;
    auto tap = [&](int yy, int xx) -> int {
        if (yy < 0 || yy >= h || xx < 0 || xx >= w) return 0;
        bool in = false;
#pragma unroll
        for (int k = 0; k < kSpans; ++k) { const int2 r = rows[(size_t)yy * kSpans + k]; in = in || (xx >= r.x && xx <= r.y); }
        return in ? 255 : 0;
    };
    const int v = tap(sy, sx) * (32 - ay) * (32 - ax) + tap(sy, sx + 1) * (32 - ay) * ax + tap(sy + 1, sx) * ay * (32 - ax) +
                  tap(sy + 1, sx + 1) * ay * ax;                       // weights / 32: (sum * 32 + 16384) >> 15 == (sum + 512) >> 10
    out[(size_t)y * w + x] = (uint8_t)((v + 512) >> 10);
}

// ---- 101 x 101 Gaussian of the warped mask: rows into 8.8, columns into 16.16, one rounding ---------------------------------
constexpr int GR = 50;
__global__ void __launch_bounds__(256)
trail_blur_rows_kernel(const uint8_t* __restrict__ src, int w, int h, unsigned short* __restrict__ dst) {
    __shared__ unsigned char row[256 + 2 * GR];
    const int y = blockIdx.y, x0 = blockIdx.x * 256;
    for (int i = threadIdx.x; i < 256 + 2 * GR; i += 256) row[i] = src[(size_t)y * w + reflect101_multi(x0 + i - GR, w)];
    __syncthreads();
    const int x = x0 + threadIdx.x;
    if (x >= w) return;
    int s = 0;
#pragma unroll 4
    for (int j = 7; j <= 93; ++j) s += (int)kGauss101[j] * (int)row[threadIdx.x + j];       // the taps outside 7..93 are zero
    dst[(size_t)y * w + x] = (unsigned short)s;
}
constexpr int VT = 64;                          // rows per tile
__global__ void __launch_bounds__(256)
trail_blur_cols_kernel(const unsigned short* __restrict__ src, int w, int h, uint8_t* __restrict__ dst) {
    __shared__ unsigned short col[VT + 2 * GR][32];
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * VT;
    const int lx = threadIdx.x & 31, ly0 = threadIdx.x >> 5;
    const int x = min(x0 + lx, w - 1);
    for (int r = ly0; r < VT + 2 * GR; r += 8) col[r][lx] = src[(size_t)reflect101_multi(y0 + r - GR, h) * w + x];
    __syncthreads();
    if (x0 + lx >= w) return;
    for (int r = ly0; r < VT; r += 8) {
        const int y = y0 + r;
        if (y >= h) break;
        int s = 0;
#pragma unroll 4
        for (int j = 7; j <= 93; ++j) s += (int)kGauss101[j] * (int)col[r + j][lx];
        dst[(size_t)y * w + x0 + lx] = (uint8_t)((s + 32768) >> 16);
    }
}

// ---- the blend: four individually rounded float operations per channel, saturate_cast<uchar> ------------------------------
__global__ void __launch_bounds__(256)
trail_blend_kernel(const uint8_t* __restrict__ warped, size_t wpitch, const uint8_t* __restrict__ alpha8, const uint8_t* __restrict__ bgg,
                   int w, int h, uint8_t* __restrict__ out, size_t out_pitch) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= w || y >= h) return;
    const float a = __fmul_rn((float)alpha8[(size_t)y * w + x], (float)(1.0 / 255.0));     // convertTo(CV_32F, 1/255)
    const float ia = __fsub_rn(1.f, a);
    const float b = (float)bgg[(size_t)y * w + x];
    const uint8_t* f = warped + (size_t)y * wpitch + (size_t)x * 3;
    uint8_t* o = out + (size_t)y * out_pitch + (size_t)x * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float v = __fadd_rn(__fmul_rn(a, (float)f[c]), __fmul_rn(ia, b));
        o[c] = (uint8_t)min(255, max(0, __float2int_rn(v)));
    }
}

// border colour 0 for the warp inside copyFeathered (cv::warpPerspective's default border value)
__global__ void trail_params_kernel(const WarpParams* __restrict__ in, WarpParams* __restrict__ out, int src_slot) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    WarpParams p = *in;
    p.border[0] = p.border[1] = p.border[2] = p.border[3] = 0;
    if (src_slot >= 0) p.src_slot = src_slot;
    *out = p;
}

}  // namespace

size_t trail_workspace_bytes(int w, int h, size_t frame_bytes) {
    const size_t px = ((size_t)w * h + 255) & ~(size_t)255;
    return frame_bytes + 256 + px /* bg' */ + px /* warped mask */ + 2 * px /* rows pass */ + px /* alpha */ +
           (((size_t)h * 5 * sizeof(int2) + 255) & ~(size_t)255) + 256 /* params */;
}

// fg: source frames (ring / array, slot from wp unless src_slot >= 0), bg: the trail background (pitch rows), out: blended frame.
void launch_trail(const uint8_t* frames, size_t pitch, size_t frame_stride, long slot_mod, const WarpParams* wp, int src_slot,
                  const uint8_t* bg, int w, int h, void* workspace, uint8_t* out, size_t out_pitch, cudaStream_t st) {
    const size_t px = ((size_t)w * h + 255) & ~(size_t)255;
    const size_t frame_bytes = ((pitch * (size_t)h) + 255) & ~(size_t)255;
    uint8_t* p = static_cast<uint8_t*>(workspace);
    uint8_t* warped = p; p += frame_bytes;
    uint8_t* bgg = p; p += px;
    uint8_t* mask = p; p += px;
    unsigned short* rows16 = reinterpret_cast<unsigned short*>(p); p += 2 * px;
    uint8_t* alpha = p; p += px;
    int2* rows = reinterpret_cast<int2*>(p); p += ((size_t)h * kSpans * sizeof(int2) + 255) & ~(size_t)255;
    WarpParams* wp0 = reinterpret_cast<WarpParams*>(p);
    count_launch(8);
    trail_params_kernel<<<1, 32, 0, st>>>(wp, wp0, src_slot);
    launch_warp(frames, pitch, frame_stride, slot_mod, wp0, 1, w, h, warped, pitch, 0, st);
    trail_bg_kernel<<<dim3((w + BT - 1) / BT, (h + BT - 1) / BT), 256, 0, st>>>(bg, pitch, w, h, bgg);
    trail_rows_reset_kernel<<<(h * kSpans + 255) / 256, 256, 0, st>>>(rows, w, h);
    trail_poly_kernel<<<1, 32, 0, st>>>(wp0, w, h, rows);
    trail_mask_kernel<<<dim3((w + 31) / 32, (h + 7) / 8), 256, 0, st>>>(wp0, rows, w, h, mask);
    trail_blur_rows_kernel<<<dim3((w + 255) / 256, h), 256, 0, st>>>(mask, w, h, rows16);
    trail_blur_cols_kernel<<<dim3((w + 31) / 32, (h + VT - 1) / VT), 256, 0, st>>>(rows16, w, h, alpha);
    trail_blend_kernel<<<dim3((w + 31) / 32, (h + 7) / 8), 256, 0, st>>>(warped, pitch, alpha, bgg, w, h, out, out_pitch);
}

}  // namespace vstabk
