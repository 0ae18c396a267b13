// K4: sparse pyramidal Lucas-Kanade, one warp per feature
//   == cv::calcOpticalFlowPyrLK(prevGray, gray, prevPts, ..., Size(21,21), 3,
//                               {COUNT+EPS, 50, 0.01}, 0, 1e-4)
// (/root/reference/src/stabilizer.cpp:170-209).  Control flow, fixed-point formats and
// float operation order follow OpenCV's LKTrackerInvoker: Q14 bilinear
// weights, Q5 intensity patches, Scharr derivatives (zero outside the image, intensity
// reflect-101 padded), float 2x2 normal equations, eps^2 = 1e-4, oscillation damping.
// Differences from OpenCV are confined to the summation order of the 441-term sums
// (exact 64-bit integer sums here) => <= 1e-3 px against the 0.05 px tolerance.
//
// The 21x21 template (I, Ix, Iy) of a level lives in registers: lane l owns pixels
// k = l + 32 s (s = 0..13).  Intensities come from the padded u8 level (reflect-101 border)
// and derivatives from the padded Scharr level (zero border) that K2 prepares, so no tap ever
// needs border logic; both are tiny and L1/L2-resident.  The 2x2 system and the mismatch
// vector are reduced exactly with redux.sync.
#include <cstdlib>
#include "kernels.h"

namespace vstabk {
namespace {

constexpr int kWarpsPerBlock = 4;
constexpr int kPix = kLkWin * kLkWin;  // 441
constexpr int kStrides = (kPix + 31) / 32;  // 14

// exact warp-wide sum of one int32 per lane (|v| < 2^31) as a 64-bit integer: two redux.sync
VSTAB_D long long warp_sum_i32(int v) {
    const int lo = __reduce_add_sync(0xffffffffu, v & 0xffff);
    const int hi = __reduce_add_sync(0xffffffffu, v >> 16);
    return ((long long)hi << 16) + (long long)lo;
}

template <int kMinBlocks>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, kMinBlocks)
lk_kernel(const uint8_t* __restrict__ prev_pyr, const uint8_t* __restrict__ next_pyr,
          size_t prev_stride, size_t next_stride, PyrDesc d,
          const float2* __restrict__ pts, const int* __restrict__ counts,
          float2* __restrict__ out_pts, uint8_t* __restrict__ status) {
    const int frame = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int fi = blockIdx.x * kWarpsPerBlock + warp;
    if (fi >= counts[frame]) return;
    const uint8_t* pI = prev_pyr + (size_t)frame * prev_stride;
    const uint8_t* pJ = next_pyr + (size_t)frame * next_stride;
    const float2 pt = pts[(size_t)frame * kMaxCorners + fi];

    float outx = 0.f, outy = 0.f;
    int st = 1;
    const float half = (float)(kLkWin - 1) * 0.5f;
    const float kFltScale = 1.f / (float)(1 << 20);

#pragma unroll 1
    for (int L = d.nlev - 1; L >= 0; --L) {
        const int cols = d.w[L], rows = d.h[L], P = d.pitch[L];
        const size_t org = (size_t)kLkPad * P + kLkPad;          // padded offset of image pixel (0,0)
        const uint8_t* I = pI + d.poff[L] + org;
        const short* dI = reinterpret_cast<const short*>(pI + d.doff[L]) + org * 2;
        const uint8_t* J = pJ + d.poff[L] + org;
        const float lscale = (float)(1. / (1 << L));
        float px = __fmul_rn(pt.x, lscale), py = __fmul_rn(pt.y, lscale);
        float nx, ny;
        if (L == d.nlev - 1) { nx = px; ny = py; }
        else { nx = __fmul_rn(outx, 2.f); ny = __fmul_rn(outy, 2.f); }
        outx = nx; outy = ny;
        px = __fsub_rn(px, half); py = __fsub_rn(py, half);
        const int ix = (int)floorf(px), iy = (int)floorf(py);
        if (ix < -kLkWin || ix >= cols || iy < -kLkWin || iy >= rows) {
            if (L == 0) st = 0;
            continue;
        }
        int w00, w01, w10, w11;
        {
            const float a = __fsub_rn(px, (float)ix), b = __fsub_rn(py, (float)iy);
            const float ia = __fsub_rn(1.f, a), ib = __fsub_rn(1.f, b);
            w00 = __float2int_rn(__fmul_rn(__fmul_rn(ia, ib), 16384.f));
            w01 = __float2int_rn(__fmul_rn(__fmul_rn(a, ib), 16384.f));
            w10 = __float2int_rn(__fmul_rn(__fmul_rn(ia, b), 16384.f));
            w11 = 16384 - w00 - w01 - w10;
        }
        // offsets of the pixels this lane owns (k = lane + 32 s -> ky * P + kx), shared by the
        // template and every iteration of this level
        int koff[kStrides];
#pragma unroll
        for (int s = 0; s < kStrides; ++s) {
            const int k = lane + 32 * s;
            const int ky = (k * 3121) >> 16;               // k / 21 for k < 448
            koff[s] = ky * P + (k - ky * kLkWin);
        }

        // ---- template patch (I in Q5, Ix/Iy) into registers, 2x2 normal matrix ---------------
        int Iw[kStrides], Ix[kStrides], Iy[kStrides];
        int sA11 = 0, sA12 = 0, sA22 = 0;          // per lane <= 14 * 4080^2 < 2^31
        {
            const int o0 = iy * P + ix;
#pragma unroll
            for (int s = 0; s < kStrides; ++s) {
                Iw[s] = 0; Ix[s] = 0; Iy[s] = 0;
                if (lane + 32 * s < kPix) {
                    const int o = o0 + koff[s];
                    const uint8_t* q = I + o;
                    const short* g = dI + 2 * o;
                    const int iv = q[0] * w00 + q[1] * w01 + q[P] * w10 + q[P + 1] * w11;
                    const int xv = g[0] * w00 + g[2] * w01 + g[2 * P] * w10 + g[2 * P + 2] * w11;
                    const int yv = g[1] * w00 + g[3] * w01 + g[2 * P + 1] * w10 + g[2 * P + 3] * w11;
                    Iw[s] = (iv + (1 << 8)) >> 9;
                    Ix[s] = (xv + (1 << 13)) >> 14;
                    Iy[s] = (yv + (1 << 13)) >> 14;
                    sA11 += Ix[s] * Ix[s];
                    sA12 += Ix[s] * Iy[s];
                    sA22 += Iy[s] * Iy[s];
                }
            }
        }
        const float A11 = __fmul_rn((float)warp_sum_i32(sA11), kFltScale);
        const float A12 = __fmul_rn((float)warp_sum_i32(sA12), kFltScale);
        const float A22 = __fmul_rn((float)warp_sum_i32(sA22), kFltScale);
        float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
        {
            const float dd = __fsub_rn(A11, A22);
            const float rad = __fsqrt_rn(__fadd_rn(__fmul_rn(dd, dd), __fmul_rn(__fmul_rn(4.f, A12), A12)));
            const float min_eig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), rad), (float)(2 * kPix));
            if ((double)min_eig < 1e-4 || D < 1.1920928955078125e-7f) {   // minEigThreshold is a double
                if (L == 0) st = 0;
                continue;
            }
        }
        D = __fdiv_rn(1.f, D);
        nx = __fsub_rn(nx, half); ny = __fsub_rn(ny, half);
        float pdx = 0.f, pdy = 0.f;
#pragma unroll 1
        for (int j = 0; j < kLkMaxIter; ++j) {
            const int jx = (int)floorf(nx), jy = (int)floorf(ny);
            if (jx < -kLkWin || jx >= cols || jy < -kLkWin || jy >= rows) {
                if (L == 0) st = 0;
                break;
            }
            int v00, v01, v10, v11;
            {
                const float a = __fsub_rn(nx, (float)jx), b = __fsub_rn(ny, (float)jy);
                const float ia = __fsub_rn(1.f, a), ib = __fsub_rn(1.f, b);
                v00 = __float2int_rn(__fmul_rn(__fmul_rn(ia, ib), 16384.f));
                v01 = __float2int_rn(__fmul_rn(__fmul_rn(a, ib), 16384.f));
                v10 = __float2int_rn(__fmul_rn(__fmul_rn(ia, b), 16384.f));
                v11 = 16384 - v00 - v01 - v10;
            }
            int sb1 = 0, sb2 = 0;                   // per lane <= 14 * 8160 * 4080 < 2^31
            const int o0 = jy * P + jx;
#pragma unroll
            for (int s = 0; s < kStrides; ++s) {
                if (lane + 32 * s < kPix) {
                    const uint8_t* q = J + (o0 + koff[s]);
                    const int jv = (q[0] * v00 + q[1] * v01 + q[P] * v10 + q[P + 1] * v11 + (1 << 8)) >> 9;
                    const int diff = jv - Iw[s];
                    sb1 += diff * Ix[s];
                    sb2 += diff * Iy[s];
                }
            }
            const float b1 = __fmul_rn((float)warp_sum_i32(sb1), kFltScale);
            const float b2 = __fmul_rn((float)warp_sum_i32(sb2), kFltScale);
            const float ddx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), D);
            const float ddy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), D);
            nx = __fadd_rn(nx, ddx); ny = __fadd_rn(ny, ddy);
            outx = __fadd_rn(nx, half); outy = __fadd_rn(ny, half);
            if ((double)ddx * (double)ddx + (double)ddy * (double)ddy <= 0.01 * 0.01) break;   // criteria.epsilon^2, Point2f::ddot in double
            if (j > 0 && (double)fabsf(__fadd_rn(ddx, pdx)) < 0.01 && (double)fabsf(__fadd_rn(ddy, pdy)) < 0.01) {
                outx = __fsub_rn(outx, __fmul_rn(ddx, 0.5f));
                outy = __fsub_rn(outy, __fmul_rn(ddy, 0.5f));
                break;
            }
            pdx = ddx; pdy = ddy;
        }
        // the `err` block of OpenCV's LKTrackerInvoker (the reference passes an err vector,
        // stabilizer.cpp:192-195): at level 0 a still-valid point whose final window origin lies
        // outside the image loses its status (coordinates are kept).
        if (L == 0 && st) {
            const int fx = (int)floorf(__fsub_rn(outx, half)), fy = (int)floorf(__fsub_rn(outy, half));
            if (fx < -kLkWin || fx >= cols || fy < -kLkWin || fy >= rows) st = 0;
        }
    }
    if (lane == 0) {
        out_pts[(size_t)frame * kMaxCorners + fi] = make_float2(outx, outy);
        status[(size_t)frame * kMaxCorners + fi] = (uint8_t)st;
    }
}

}  // namespace

void launch_lk(const uint8_t* prev_pyr, const uint8_t* next_pyr, size_t prev_stride, size_t next_stride,
               const PyrDesc& d, const float2* pts, const int* counts, int nframes,
               float2* out_pts, uint8_t* status, cudaStream_t st) {
    if (nframes <= 0) return;
    dim3 grid((kMaxCorners + kWarpsPerBlock - 1) / kWarpsPerBlock, nframes);
    count_launch(1);
    // resident CTAs per SM the register allocation is tuned for (4: 128 regs, 5: 96, 6: 80 + small spills);
    // measured on B200: 4 -> 8.2 ms, 5 -> 10.2 ms, 6 -> 8.8 ms per 512 frames, so 4 is the default
    static int mb = 0;
    if (mb == 0) { const char* e = getenv("VSTAB_LK_MINBLOCKS"); mb = e ? atoi(e) : 4; }
    if (mb == 4)
        lk_kernel<4><<<grid, kWarpsPerBlock * 32, 0, st>>>(prev_pyr, next_pyr, prev_stride, next_stride, d, pts, counts, out_pts, status);
    else if (mb == 6)
        lk_kernel<6><<<grid, kWarpsPerBlock * 32, 0, st>>>(prev_pyr, next_pyr, prev_stride, next_stride, d, pts, counts, out_pts, status);
    else
        lk_kernel<5><<<grid, kWarpsPerBlock * 32, 0, st>>>(prev_pyr, next_pyr, prev_stride, next_stride, d, pts, counts, out_pts, status);
}

}  // namespace vstabk
