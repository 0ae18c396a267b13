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
// dp2a with signed 16-bit weights and unsigned 8-bit pixels (w11 = 2^14 - w00 - w01 - w10 can be -1):
// d = c + a.s16[0] * b.u8[2*hi] + a.s16[1] * b.u8[2*hi+1]
VSTAB_D int dp2a_lo_su(unsigned a, unsigned b, int c) {
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
VSTAB_D int dp2a_hi_su(unsigned a, unsigned b, int c) {
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// keep a computed pointer in registers (ptxas otherwise rematerialises the 64-bit base arithmetic at every tap)
VSTAB_D const unsigned* pin(const unsigned* p) {
    asm volatile("" : "+l"(p));
    return p;
}
VSTAB_D unsigned pack_w(int lo, int hi) { return ((unsigned)lo & 0xffffu) | ((unsigned)hi << 16); }

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
        const unsigned* I = reinterpret_cast<const unsigned*>(pI + d.qoff[L]) + org;      // bilinear quads of the previous frame
        const unsigned* dI = reinterpret_cast<const unsigned*>(pI + d.doff[L]) + org;    // {dx, dy} short2
        const unsigned* J = reinterpret_cast<const unsigned*>(pJ + d.qoff[L]) + org;      // bilinear quads of the next frame
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
        const unsigned wt01 = pack_w(w00, w01), wt23 = pack_w(w10, w11);
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
            // window origin as pinned 64-bit pointers: one 32->64-bit multiply-add per tap, nothing rematerialised
            const unsigned* Ip = pin(I + (iy * P + ix));
            const unsigned* Dp = pin(dI + (iy * P + ix));
            const unsigned* DpP = pin(Dp + P);
            if (w00 == (1 << 14)) {
                // integer window origin (always at level 0: Shi-Tomasi corners are integer-valued): the bilinear taps
                // collapse to the pixel itself -- (16384 p + 256) >> 9 == 32 p, (16384 d + 8192) >> 14 == d
#pragma unroll
                for (int s = 0; s < kStrides; ++s) {
                    Iw[s] = 0; Ix[s] = 0; Iy[s] = 0;
                    if (lane + 32 * s < kPix) {
                        const unsigned qi = __ldg(Ip + koff[s]);
                        const unsigned d00 = __ldg(Dp + koff[s]);
                        Iw[s] = (int)(qi & 0xffu) << 5;
                        Ix[s] = (int)(short)(d00 & 0xffffu);
                        Iy[s] = (int)d00 >> 16;
                        sA11 += Ix[s] * Ix[s];
                        sA12 += Ix[s] * Iy[s];
                        sA22 += Iy[s] * Iy[s];
                    }
                }
            } else {
#pragma unroll
                for (int s = 0; s < kStrides; ++s) {
                    Iw[s] = 0; Ix[s] = 0; Iy[s] = 0;
                    if (lane + 32 * s < kPix) {
                        const unsigned* g = Dp + koff[s];
                        const unsigned* gP = DpP + koff[s];
                        // the 2x2 neighbourhood in one word: 16-bit weights x 8-bit pixels, two dp2a
                        const unsigned qi = __ldg(Ip + koff[s]);
                        const int iv = dp2a_lo_su(wt01, qi, dp2a_hi_su(wt23, qi, 1 << 8));
                        const unsigned d00 = __ldg(g), d01 = __ldg(g + 1), d10 = __ldg(gP), d11 = __ldg(gP + 1);
                        const int xv = (int)(short)(d00 & 0xffffu) * w00 + (int)(short)(d01 & 0xffffu) * w01 +
                                       (int)(short)(d10 & 0xffffu) * w10 + (int)(short)(d11 & 0xffffu) * w11;
                        const int yv = ((int)d00 >> 16) * w00 + ((int)d01 >> 16) * w01 + ((int)d10 >> 16) * w10 + ((int)d11 >> 16) * w11;
                        Iw[s] = iv >> 9;
                        Ix[s] = (xv + (1 << 13)) >> 14;
                        Iy[s] = (yv + (1 << 13)) >> 14;
                        sA11 += Ix[s] * Ix[s];
                        sA12 += Ix[s] * Iy[s];
                        sA22 += Iy[s] * Iy[s];
                    }
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
            unsigned vt01, vt23;
            {
                const float a = __fsub_rn(nx, (float)jx), b = __fsub_rn(ny, (float)jy);
                const float ia = __fsub_rn(1.f, a), ib = __fsub_rn(1.f, b);
                const int v00 = __float2int_rn(__fmul_rn(__fmul_rn(ia, ib), 16384.f));
                const int v01 = __float2int_rn(__fmul_rn(__fmul_rn(a, ib), 16384.f));
                const int v10 = __float2int_rn(__fmul_rn(__fmul_rn(ia, b), 16384.f));
                const int v11 = 16384 - v00 - v01 - v10;
                vt01 = pack_w(v00, v01); vt23 = pack_w(v10, v11);
            }
            int sb1 = 0, sb2 = 0;                   // per lane <= 14 * 8160 * 4080 < 2^31
            const unsigned* Jp = pin(J + (jy * P + jx));
#pragma unroll
            for (int s = 0; s < kStrides; ++s) {
                if (lane + 32 * s < kPix) {
                    const unsigned q = __ldg(Jp + koff[s]);
                    const int jv = dp2a_lo_su(vt01, q, dp2a_hi_su(vt23, q, 1 << 8)) >> 9;
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
