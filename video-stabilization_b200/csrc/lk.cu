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
// The 21x21 template (I, Ix, Iy) of a level lives in registers: the window is cut into 63 vertical
// strips of 7 pixels (3 strip rows x 21 columns) and lane l owns strips l and l + 32 (lane 31 has one).
// Every level of a frame's pyramid block has the same row pitch, so with the pitch as a template
// argument the 7 taps of a strip are ONE pointer + immediate offsets: two address computations per
// iteration instead of fourteen, no per-tap offset registers, all 14 loads of an iteration issued
// back to back.  A request touches two image rows (lanes 0..20 one row, lanes 21..31 the row seven
// below).  Intensities come from the padded u8 level (reflect-101 border) and derivatives from the
// padded Scharr level (zero border) that K2 prepares, so no tap ever needs border logic; both are
// tiny and L1/L2-resident.  The 2x2 system and the mismatch vector are reduced exactly with
// redux.sync.  (Round-1 history: pixel-major mapping, lane owns pixels l + 32 s with 14 offset
// register pairs and a divergent last slot: 2.62 ms per 256 frames; this mapping: 1.84 ms.)
#include <cstdlib>
#include "kernels.h"

namespace vstabk {
namespace {

constexpr int kPix = kLkWin * kLkWin;  // 441

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

// (float)(exact 64-bit warp sum of v) * 2^-20 without 64-bit integers: the sum is hi * 2^16 + lo with the two redux.sync
// results lo < 2^21 and |hi| <= 2^20, which convert exactly; the products by powers of two are exact and the fma rounds
// the exact sum once -- the same float as an int64 -> float conversion followed by the scaling.
VSTAB_D float warp_sum_scaled(int v) {
    const int lo = __reduce_add_sync(0xffffffffu, v & 0xffff);
    const int hi = __reduce_add_sync(0xffffffffu, v >> 16);
    return __fmaf_rn((float)hi, 0.0625f, __fmul_rn((float)lo, 1.f / (float)(1 << 20)));
}

// kP = 0: run-time pitch (any working width); kP > 0: the tap offsets are immediates.
constexpr int kStripLen = 7, kStrips = 63;

template <int kWarpsPerBlock, int kMinBlocks, int kP>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, kMinBlocks)
lk_strip_kernel(const uint8_t* __restrict__ prev_pyr, const uint8_t* __restrict__ next_pyr,
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

    // the two strips of this lane: strip i covers rows 7 (i / 21) .. + 6 of column i % 21
    const bool v1 = lane + 32 < kStrips;                       // lane 31: its second strip aliases strip 62 with zero derivatives
    const int i1 = v1 ? lane + 32 : kStrips - 1;
    const int q0 = (lane * 3121) >> 16, q1 = (i1 * 3121) >> 16;   // i / 21 for i < 448
    const int c0 = lane - kLkWin * q0, c1 = i1 - kLkWin * q1;

    float outx = 0.f, outy = 0.f;
    int st = 1;
    const float half = (float)(kLkWin - 1) * 0.5f;

#pragma unroll 1
    for (int L = d.nlev - 1; L >= 0; --L) {
        const int cols = d.w[L], rows = d.h[L];
        const int P = kP > 0 ? kP : d.pitch[L];
        const size_t org = (size_t)kLkPad * P + kLkPad;          // padded offset of image pixel (0,0)
        const unsigned* I = reinterpret_cast<const unsigned*>(pI + d.qoff[L]) + org;      // bilinear quads of the previous frame
        const unsigned* dI = reinterpret_cast<const unsigned*>(pI + d.doff[L]) + org;    // {dx, dy} short2
        const unsigned* J = reinterpret_cast<const unsigned*>(pJ + d.qoff[L]) + org;      // bilinear quads of the next frame
        const int so0 = kStripLen * q0 * P + c0, so1 = kStripLen * q1 * P + c1;
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

        // ---- template patch (I in Q5, Ix/Iy) into registers, 2x2 normal matrix ---------------
        int Iw[2 * kStripLen], Ix[2 * kStripLen], Iy[2 * kStripLen];
        int sA11 = 0, sA12 = 0, sA22 = 0;          // per lane <= 14 * 4080^2 < 2^31
        {
            const int wo = iy * P + ix;
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const int so = b ? so1 : so0;
                const unsigned* Ip = pin(I + (wo + so));
                const unsigned* Dp = pin(dI + (wo + so));
                if (w00 == (1 << 14)) {
                    // integer window origin (always at level 0: Shi-Tomasi corners are integer-valued): the bilinear taps
                    // collapse to the pixel itself -- (16384 p + 256) >> 9 == 32 p, (16384 d + 8192) >> 14 == d
#pragma unroll
                    for (int t = 0; t < kStripLen; ++t) {
                        const int s = b * kStripLen + t;
                        const unsigned qi = __ldg(Ip + t * P);
                        const unsigned d00 = __ldg(Dp + t * P);
                        Iw[s] = (int)(qi & 0xffu) << 5;
                        Ix[s] = (int)(short)(d00 & 0xffffu);
                        Iy[s] = (int)d00 >> 16;
                    }
                } else {
                    // derivative rows 0..7 of the strip, two columns: the lower taps of row t are the upper taps of row t + 1
                    unsigned dl[kStripLen + 1], dr[kStripLen + 1];
#pragma unroll
                    for (int t = 0; t <= kStripLen; ++t) { dl[t] = __ldg(Dp + t * P); dr[t] = __ldg(Dp + t * P + 1); }
#pragma unroll
                    for (int t = 0; t < kStripLen; ++t) {
                        const int s = b * kStripLen + t;
                        // the 2x2 neighbourhood in one word: 16-bit weights x 8-bit pixels, two dp2a
                        const unsigned qi = __ldg(Ip + t * P);
                        const int iv = dp2a_lo_su(wt01, qi, dp2a_hi_su(wt23, qi, 1 << 8));
                        const unsigned d00 = dl[t], d01 = dr[t], d10 = dl[t + 1], d11 = dr[t + 1];
                        const int xv = (int)(short)(d00 & 0xffffu) * w00 + (int)(short)(d01 & 0xffffu) * w01 +
                                       (int)(short)(d10 & 0xffffu) * w10 + (int)(short)(d11 & 0xffffu) * w11;
                        const int yv = ((int)d00 >> 16) * w00 + ((int)d01 >> 16) * w01 + ((int)d10 >> 16) * w10 + ((int)d11 >> 16) * w11;
                        Iw[s] = iv >> 9;
                        Ix[s] = (xv + (1 << 13)) >> 14;
                        Iy[s] = (yv + (1 << 13)) >> 14;
                    }
                }
            }
            if (!v1) {
#pragma unroll
                for (int t = 0; t < kStripLen; ++t) { Ix[kStripLen + t] = 0; Iy[kStripLen + t] = 0; }
            }
#pragma unroll
            for (int s = 0; s < 2 * kStripLen; ++s) {
                sA11 += Ix[s] * Ix[s];
                sA12 += Ix[s] * Iy[s];
                sA22 += Iy[s] * Iy[s];
            }
        }
        const float A11 = warp_sum_scaled(sA11);
        const float A12 = warp_sum_scaled(sA12);
        const float A22 = warp_sum_scaled(sA22);
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
            const int jo = jy * P + jx;
            const unsigned* Jp0 = pin(J + (jo + so0));
            const unsigned* Jp1 = pin(J + (jo + so1));
            unsigned q[2 * kStripLen];
#pragma unroll
            for (int t = 0; t < kStripLen; ++t) { q[t] = __ldg(Jp0 + t * P); q[kStripLen + t] = __ldg(Jp1 + t * P); }
#pragma unroll
            for (int s = 0; s < 2 * kStripLen; ++s) {
                const int jv = dp2a_lo_su(vt01, q[s], dp2a_hi_su(vt23, q[s], 1 << 8)) >> 9;
                const int diff = jv - Iw[s];
                sb1 += diff * Ix[s];
                sb2 += diff * Iy[s];
            }
            const float b1 = warp_sum_scaled(sb1);
            const float b2 = warp_sum_scaled(sb2);
            const float ddx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), D);
            const float ddy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), D);
            nx = __fadd_rn(nx, ddx); ny = __fadd_rn(ny, ddy);
            outx = __fadd_rn(nx, half); outy = __fadd_rn(ny, half);
            // delta.ddot(delta) <= criteria.epsilon^2 with Point2f::ddot in double: decided in float when the float value
            // is outside a 1e-5 relative band around the threshold (float error <= 2^-22), in double inside it
            {
                const float sf = __fmaf_rn(ddx, ddx, __fmul_rn(ddy, ddy));
                bool conv = sf < 0.99999e-4f;
                if (!conv && sf < 1.00001e-4f) conv = (double)ddx * (double)ddx + (double)ddy * (double)ddy <= 0.01 * 0.01;
                if (conv) break;
            }
            // float < 0.01 (a double): 0x3C23D70A = 0.0099999998 is the largest float below it
            if (j > 0 && fabsf(__fadd_rn(ddx, pdx)) <= 0.01f && fabsf(__fadd_rn(ddy, pdy)) <= 0.01f) {
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

template <int kWarpsPerBlock, int kMinBlocks>
void launch_strip(int P, int nframes, cudaStream_t st, const uint8_t* prev_pyr, const uint8_t* next_pyr, size_t prev_stride,
                  size_t next_stride, const PyrDesc& d, const float2* pts, const int* counts, float2* out_pts, uint8_t* status) {
    const int T = kWarpsPerBlock * 32;
    const dim3 grid((kMaxCorners + kWarpsPerBlock - 1) / kWarpsPerBlock, nframes);
    // working widths 640 (wh 360 of 16:9 input), 1920, 3840 (+ 2 * kLkPad) get immediate tap offsets
    if (P == 688) lk_strip_kernel<kWarpsPerBlock, kMinBlocks, 688><<<grid, T, 0, st>>>(prev_pyr, next_pyr, prev_stride, next_stride, d, pts, counts, out_pts, status);
    else if (P == 1968) lk_strip_kernel<kWarpsPerBlock, kMinBlocks, 1968><<<grid, T, 0, st>>>(prev_pyr, next_pyr, prev_stride, next_stride, d, pts, counts, out_pts, status);
    else if (P == 3888) lk_strip_kernel<kWarpsPerBlock, kMinBlocks, 3888><<<grid, T, 0, st>>>(prev_pyr, next_pyr, prev_stride, next_stride, d, pts, counts, out_pts, status);
    else lk_strip_kernel<kWarpsPerBlock, kMinBlocks, 0><<<grid, T, 0, st>>>(prev_pyr, next_pyr, prev_stride, next_stride, d, pts, counts, out_pts, status);
}

}  // namespace

void launch_lk(const uint8_t* prev_pyr, const uint8_t* next_pyr, size_t prev_stride, size_t next_stride,
               const PyrDesc& d, const float2* pts, const int* counts, int nframes,
               float2* out_pts, uint8_t* status, cudaStream_t st) {
    if (nframes <= 0) return;
    count_launch(1);
    // VSTAB_LK_WARPS: features (warps) per CTA -- a CTA's registers are held until its slowest feature converges, so
    // one-warp CTAs refill soonest; VSTAB_LK_REGS: 96 or 80 registers per thread.  Measured on B200, ms per 256 frames
    // (warps per CTA / registers): 4/96 1.73, 4/80 1.91, 2/96 1.71, 2/80 1.88, 1/96 1.66, 1/80 1.62 (default).
    // An L1 prefetch (CCTL.PF1) of the next level's windows cost +0.1 ms and was not kept.
    static int wpb = 0, regs = 80;
    if (wpb == 0) {
        const char* e = getenv("VSTAB_LK_WARPS"); wpb = e ? atoi(e) : 1;
        if (const char* r = getenv("VSTAB_LK_REGS")) regs = atoi(r);
    }
    bool uniform = true;
    for (int l = 1; l < d.nlev; ++l) uniform = uniform && d.pitch[l] == d.pitch[0];
    const int P = uniform ? d.pitch[0] : 0;
#define VSTAB_LK_ARGS P, nframes, st, prev_pyr, next_pyr, prev_stride, next_stride, d, pts, counts, out_pts, status
    // min blocks per SM = 65536 / (registers * threads per CTA)
    if (wpb == 1) {
        if (regs == 64) launch_strip<1, 32>(VSTAB_LK_ARGS); else if (regs == 72) launch_strip<1, 28>(VSTAB_LK_ARGS);
        else if (regs == 80) launch_strip<1, 24>(VSTAB_LK_ARGS); else launch_strip<1, 20>(VSTAB_LK_ARGS);
    }
    else if (wpb == 2) { if (regs == 80) launch_strip<2, 12>(VSTAB_LK_ARGS); else launch_strip<2, 10>(VSTAB_LK_ARGS); }
    else { if (regs == 80) launch_strip<4, 6>(VSTAB_LK_ARGS); else launch_strip<4, 5>(VSTAB_LK_ARGS); }
#undef VSTAB_LK_ARGS
}

}  // namespace vstabk
