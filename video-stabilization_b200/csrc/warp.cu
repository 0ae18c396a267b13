// K7: fused output warp
//   cv::warpPerspective(presentation_image, warped, H_stabilize_scaled, frame.size(),
//                       INTER_LINEAR, BORDER_CONSTANT, 0.5*mean)      /root/reference/src/stabilizer.cpp:1309-1313
// Bit-exact restatement of OpenCV's fixed-point path: inverse map in f64,
// (X,Y)*32/W rounded half-even to Q5, 2x2 taps with Q15 weights
//   w = {(32-ay)(32-ax), (32-ay)ax, ay(32-ax), ay ax} * 32     (== OpenCV's int16 table; the one
//   saturated entry (0,0) -> {32767,1,0,0} yields the same pixel),
// out = (sum + 16384) >> 15, taps outside the source take the constant border colour.
// HBM-bound: reads 3WH, writes 3WH per frame.  Each thread produces 4 consecutive pixels
// (three 32-bit stores); source taps are fetched as aligned 32-bit words and funnel-shifted.
#include "kernels.h"

namespace vstabk {
namespace {

constexpr int WTX = 64, WTY = 4;   // threads; each thread -> 4 px  => 256 x 4 px per CTA

struct Taps6 { unsigned lo, hi; };  // bytes A..A+7 (6 used: two BGR pixels)

VSTAB_D Taps6 load6(const uint8_t* row, int byte_off) {
    const int al = byte_off & ~3;
    const int sh = (byte_off & 3) * 8;
    const unsigned* p = reinterpret_cast<const unsigned*>(row + al);
    const unsigned w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
    Taps6 t;
    t.lo = __funnelshift_r(w0, w1, sh);
    t.hi = __funnelshift_r(w1, w2, sh);
    return t;
}

VSTAB_D int byte_of(const Taps6& t, int i) {   // i in 0..5
    return i < 4 ? (int)((t.lo >> (8 * i)) & 0xffu) : (int)((t.hi >> (8 * (i - 4))) & 0xffu);
}

__global__ void __launch_bounds__(WTX * WTY)
warp_kernel(const uint8_t* __restrict__ frames, size_t pitch, size_t frame_stride, long slot_mod,
            const WarpParams* __restrict__ wps, int w, int h,
            uint8_t* __restrict__ out, size_t out_pitch, size_t out_frame_stride) {
    const int oi = blockIdx.z;
    __shared__ double sM[9];
    __shared__ int sB[3];
    __shared__ int sSlot;
    if (threadIdx.y == 0 && threadIdx.x < 9) sM[threadIdx.x] = wps[oi].Minv[threadIdx.x];
    if (threadIdx.y == 0 && threadIdx.x >= 16 && threadIdx.x < 19) sB[threadIdx.x - 16] = wps[oi].border[threadIdx.x - 16];
    if (threadIdx.y == 0 && threadIdx.x == 32) sSlot = wps[oi].src_slot;
    __syncthreads();
    const uint8_t* src = frames + (size_t)(slot_mod > 0 ? (sSlot % slot_mod) : sSlot) * frame_stride;
    uint8_t* dst = out + (size_t)oi * out_frame_stride;

    const int y = blockIdx.y * WTY + threadIdx.y;
    const int x0 = (blockIdx.x * WTX + threadIdx.x) * 4;
    if (y >= h || x0 >= w) return;

    const double M0 = sM[0], M1 = sM[1], M2 = sM[2], M3 = sM[3], M4 = sM[4], M5 = sM[5],
                 M6 = sM[6], M7 = sM[7], M8 = sM[8];
    const double yd = (double)y;
    const bool al_ok = ((pitch & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 3) == 0);
    unsigned char res[12];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int x = x0 + i;
        // OpenCV evaluates per 32-px block: X0 = M0*bx + M1*y + M2, then X0 + M0*x1
        const double bx = (double)(x & ~31), x1 = (double)(x & 31);
        const double X = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(M0, bx), __dmul_rn(M1, yd)), M2), __dmul_rn(M0, x1));
        const double Y = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(M3, bx), __dmul_rn(M4, yd)), M5), __dmul_rn(M3, x1));
        double W = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(M6, bx), __dmul_rn(M7, yd)), M8), __dmul_rn(M6, x1));
        W = W != 0.0 ? __ddiv_rn(32.0, W) : 0.0;
        double fX = __dmul_rn(X, W), fY = __dmul_rn(Y, W);
        fX = fmax(-2147483648.0, fmin(2147483647.0, fX));
        fY = fmax(-2147483648.0, fmin(2147483647.0, fY));
        const int iX = __double2int_rn(fX), iY = __double2int_rn(fY);
        int sx = iX >> 5, sy = iY >> 5;
        const int ax = iX & 31, ay = iY & 31;
        sx = max(-32768, min(32767, sx));      // remap stores int16 coordinates
        sy = max(-32768, min(32767, sy));
        const int w00 = (32 - ay) * (32 - ax) * 32, w01 = (32 - ay) * ax * 32;
        const int w10 = ay * (32 - ax) * 32, w11 = ay * ax * 32;
        int p00[3], p01[3], p10[3], p11[3];
        const bool fast = al_ok && sx >= 0 && sy >= 0 && sy + 1 < h && 3 * sx + 14 <= 3 * w;
        if (fast) {
            const uint8_t* r0 = src + (size_t)sy * pitch;
            const Taps6 a = load6(r0, 3 * sx);
            const Taps6 b = load6(r0 + pitch, 3 * sx);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                p00[c] = byte_of(a, c); p01[c] = byte_of(a, 3 + c);
                p10[c] = byte_of(b, c); p11[c] = byte_of(b, 3 + c);
            }
        } else {
            const bool y0in = sy >= 0 && sy < h, y1in = sy + 1 >= 0 && sy + 1 < h;
            const bool x0in = sx >= 0 && sx < w, x1in = sx + 1 >= 0 && sx + 1 < w;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                p00[c] = (y0in && x0in) ? (int)__ldg(src + (size_t)sy * pitch + 3 * sx + c) : sB[c];
                p01[c] = (y0in && x1in) ? (int)__ldg(src + (size_t)sy * pitch + 3 * (sx + 1) + c) : sB[c];
                p10[c] = (y1in && x0in) ? (int)__ldg(src + (size_t)(sy + 1) * pitch + 3 * sx + c) : sB[c];
                p11[c] = (y1in && x1in) ? (int)__ldg(src + (size_t)(sy + 1) * pitch + 3 * (sx + 1) + c) : sB[c];
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int v = (p00[c] * w00 + p01[c] * w01 + p10[c] * w10 + p11[c] * w11 + 16384) >> 15;
            res[i * 3 + c] = (unsigned char)min(255, max(0, v));
        }
    }
    uint8_t* o = dst + (size_t)y * out_pitch + (size_t)x0 * 3;
    if (x0 + 4 <= w && ((out_pitch & 3) == 0)) {
        unsigned* o32 = reinterpret_cast<unsigned*>(o);
        o32[0] = res[0] | (res[1] << 8) | (res[2] << 16) | ((unsigned)res[3] << 24);
        o32[1] = res[4] | (res[5] << 8) | (res[6] << 16) | ((unsigned)res[7] << 24);
        o32[2] = res[8] | (res[9] << 8) | (res[10] << 16) | ((unsigned)res[11] << 24);
    } else {
        const int n = min(4, w - x0) * 3;
        for (int k = 0; k < n; ++k) o[k] = res[k];
    }
}

}  // namespace

void launch_warp(const uint8_t* frames, size_t pitch, size_t frame_stride, long slot_mod,
                 const WarpParams* wp, int nout, int w, int h,
                 uint8_t* out, size_t out_pitch, size_t out_frame_stride, cudaStream_t st) {
    if (nout <= 0) return;
    dim3 block(WTX, WTY);
    dim3 grid((w + WTX * 4 - 1) / (WTX * 4), (h + WTY - 1) / WTY, nout);
    count_launch(1);
    warp_kernel<<<grid, block, 0, st>>>(frames, pitch, frame_stride, slot_mod, wp, w, h, out, out_pitch,
                                        out_frame_stride);
}

}  // namespace vstabk
