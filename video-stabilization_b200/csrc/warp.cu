// K7: fused output warp
//   cv::warpPerspective(presentation_image, warped, H_stabilize_scaled, frame.size(),
//                       INTER_LINEAR, BORDER_CONSTANT, 0.5*mean)      /root/reference/src/stabilizer.cpp:1309-1313
// Bit-exact restatement of OpenCV's fixed-point path: inverse map in f64,
// (X,Y)*32/W rounded half-even to Q5, 2x2 taps with Q15 weights
//   w = {(32-ay)(32-ax), (32-ay)ax, ay(32-ax), ay ax} * 32     (== OpenCV's int16 table; the one
//   saturated entry (0,0) -> {32767,1,0,0} yields the same pixel),
// out = (sum + 16384) >> 15, taps outside the source take the constant border colour.
// HBM-bound: reads 3WH, writes 3WH per frame.
//
// One CTA produces a 128 x 32 destination tile.  Every H the stabilizer produces is affine
// (SURVEY B.8), so the source footprint of a tile is a small parallelogram: its bounding box
// (clamped to the image) is staged in shared memory with 16-byte loads, and the taps are read
// from there as aligned 32-bit words + funnel shifts.  Because the Q15 weights are exact
// products, the bilinear sum factors into a horizontal and a vertical blend
//   h_r = p_r0 (32-ax) + p_r1 ax,   out = (h_0 (32-ay) + h_1 ay + 512) >> 10
// which is the same integer; B and R ride in one register (16-bit fields) for the horizontal
// blend.  Pixels whose taps leave the staged box (image border, non-affine H, large rotations)
// take the generic per-tap path that reads global memory and substitutes the border colour.
#include <cstdlib>
#include "kernels.h"

namespace vstabk {
namespace {

constexpr int TW = 128, TH = 32;               // destination tile
constexpr int NTX = 32, NTY = 8;               // threads: 4 px per thread in x, TH/NTY rows per thread
constexpr int kStageBytes = 28 * 1024;         // staged source box per pipeline stage

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

// Generic pixel: taps from global memory, border substitution (OpenCV remap semantics incl. the
// int16 saturation of the integer coordinates).  Returns packed B | G<<8 | R<<16.
VSTAB_D unsigned generic_pixel(const uint8_t* __restrict__ src, size_t pitch, int w, int h, bool al_ok,
                               int iX, int iY, const int* sB) {
    int sx = iX >> 5, sy = iY >> 5;
    // all four taps outside the source: the weights sum to 2^15, so the pixel is the border colour
    if (sx < -1 || sx >= w || sy < -1 || sy >= h) return (unsigned)sB[0] | ((unsigned)sB[1] << 8) | ((unsigned)sB[2] << 16);
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
    unsigned res = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int v = (p00[c] * w00 + p01[c] * w01 + p10[c] * w10 + p11[c] * w11 + 16384) >> 15;
        res |= (unsigned)min(255, max(0, v)) << (8 * c);
    }
    return res;
}

// 6 bytes (two BGR pixels) at byte offset `o` of the staged box -> lo = B0 G0 R0 B1, hi = G1 R1 . .
VSTAB_D void staged_row(const uint8_t* __restrict__ sm, int o, unsigned& lo, unsigned& hi) {
    const unsigned* p = reinterpret_cast<const unsigned*>(sm + (o & ~3));
    const unsigned w0 = p[0], w1 = p[1], w2 = p[2];
    lo = __funnelshift_r(w0, w1, o << 3);                  // the shifter uses (o << 3) & 31 = 8 * (o & 3)
    hi = __funnelshift_r(w1, w2, o << 3);
}

struct TileBox { int fx0, fxn, fy0, fyn, b0, SP; };       // staged box: x in [fx0, fx0+fxn], y in [fy0, fy0+fyn]

// Everything the consumer warps need to know about one destination tile (written by the producer
// lane before it arrives on the tile's "full" barrier).
struct TileInfo {
    double M[9];
    const uint8_t* src;
    TileBox box;
    int border[3];
    int oi, tx0, ty0;
};

constexpr int kStages = 2;
constexpr int kConsumerWarps = NTY;                        // 8 warps x 32 lanes = the 32 x 8 thread tile
constexpr int kThreads = 32 * (1 + kConsumerWarps);        // + 1 producer warp
constexpr int kStageStride = kStageBytes + 16;             // + slack for the 3-word tap loads

struct __align__(16) WarpSmem {
    unsigned char stage[kStages][kStageStride];
    TileInfo info[kStages];
    unsigned long long full[kStages], empty[kStages];
};

VSTAB_D unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
VSTAB_D void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
VSTAB_D void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
VSTAB_D void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
VSTAB_D void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// TMA engine, 1-D bulk copy global -> shared, completion counted on an mbarrier (UBLKCP in SASS)
VSTAB_D void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Source bounding box (clamped to the image) of the destination tile at (tx0, ty0) under the inverse
// map M, when M is affine and the box fits a stage; otherwise the empty box.
VSTAB_D TileBox tile_box(const double* M, const uint8_t* src, size_t pitch, int w, int h, int tx0, int ty0) {
    TileBox bx{0, 0, 0, 0, 0, 0};                                            // empty: nothing staged, no pixel is "inside"
    const bool al_ok = ((pitch & 15) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    if (M[6] == 0.0 && M[7] == 0.0 && M[8] != 0.0 && al_ok) {
        // every H the stabilizer produces is affine: the tile's source footprint is a
        // parallelogram whose extremes are at the tile corners
        const int tx1 = min(tx0 + TW, w) - 1, ty1 = min(ty0 + TH, h) - 1;
        const double wd = 1.0 / M[8];
        double xmn = 1e300, xmx = -1e300, ymn = 1e300, ymx = -1e300;
        for (int c = 0; c < 4; ++c) {
            const double cx = (c & 1) ? (double)tx1 : (double)tx0, cy = (c & 2) ? (double)ty1 : (double)ty0;
            const double X = (M[0] * cx + M[1] * cy + M[2]) * wd, Y = (M[3] * cx + M[4] * cy + M[5]) * wd;
            xmn = fmin(xmn, X); xmx = fmax(xmx, X); ymn = fmin(ymn, Y); ymx = fmax(ymx, Y);
        }
        if (xmx > -4.0 && ymx > -4.0 && xmn < (double)w + 4.0 && ymn < (double)h + 4.0) {
            const int fx0 = max(0, (int)floor(xmn) - 1), fx1 = min(w - 1, (int)floor(xmx) + 2);
            const int fy0 = max(0, (int)floor(ymn) - 1), fy1 = min(h - 1, (int)floor(ymx) + 2);
            const int b0 = (3 * fx0) & ~15;
            const int b1 = min((int)pitch, (3 * (fx1 + 1) + 15) & ~15);
            const int SP = b1 - b0;
            if (fx1 > fx0 && fy1 > fy0 && SP * (fy1 - fy0 + 1) <= kStageBytes)
                bx = TileBox{fx0, fx1 - fx0, fy0, fy1 - fy0, b0, SP};
        }
    }
    return bx;
}

// The 128 x 32 destination tile at (tx0, ty0): thread (tx, ty) of the 32 x 8 thread tile produces
// 4 consecutive pixels on each of the rows ty, ty+8, ty+16, ty+24.  `sm` holds the staged source
// box (may be empty: box.SP == 0).
VSTAB_D void compute_tile(const uint8_t* __restrict__ sm, const TileBox box, const double* M, const int* border,
                          const uint8_t* __restrict__ src, size_t pitch, int w, int h,
                          uint8_t* __restrict__ dst, size_t out_pitch, int tx0, int ty0, int tx, int ty) {
const int x0 = tx0 + tx * 4;
if (x0 < w) {
        const double M0 = M[0], M1 = M[1], M2 = M[2], M3 = M[3], M4 = M[4], M5 = M[5],
                     M6 = M[6], M7 = M[7], M8 = M[8];
        const bool affine = M6 == 0.0 && M7 == 0.0 && M8 != 0.0;
        const bool al_ok = ((pitch & 15) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
        const double bx = (double)(x0 & ~31);
        const double wdiv = affine ? __ddiv_rn(32.0, M8) : 0.0;
        const bool out_al = (out_pitch & 3) == 0 && ((reinterpret_cast<uintptr_t>(dst) & 3) == 0);
        const int obase = -box.fy0 * box.SP - box.b0;
        double mx1[4], my1[4];                        // M0*x1, M3*x1 for the 4 pixels of this thread
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const double x1 = (double)((x0 + i) & 31);
            mx1[i] = __dmul_rn(M0, x1);
            my1[i] = __dmul_rn(M3, x1);
        }
#pragma unroll
        for (int rr = 0; rr < TH / NTY; ++rr) {
            const int y = ty0 + ty + rr * NTY;
            if (y >= h) break;
            const double yd = (double)y;
            // OpenCV evaluates per 32-px block: X0 = M0*bx + M1*y + M2, then X0 + M0*x1
            const double X0 = __dadd_rn(__dadd_rn(__dmul_rn(M0, bx), __dmul_rn(M1, yd)), M2);
            const double Y0 = __dadd_rn(__dadd_rn(__dmul_rn(M3, bx), __dmul_rn(M4, yd)), M5);
            const double W0 = __dadd_rn(__dadd_rn(__dmul_rn(M6, bx), __dmul_rn(M7, yd)), M8);
            unsigned px[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double X = __dadd_rn(X0, mx1[i]);
                const double Y = __dadd_rn(Y0, my1[i]);
                double Wd = wdiv;
                if (!affine) {
                    const double W = __dadd_rn(W0, __dmul_rn(M6, (double)((x0 + i) & 31)));
                    Wd = W != 0.0 ? __ddiv_rn(32.0, W) : 0.0;
                }
                // saturate_cast<int>(X * 32/W): cvt.rni.s32.f64 saturates like OpenCV's explicit clamp
                const int iX = __double2int_rn(__dmul_rn(X, Wd)), iY = __double2int_rn(__dmul_rn(Y, Wd));
                const int sx = iX >> 5, sy = iY >> 5;
                if ((unsigned)(sx - box.fx0) < (unsigned)box.fxn && (unsigned)(sy - box.fy0) < (unsigned)box.fyn) {
                    const unsigned ax = iX & 31, ay = iY & 31;
                    const int o = sy * box.SP + 3 * sx + obase;
                    unsigned loA, hiA, loB, hiB;
                    staged_row(sm, o, loA, hiA);
                    staged_row(sm, o + box.SP, loB, hiB);
                    // per channel the four taps in one word {p00, p01, p10, p11}
                    const unsigned tA = __byte_perm(loA, hiA, 0x5241);          // G0 G1 R0 R1 of row A
                    const unsigned tB = __byte_perm(loB, hiB, 0x5241);
                    const unsigned WB = __byte_perm(loA, loB, 0x7430);
                    const unsigned WG = __byte_perm(tA, tB, 0x5410);
                    const unsigned WR = __byte_perm(tA, tB, 0x7632);
                    const unsigned wx0 = (32u - ax) | (ax << 8), wx1 = wx0 << 16;
                    const unsigned iay = 32u - ay;
                    // h_r = p_r0 (32-ax) + p_r1 ax  (<= 8160);  out = (h_0 (32-ay) + h_1 ay + 512) >> 10
                    const unsigned vb = (__dp4a(WB, wx0, 0u) * iay + __dp4a(WB, wx1, 0u) * ay + 512u) >> 10;
                    const unsigned vg = (__dp4a(WG, wx0, 0u) * iay + __dp4a(WG, wx1, 0u) * ay + 512u) >> 10;
                    const unsigned vr = (__dp4a(WR, wx0, 0u) * iay + __dp4a(WR, wx1, 0u) * ay + 512u) >> 10;
                    px[i] = vb + (vg << 8) + (vr << 16);
                } else {
                    px[i] = generic_pixel(src, pitch, w, h, al_ok, iX, iY, border);
                }
            }
            uint8_t* o = dst + (size_t)y * out_pitch + (size_t)x0 * 3;
            if (x0 + 4 <= w && out_al) {
                unsigned* o32 = reinterpret_cast<unsigned*>(o);
                o32[0] = px[0] | (px[1] << 24);
                o32[1] = (px[1] >> 8) | (px[2] << 16);
                o32[2] = (px[2] >> 16) | (px[3] << 8);
            } else {
                const int n = min(4, w - x0);
                for (int k = 0; k < n; ++k) {
                    o[3 * k] = (uint8_t)(px[k] & 0xff); o[3 * k + 1] = (uint8_t)((px[k] >> 8) & 0xff); o[3 * k + 2] = (uint8_t)(px[k] >> 16);
                }
            }
        }
    }
}

// Default variant (0): one CTA per tile, synchronous 16-byte loads into the stage; the block scheduler
// overlaps the fill of one CTA with the arithmetic of the others resident on the SM.
__global__ void __launch_bounds__(NTX * NTY, 4)
warp_tile_kernel(const uint8_t* __restrict__ frames, size_t pitch, size_t frame_stride, long slot_mod,
                 const WarpParams* __restrict__ wps, int w, int h,
                 uint8_t* __restrict__ out, size_t out_pitch, size_t out_frame_stride) {
    __shared__ uint4 stage[kStageStride / 16];
    __shared__ TileInfo I;
    const int oi = blockIdx.z;
    const int tid = threadIdx.y * NTX + threadIdx.x;
    const int tx0 = blockIdx.x * TW, ty0 = blockIdx.y * TH;
    if (tid == 0) {
        const WarpParams& P = wps[oi];
#pragma unroll
        for (int i = 0; i < 9; ++i) I.M[i] = P.Minv[i];
        I.border[0] = P.border[0]; I.border[1] = P.border[1]; I.border[2] = P.border[2];
        I.src = frames + (size_t)(slot_mod > 0 ? (P.src_slot % slot_mod) : P.src_slot) * frame_stride;
        I.box = tile_box(I.M, I.src, pitch, w, h, tx0, ty0);
    }
    __syncthreads();
    const TileBox box = I.box;
    if (box.SP > 0) {
        const int vpr = box.SP >> 4;
        for (int r = threadIdx.y; r <= box.fyn; r += NTY) {
            const uint4* g = reinterpret_cast<const uint4*>(I.src + (size_t)(box.fy0 + r) * pitch + box.b0);
            for (int c = threadIdx.x; c < vpr; c += NTX) stage[r * vpr + c] = __ldg(g + c);
        }
    }
    __syncthreads();
    compute_tile(reinterpret_cast<const uint8_t*>(stage), box, I.M, I.border, I.src, pitch, w, h,
                 out + (size_t)oi * out_frame_stride, out_pitch, tx0, ty0, threadIdx.x, threadIdx.y);
}

// Experimental variant (VSTAB_WARP_VARIANT=2; bit-exact, measured slower than variant 0 on B200: 7.2 vs
// 5.7 ms per 512 frames): persistent CTAs, each walking tiles blockIdx.x, +gridDim.x, ... with a two-stage
// shared-memory pipeline: the staged source box of the NEXT tile is fetched with cp.async (LDGSTS, no
// register staging) while the current tile is computed, so the global-load latency of the fill --
// 30 % of the one-tile-per-CTA kernel's stall samples -- is hidden behind the arithmetic.
struct PipeSmem {
    unsigned char stage[2][kStageStride];
    TileInfo info[3];
};

VSTAB_D void make_tile_info(TileInfo& I, long t, int tiles_per_frame, int ntx, const uint8_t* frames, size_t pitch,
                            size_t frame_stride, long slot_mod, const WarpParams* wps, int w, int h) {
    const int oi = (int)(t / tiles_per_frame);
    const int tt = (int)(t - (long)oi * tiles_per_frame);
    const int tyi = tt / ntx, txi = tt - tyi * ntx;
    const WarpParams& P = wps[oi];
#pragma unroll
    for (int i = 0; i < 9; ++i) I.M[i] = P.Minv[i];
    I.border[0] = P.border[0]; I.border[1] = P.border[1]; I.border[2] = P.border[2];
    I.src = frames + (size_t)(slot_mod > 0 ? (P.src_slot % slot_mod) : P.src_slot) * frame_stride;
    I.oi = oi; I.tx0 = txi * TW; I.ty0 = tyi * TH;
    I.box = tile_box(I.M, I.src, pitch, w, h, I.tx0, I.ty0);
}

VSTAB_D void prefetch_box(const TileInfo& I, unsigned char* stage, size_t pitch, int tx, int ty) {
    const TileBox box = I.box;
    if (box.SP > 0) {
        const int vpr = box.SP >> 4;
        for (int r = ty; r <= box.fyn; r += NTY) {
            const uint8_t* g = I.src + (size_t)(box.fy0 + r) * pitch + box.b0;
            const unsigned d = smem_u32(stage + (size_t)r * box.SP);
            for (int c = tx; c < vpr; c += NTX)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + c * 16), "l"(g + c * 16) : "memory");
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

__global__ void __launch_bounds__(NTX * NTY, 3)
warp_pipe_kernel(const uint8_t* __restrict__ frames, size_t pitch, size_t frame_stride, long slot_mod,
                 const WarpParams* __restrict__ wps, int nout, int w, int h,
                 uint8_t* __restrict__ out, size_t out_pitch, size_t out_frame_stride) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PipeSmem& S = *reinterpret_cast<PipeSmem*>(smem_raw);
    const int ntx = (w + TW - 1) / TW, nty = (h + TH - 1) / TH;
    const int tiles_per_frame = ntx * nty;
    const long ntiles = (long)tiles_per_frame * nout;
    const int tid = threadIdx.y * NTX + threadIdx.x;
    const long t0 = blockIdx.x, stride = gridDim.x;
    if (t0 >= ntiles) return;
    if (tid == 0) {
        make_tile_info(S.info[0], t0, tiles_per_frame, ntx, frames, pitch, frame_stride, slot_mod, wps, w, h);
        if (t0 + stride < ntiles)
            make_tile_info(S.info[1], t0 + stride, tiles_per_frame, ntx, frames, pitch, frame_stride, slot_mod, wps, w, h);
    }
    __syncthreads();
    prefetch_box(S.info[0], S.stage[0], pitch, threadIdx.x, threadIdx.y);
    int it = 0;
    for (long t = t0; t < ntiles; t += stride, ++it) {
        const int cur = it & 1;
        const bool has_next = t + stride < ntiles;
        if (has_next) prefetch_box(S.info[(it + 1) % 3], S.stage[cur ^ 1], pitch, threadIdx.x, threadIdx.y);
        if (has_next) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                                   // stage[cur] is complete and visible
        // the tile after next: thread 0 prepares its box while the CTA computes
        if (tid == 0 && t + 2 * stride < ntiles)
            make_tile_info(S.info[(it + 2) % 3], t + 2 * stride, tiles_per_frame, ntx, frames, pitch, frame_stride, slot_mod, wps, w, h);
        const TileInfo& I = S.info[it % 3];
        compute_tile(S.stage[cur], I.box, I.M, I.border, I.src, pitch, w, h, out + (size_t)I.oi * out_frame_stride, out_pitch,
                     I.tx0, I.ty0, threadIdx.x, threadIdx.y);
        __syncthreads();                                   // stage[cur] may be refilled, info[(it+2)%3] is published
    }
}

// Experimental variant (VSTAB_WARP_VARIANT=1; measured slower than the default on B200, see
// DESIGN.md): persistent, warp-specialised: warp 0 walks this CTA's tiles one ahead of the consumers, computes
// each tile's source bounding box and has the TMA engine bulk-copy its rows into the free stage;
// warps 1..8 wait on the stage's mbarrier, produce the 128 x 32 destination pixels and release it.
__global__ void __launch_bounds__(kThreads, 3)
warp_kernel(const uint8_t* __restrict__ frames, size_t pitch, size_t frame_stride, long slot_mod,
            const WarpParams* __restrict__ wps, int nout, int w, int h,
            uint8_t* __restrict__ out, size_t out_pitch, size_t out_frame_stride) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WarpSmem& S = *reinterpret_cast<WarpSmem*>(smem_raw);
    const int ntx = (w + TW - 1) / TW, nty = (h + TH - 1) / TH;
    const int tiles_per_frame = ntx * nty;
    const long ntiles = (long)tiles_per_frame * nout;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) { mbar_init(&S.full[s], 1); mbar_init(&S.empty[s], kConsumerWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == 0) {
        // ================================ producer ==============================================
        int it = 0;
        for (long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int s = it % kStages;
            if (it >= kStages) mbar_wait(&S.empty[s], ((it / kStages) - 1) & 1);
            const int oi = (int)(t / tiles_per_frame);
            const int tt = (int)(t - (long)oi * tiles_per_frame);
            const int tyi = tt / ntx, txi = tt - tyi * ntx;
            const int tx0 = txi * TW, ty0 = tyi * TH;
            TileBox bx{0, 0, 0, 0, 0, 0};                                    // empty: nothing staged, no pixel is "inside"
            const uint8_t* src = nullptr;
            if (lane == 0) {
                const WarpParams& P = wps[oi];
                TileInfo& I = S.info[s];
                double M[9];
#pragma unroll
                for (int i = 0; i < 9; ++i) { M[i] = P.Minv[i]; I.M[i] = M[i]; }
                I.border[0] = P.border[0]; I.border[1] = P.border[1]; I.border[2] = P.border[2];
                src = frames + (size_t)(slot_mod > 0 ? (P.src_slot % slot_mod) : P.src_slot) * frame_stride;
                I.src = src; I.oi = oi; I.tx0 = tx0; I.ty0 = ty0;
                bx = tile_box(M, src, pitch, w, h, tx0, ty0);
                I.box = bx;
                // one arrival (this lane) + the bytes the bulk copies will deliver
                if (bx.SP > 0) mbar_arrive_expect_tx(&S.full[s], (unsigned)(bx.SP * (bx.fyn + 1)));
                else mbar_arrive(&S.full[s]);
            }
            const int SP = __shfl_sync(0xffffffffu, bx.SP, 0);
            if (SP > 0) {
                const int fy0 = __shfl_sync(0xffffffffu, bx.fy0, 0), fyn = __shfl_sync(0xffffffffu, bx.fyn, 0);
                const int b0 = __shfl_sync(0xffffffffu, bx.b0, 0);
                const unsigned long long sp = __shfl_sync(0xffffffffu, (unsigned long long)src, 0);
                const uint8_t* g = reinterpret_cast<const uint8_t*>(sp) + b0;
                for (int r = lane; r <= fyn; r += 32)
                    bulk_g2s(&S.stage[s][r * SP], g + (size_t)(fy0 + r) * pitch, (unsigned)SP, &S.full[s]);
            }
        }
        return;
    }

    // ==================================== consumers ==============================================
    const int tx = lane, ty = warp - 1;                       // 32 x 8 thread tile
    int it = 0;
    for (long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const int s = it % kStages;
        mbar_wait(&S.full[s], (it / kStages) & 1);
        const TileInfo& I = S.info[s];
        const uint8_t* sm = S.stage[s];
        const uint8_t* src = I.src;
        const TileBox box = I.box;
        compute_tile(sm, I.box, I.M, I.border, I.src, pitch, w, h, out + (size_t)I.oi * out_frame_stride, out_pitch,
                     I.tx0, I.ty0, tx, ty);
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.empty[s]);            // this warp is done reading stage s
    }
}

}  // namespace

void launch_warp(const uint8_t* frames, size_t pitch, size_t frame_stride, long slot_mod,
                 const WarpParams* wp, int nout, int w, int h,
                 uint8_t* out, size_t out_pitch, size_t out_frame_stride, cudaStream_t st) {
    if (nout <= 0) return;
    static int num_sms = 0, variant = 0, ctas_per_sm = 3;
    if (num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        cudaFuncSetAttribute(warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WarpSmem));
        cudaFuncSetAttribute(warp_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PipeSmem));
        if (num_sms <= 0) num_sms = 148;
        if (const char* e = getenv("VSTAB_WARP_VARIANT")) variant = atoi(e);
        if (const char* e = getenv("VSTAB_WARP_CTAS")) ctas_per_sm = atoi(e) > 0 ? atoi(e) : 3;
    }
    count_launch(1);
    if (variant == 0) {
        dim3 grid((w + TW - 1) / TW, (h + TH - 1) / TH, nout);
        warp_tile_kernel<<<grid, dim3(NTX, NTY), 0, st>>>(frames, pitch, frame_stride, slot_mod, wp, w, h, out, out_pitch,
                                                          out_frame_stride);
        return;
    }
    const long ntiles = (long)((w + TW - 1) / TW) * ((h + TH - 1) / TH) * nout;
    const int grid = (int)(ntiles < (long)num_sms * ctas_per_sm ? ntiles : (long)num_sms * ctas_per_sm);
    if (variant == 2) {
        warp_pipe_kernel<<<grid, dim3(NTX, NTY), sizeof(PipeSmem), st>>>(frames, pitch, frame_stride, slot_mod, wp, nout, w, h,
                                                                         out, out_pitch, out_frame_stride);
        return;
    }
    warp_kernel<<<grid, kThreads, sizeof(WarpSmem), st>>>(frames, pitch, frame_stride, slot_mod, wp, nout, w, h, out,
                                                          out_pitch, out_frame_stride);
}

}  // namespace vstabk
