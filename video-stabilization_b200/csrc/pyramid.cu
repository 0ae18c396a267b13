// K2: the 8-bit image pyramid that cv::calcOpticalFlowPyrLK builds internally
// (/root/reference/src/stabilizer.cpp:192-195, maxLevel 3): level l+1 = pyrDown(level l),
// separable [1 4 6 4 1], BORDER_REFLECT_101, out = (sum + 128) >> 8 at even sites,
// dst size ((w+1)/2, (h+1)/2).  Bit-exact.
#include <cstdlib>
#include "kernels.h"

namespace vstabk {

PyrDesc make_pyr_desc(int w, int h) {
    PyrDesc d;
    size_t off = 0;
    d.nlev = kLkLevels;
    for (int l = 0; l < kLkLevels; ++l) {
        d.w[l] = w;
        d.h[l] = h;
        d.off[l] = off;
        off += ((size_t)w * h + 255) & ~(size_t)255;   // keep levels 256-byte aligned
        w = (w + 1) / 2;
        h = (h + 1) / 2;
        // buildOpticalFlowPyramid drops a level that is not larger than the window
        if (d.nlev == kLkLevels && l + 1 < kLkLevels && (w <= kLkWin || h <= kLkWin)) d.nlev = l + 1;
    }
    for (int l = 0; l < kLkLevels; ++l) {
        // one row pitch for all levels (that of level 0): the tracker addresses a column of taps as one pointer +
        // immediate multiples of the pitch (lk.cu); rows beyond a level's own padded width are never touched
        d.pitch[l] = (d.w[0] + 2 * kLkPad + 3) & ~3;
        const size_t px = (size_t)d.pitch[l] * (d.h[l] + 2 * kLkPad);
        d.poff[l] = 0;                              // (plain padded u8 levels are not materialised: the quads carry them)
        d.doff[l] = off;
        off += (px * 4 + 255) & ~(size_t)255;
        d.qoff[l] = off;
        off += (px * 4 + 255) & ~(size_t)255;
    }
    d.frame_bytes = off;
    return d;
}

namespace {

constexpr int TX = 32, TY = 8;                 // dst tile

// pyrDown without shared memory: a thread owns one destination column and 4 destination rows.  The 11 source rows it
// needs are filtered horizontally straight from global memory -- interior lanes read the 5 taps as two aligned words
// (funnel shift + one dp4a with the weights {1, 4, 6, 4} + the fifth tap), the lanes at the left / right image edge read
// 5 reflected bytes; row indices are reflected per row (warp-uniform) -- then combined vertically in registers.
// ~27 instructions per destination pixel (a shared-memory tile version with byte staging through reflect indices and two
// barriers needed ~140: 0.18 -> 0.12 ms per 256 frames for the three levels); integer arithmetic as in the header comment.
constexpr int DRY = 4;                           // destination rows per thread
__global__ void __launch_bounds__(TX * TY)
pyrdown_kernel(const uint8_t* __restrict__ pyr, uint8_t* __restrict__ pyr_out, size_t frame_bytes,
                      size_t src_off, size_t dst_off, int sw, int sh, int dw, int dh) {
    const int frame = blockIdx.z;
    const uint8_t* src = pyr + (size_t)frame * frame_bytes + src_off;
    uint8_t* dst = pyr_out + (size_t)frame * frame_bytes + dst_off;
    const int x = blockIdx.x * TX + threadIdx.x;
    const int y0 = (blockIdx.y * TY + threadIdx.y) * DRY;
    if (x >= dw || y0 >= dh) return;
    const int sx = 2 * x - 2;
    const bool word_ok = (sw & 3) == 0 && sx >= 0 && sx + 4 <= sw - 1;
    const int a = sx & ~3;                                       // aligned start of the two words holding taps 0..4
    const unsigned sh16 = (unsigned)(sx & 2) * 8u;               // taps start at byte 0 or 2 of the first word
    const unsigned sel4 = (sx & 2) ? 0x4442u : 0x4440u;          // the fifth tap: byte 2 or byte 0 of the second word
    int xo[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) xo[k] = reflect101(min(max(sx + k, -(sw - 1)), 2 * sw - 2), sw);
    int hs[2 * DRY + 3];
#pragma unroll
    for (int j = 0; j < 2 * DRY + 3; ++j) {
        const int ry = reflect101(min(max(2 * y0 - 2 + j, -(sh - 1)), 2 * sh - 2), sh);
        const uint8_t* row = src + (size_t)ry * sw;
        if (word_ok) {
            const unsigned w0 = __ldg(reinterpret_cast<const unsigned*>(row + a));
            const unsigned w1 = __ldg(reinterpret_cast<const unsigned*>(row + a + 4));
            hs[j] = (int)__dp4a(__funnelshift_r(w0, w1, sh16), 0x04060401u, __byte_perm(w1, 0u, sel4));
        } else {
            hs[j] = (int)__ldg(row + xo[0]) + 4 * (int)__ldg(row + xo[1]) + 6 * (int)__ldg(row + xo[2]) + 4 * (int)__ldg(row + xo[3]) +
                    (int)__ldg(row + xo[4]);
        }
    }
#pragma unroll
    for (int i = 0; i < DRY; ++i) {
        const int y = y0 + i;
        if (y < dh) {
            const int s = hs[2 * i] + 4 * hs[2 * i + 1] + 6 * hs[2 * i + 2] + 4 * hs[2 * i + 3] + hs[2 * i + 4];
            dst[(size_t)y * dw + x] = (uint8_t)((s + 128) >> 8);
        }
    }
}

// LK preparation: for every level the padded intensity image (BORDER_REFLECT_101, what
// cv::buildOpticalFlowPyramid stores) and the padded Scharr derivative image
//   dx = 3(p[-1,+1]-p[-1,-1]) + 10(p[0,+1]-p[0,-1]) + 3(p[+1,+1]-p[+1,-1]),  dy transposed,
// reflect-101 inside the image, zero outside.  One CTA = 128 x 8 padded pixels of
// one level of one frame; the (8+2) x (128+2) source footprint is staged in shared memory.
constexpr int PTX = 32, PTY = 8, PPX = 4;            // threads x, threads y, pixels per thread
constexpr int PTW = PTX * PPX;                        // 128
struct PrepLevels {
    int w[kLkLevels], h[kLkLevels], pitch[kLkLevels];
    int pw[kLkLevels];               // padded width of the level (multiple of 4, <= pitch)
    int tiles_x[kLkLevels];
    unsigned tiles_x_inv[kLkLevels]; // ceil(2^32 / tiles_x): tile / tiles_x == umulhi(tile, inv) for tile < 2^16
    int first[kLkLevels + 1];        // first tile of each level
    unsigned off[kLkLevels], poff[kLkLevels], doff[kLkLevels], qoff[kLkLevels];
};

// u8 pixels x s8 weights
VSTAB_D int dp4a_us(unsigned a, unsigned b, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

constexpr int PTS = PTW + 8;                          // staged row stride (bytes, multiple of 4): PTW + 2 used

__global__ void __launch_bounds__(PTX * PTY)
lkprep_kernel(uint8_t* __restrict__ pyr, size_t frame_bytes, PrepLevels L) {
    __shared__ __align__(16) uint8_t t[PTY + 2][PTS];
    __shared__ int sxs[PTW + 2], sys[PTY + 2];        // reflected source column / row offset of every staged column / row
    const int tile = blockIdx.x;
    int l = 0;
#pragma unroll
    for (int i = 1; i < kLkLevels; ++i) l += (tile >= L.first[i]);
    const int w = L.w[l], h = L.h[l], pitch = L.pitch[l];
    const int ti = tile - L.first[l];
    const int ty = L.tiles_x[l] == 1 ? ti : (int)__umulhi((unsigned)ti, L.tiles_x_inv[l]), tx = ti - ty * L.tiles_x[l];
    const int X0 = tx * PTW, Y0 = ty * PTY;                  // padded coordinates of the tile
    uint8_t* base = pyr + (size_t)blockIdx.y * frame_bytes;
    const uint8_t* src = base + L.off[l];
    const int tid = threadIdx.y * PTX + threadIdx.x;
    // Interior tile (57 % of level 0 at 640 x 360): every tap of every pixel of the tile lies inside the image and the
    // rows are word-aligned, so a thread takes its 3 x 6 bytes straight from global memory as 3 aligned words per row
    // (x0 = X - kLkPad is a multiple of 4: bytes x0-1 .. x0+4 sit in the words at x0-4, x0, x0+4) -- no staging, no
    // reflection tables, no barrier.
    if ((w & 3) == 0 && X0 >= kLkPad + 4 && X0 + PTW - kLkPad <= w - 1 && Y0 >= kLkPad + 1 && Y0 + PTY - kLkPad <= h - 1) {
        const int Y = Y0 + threadIdx.y, X = X0 + threadIdx.x * PPX;
        const int wq = w >> 2;
        const unsigned* rw = reinterpret_cast<const unsigned*>(src + (size_t)(Y - kLkPad - 1) * w + (X - kLkPad));
        unsigned a[3][3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            a[r][0] = __ldg(rw + r * wq - 1); a[r][1] = __ldg(rw + r * wq); a[r][2] = __ldg(rw + r * wq + 1);
        }
        int dq[PPX], qq[PPX];
#pragma unroll
        for (int k = 0; k < PPX; ++k) {
            // columns (c-1, c, c+1, c+2) of the three rows, c = x0 + k
            unsigned t3[3];
#pragma unroll
            for (int r = 0; r < 3; ++r)
                t3[r] = k == 0 ? __funnelshift_r(a[r][0], a[r][1], 24) : k == 1 ? a[r][1] : __funnelshift_r(a[r][1], a[r][2], 8 * (k - 1));
            const unsigned wm = t3[0], w0 = t3[1], wp = t3[2];
            qq[k] = (int)__byte_perm(w0, wp, 0x6521);
            const int dx = dp4a_us(wp, 0x000300fdu, dp4a_us(w0, 0x000a00f6u, dp4a_us(wm, 0x000300fdu, 0)));
            const int dy = dp4a_us(wp, 0x00030a03u, dp4a_us(wm, 0x00fdf6fdu, 0));
            dq[k] = (dx & 0xffff) | (dy << 16);
        }
        *reinterpret_cast<int4*>(base + L.doff[l] + ((size_t)Y * pitch + X) * 4) = make_int4(dq[0], dq[1], dq[2], dq[3]);
        *reinterpret_cast<int4*>(base + L.qoff[l] + ((size_t)Y * pitch + X) * 4) = make_int4(qq[0], qq[1], qq[2], qq[3]);
        return;
    }
    if (tid < PTW + 2) sxs[tid] = reflect101_multi(X0 + tid - 1 - kLkPad, w);
    else if (tid < PTW + 2 + PTY + 2) sys[tid - (PTW + 2)] = reflect101_multi(Y0 + (tid - (PTW + 2)) - 1 - kLkPad, h) * w;
    __syncthreads();
    {
        // all loads of a thread in flight together
        constexpr int kN = ((PTY + 2) * (PTW + 2) + PTX * PTY - 1) / (PTX * PTY);
        uint8_t v[kN];
#pragma unroll
        for (int k = 0; k < kN; ++k) {
            const int i = tid + k * PTX * PTY;
            const int r = i / (PTW + 2), c = i - r * (PTW + 2);
            if (i < (PTY + 2) * (PTW + 2)) v[k] = src[sys[r] + sxs[c]];
        }
#pragma unroll
        for (int k = 0; k < kN; ++k) {
            const int i = tid + k * PTX * PTY;
            const int r = i / (PTW + 2), c = i - r * (PTW + 2);
            if (i < (PTY + 2) * (PTW + 2)) t[r][c] = v[k];
        }
    }
    __syncthreads();
    const int Y = Y0 + threadIdx.y, X = X0 + threadIdx.x * PPX;
    if (Y >= h + 2 * kLkPad || X >= L.pw[l]) return;
    // staged bytes 4*tx .. 4*tx+5 of rows r-1, r, r+1 hold columns c-1 .. c+4 of this thread's 4 pixels
    const int r = threadIdx.y + 1;
    const unsigned* rm = reinterpret_cast<const unsigned*>(&t[r - 1][0]) + threadIdx.x;
    const unsigned* r0 = reinterpret_cast<const unsigned*>(&t[r][0]) + threadIdx.x;
    const unsigned* rp = reinterpret_cast<const unsigned*>(&t[r + 1][0]) + threadIdx.x;
    const unsigned mlo = rm[0], mhi = rm[1], zlo = r0[0], zhi = r0[1], plo = rp[0], phi = rp[1];
    const int iy = Y - kLkPad;
    const bool yin = iy >= 0 && iy < h;
    int dq[PPX], qq[PPX];
#pragma unroll
    for (int k = 0; k < PPX; ++k) {
        // columns (c-1, c, c+1, c+2) of the three rows
        const unsigned wm = __funnelshift_r(mlo, mhi, 8 * k), w0 = __funnelshift_r(zlo, zhi, 8 * k),
                       wp = __funnelshift_r(plo, phi, 8 * k);
        qq[k] = (int)__byte_perm(w0, wp, 0x6521);           // {p(y,x), p(y,x+1), p(y+1,x), p(y+1,x+1)}
        // Scharr: dx = 3 (p[-1,+1] - p[-1,-1]) + 10 (p[0,+1] - p[0,-1]) + 3 (p[+1,+1] - p[+1,-1]),  dy transposed
        const int dx = dp4a_us(wp, 0x000300fdu, dp4a_us(w0, 0x000a00f6u, dp4a_us(wm, 0x000300fdu, 0)));
        const int dy = dp4a_us(wp, 0x00030a03u, dp4a_us(wm, 0x00fdf6fdu, 0));
        const int ix = X + k - kLkPad;
        dq[k] = (yin && ix >= 0 && ix < w) ? ((dx & 0xffff) | (dy << 16)) : 0;
    }
    // pitch is a multiple of 4 and X is a multiple of 4: aligned vector stores
    *reinterpret_cast<int4*>(base + L.doff[l] + ((size_t)Y * pitch + X) * 4) = make_int4(dq[0], dq[1], dq[2], dq[3]);
    *reinterpret_cast<int4*>(base + L.qoff[l] + ((size_t)Y * pitch + X) * 4) = make_int4(qq[0], qq[1], qq[2], qq[3]);
}

}  // namespace

void launch_pyramid(const PyrDesc& d, uint8_t* pyr, int nframes, cudaStream_t st) {
    if (nframes <= 0) return;
    for (int l = 0; l + 1 < kLkLevels; ++l) {
        dim3 grid((d.w[l + 1] + TX - 1) / TX, (d.h[l + 1] + TY * DRY - 1) / (TY * DRY), nframes);
        dim3 block(TX, TY);
        count_launch(1);
        pyrdown_kernel<<<grid, block, 0, st>>>(pyr, pyr, d.frame_bytes, d.off[l], d.off[l + 1],
                                               d.w[l], d.h[l], d.w[l + 1], d.h[l + 1]);
    }
    PrepLevels L;
    int tiles = 0;
    for (int l = 0; l < kLkLevels; ++l) {
        L.w[l] = d.w[l]; L.h[l] = d.h[l]; L.pitch[l] = d.pitch[l];
        L.off[l] = (unsigned)d.off[l]; L.poff[l] = (unsigned)d.poff[l]; L.doff[l] = (unsigned)d.doff[l]; L.qoff[l] = (unsigned)d.qoff[l];
        L.first[l] = tiles;
        L.pw[l] = (d.w[l] + 2 * kLkPad + 3) & ~3;
        L.tiles_x[l] = (L.pw[l] + PTW - 1) / PTW;
        L.tiles_x_inv[l] = (unsigned)(((1ull << 32) + L.tiles_x[l] - 1) / L.tiles_x[l]);
        tiles += l < d.nlev ? L.tiles_x[l] * ((d.h[l] + 2 * kLkPad + PTY - 1) / PTY) : 0;
    }
    L.first[kLkLevels] = tiles;
    count_launch(1);
    lkprep_kernel<<<dim3(tiles, nframes), dim3(PTX, PTY), 0, st>>>(pyr, d.frame_bytes, L);
}

}  // namespace vstabk
