// K2: the 8-bit image pyramid that cv::calcOpticalFlowPyrLK builds internally
// (/root/reference/src/stabilizer.cpp:192-195, maxLevel 3): level l+1 = pyrDown(level l),
// separable [1 4 6 4 1], BORDER_REFLECT_101, out = (sum + 128) >> 8 at even sites,
// dst size ((w+1)/2, (h+1)/2).  Bit-exact.
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
        d.pitch[l] = (d.w[l] + 2 * kLkPad + 3) & ~3;
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
constexpr int SW = 2 * TX + 3, SH = 2 * TY + 3; // src footprint 67 x 19

__global__ void __launch_bounds__(TX * TY)
pyrdown_kernel(const uint8_t* __restrict__ pyr, uint8_t* __restrict__ pyr_out, size_t frame_bytes,
               size_t src_off, size_t dst_off, int sw, int sh, int dw, int dh) {
    __shared__ uint8_t tile[SH][SW + 1];
    __shared__ unsigned short hsum[SH][TX];
    const int frame = blockIdx.z;
    const uint8_t* src = pyr + (size_t)frame * frame_bytes + src_off;
    uint8_t* dst = pyr_out + (size_t)frame * frame_bytes + dst_off;
    const int dx0 = blockIdx.x * TX, dy0 = blockIdx.y * TY;
    const int sx0 = 2 * dx0 - 2, sy0 = 2 * dy0 - 2;
    const int tid = threadIdx.y * TX + threadIdx.x;
    for (int i = tid; i < SH * SW; i += TX * TY) {
        const int r = i / SW, c = i - r * SW;
        // rows/cols beyond what the clipped tile needs are still valid reflect indices
        int yy = sy0 + r, xx = sx0 + c;
        yy = reflect101(min(max(yy, -(sh - 1)), 2 * sh - 2), sh);
        xx = reflect101(min(max(xx, -(sw - 1)), 2 * sw - 2), sw);
        tile[r][c] = src[(size_t)yy * sw + xx];
    }
    __syncthreads();
    // horizontal pass: SH rows x TX dst columns
    for (int i = tid; i < SH * TX; i += TX * TY) {
        const int r = i / TX, c = i - r * TX;
        const uint8_t* t = &tile[r][2 * c];
        hsum[r][c] = (unsigned short)(t[0] + 4 * t[1] + 6 * t[2] + 4 * t[3] + t[4]);
    }
    __syncthreads();
    const int x = dx0 + threadIdx.x, y = dy0 + threadIdx.y;
    if (x < dw && y < dh) {
        const int r = 2 * threadIdx.y;
        const int c = threadIdx.x;
        const int s = hsum[r][c] + 4 * hsum[r + 1][c] + 6 * hsum[r + 2][c] + 4 * hsum[r + 3][c] + hsum[r + 4][c];
        dst[(size_t)y * dw + x] = (uint8_t)((s + 128) >> 8);
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
    int tiles_x[kLkLevels];
    int first[kLkLevels + 1];        // first tile of each level
    unsigned off[kLkLevels], poff[kLkLevels], doff[kLkLevels], qoff[kLkLevels];
};

__global__ void __launch_bounds__(PTX * PTY)
lkprep_kernel(uint8_t* __restrict__ pyr, size_t frame_bytes, PrepLevels L) {
    __shared__ uint8_t t[PTY + 2][PTW + 4];
    const int tile = blockIdx.x;
    int l = 0;
#pragma unroll
    for (int i = 1; i < kLkLevels; ++i) l += (tile >= L.first[i]);
    const int w = L.w[l], h = L.h[l], pitch = L.pitch[l];
    const int ti = tile - L.first[l];
    const int ty = ti / L.tiles_x[l], tx = ti - ty * L.tiles_x[l];
    const int X0 = tx * PTW, Y0 = ty * PTY;                  // padded coordinates of the tile
    uint8_t* base = pyr + (size_t)blockIdx.y * frame_bytes;
    const uint8_t* src = base + L.off[l];
    const int tid = threadIdx.y * PTX + threadIdx.x;
    for (int i = tid; i < (PTY + 2) * (PTW + 2); i += PTX * PTY) {
        const int r = i / (PTW + 2), c = i - r * (PTW + 2);
        const int y = reflect101_multi(Y0 + r - 1 - kLkPad, h);
        const int x = reflect101_multi(X0 + c - 1 - kLkPad, w);
        t[r][c] = src[(size_t)y * w + x];
    }
    __syncthreads();
    const int Y = Y0 + threadIdx.y, X = X0 + threadIdx.x * PPX;
    if (Y >= h + 2 * kLkPad || X >= pitch) return;
    const int r = threadIdx.y + 1, c0 = threadIdx.x * PPX + 1;
    const int iy = Y - kLkPad;
    int dq[PPX], qq[PPX];
#pragma unroll
    for (int k = 0; k < PPX; ++k) {
        const int c = c0 + k, ix = X + k - kLkPad;
        qq[k] = (int)((unsigned)t[r][c] | ((unsigned)t[r][c + 1] << 8) | ((unsigned)t[r + 1][c] << 16) | ((unsigned)t[r + 1][c + 1] << 24));
        int dx = 0, dy = 0;
        if (ix >= 0 && ix < w && iy >= 0 && iy < h) {
            const int a00 = t[r - 1][c - 1], a01 = t[r - 1][c], a02 = t[r - 1][c + 1];
            const int a10 = t[r][c - 1], a12 = t[r][c + 1];
            const int a20 = t[r + 1][c - 1], a21 = t[r + 1][c], a22 = t[r + 1][c + 1];
            dx = 3 * (a02 - a00) + 10 * (a12 - a10) + 3 * (a22 - a20);
            dy = 3 * (a20 - a00) + 10 * (a21 - a01) + 3 * (a22 - a02);
        }
        dq[k] = (dx & 0xffff) | (dy << 16);
    }
    // pitch is a multiple of 4 and X is a multiple of 4: aligned vector stores
    *reinterpret_cast<int4*>(base + L.doff[l] + ((size_t)Y * pitch + X) * 4) = make_int4(dq[0], dq[1], dq[2], dq[3]);
    *reinterpret_cast<int4*>(base + L.qoff[l] + ((size_t)Y * pitch + X) * 4) = make_int4(qq[0], qq[1], qq[2], qq[3]);
}

}  // namespace

void launch_pyramid(const PyrDesc& d, uint8_t* pyr, int nframes, cudaStream_t st) {
    if (nframes <= 0) return;
    for (int l = 0; l + 1 < kLkLevels; ++l) {
        dim3 grid((d.w[l + 1] + TX - 1) / TX, (d.h[l + 1] + TY - 1) / TY, nframes);
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
        L.tiles_x[l] = (d.pitch[l] + PTW - 1) / PTW;
        tiles += l < d.nlev ? L.tiles_x[l] * ((d.h[l] + 2 * kLkPad + PTY - 1) / PTY) : 0;
    }
    L.first[kLkLevels] = tiles;
    count_launch(1);
    lkprep_kernel<<<dim3(tiles, nframes), dim3(PTX, PTY), 0, st>>>(pyr, d.frame_bytes, L);
}

}  // namespace vstabk
