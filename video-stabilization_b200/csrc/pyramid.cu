// K2: the 8-bit image pyramid that cv::calcOpticalFlowPyrLK builds internally
// (/root/reference/src/stabilizer.cpp:192-195, maxLevel 3): level l+1 = pyrDown(level l),
// separable [1 4 6 4 1], BORDER_REFLECT_101, out = (sum + 128) >> 8 at even sites,
// dst size ((w+1)/2, (h+1)/2).  Bit-exact.
#include "kernels.h"

namespace vstabk {

PyrDesc make_pyr_desc(int w, int h) {
    PyrDesc d;
    size_t off = 0;
    for (int l = 0; l < kLkLevels; ++l) {
        d.w[l] = w;
        d.h[l] = h;
        d.off[l] = off;
        off += ((size_t)w * h + 255) & ~(size_t)255;   // keep levels 256-byte aligned
        w = (w + 1) / 2;
        h = (h + 1) / 2;
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
}

}  // namespace vstabk
