// Shared helpers for the sm_100a kernels of the stabilizer hot path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vstabk {

#define VSTAB_HD __host__ __device__ __forceinline__
#define VSTAB_D __device__ __forceinline__

constexpr int kMaxCorners = 1300;          // MAX_FEATURES_TO_DETECT, src/stabilizer.cpp:935
constexpr int kLkWin = 21;                 // WINDOW_SIZE, src/stabilizer.cpp:185
constexpr int kLkLevels = 4;               // MAX_PYRAMID_LEVEL 3 => 4 levels, :186
constexpr int kLkMaxIter = 50;             // TERM_CRITERIA count, :187-188
constexpr int kMinPointsForMotion = 10;    // MIN_POINTS_FOR_MOTION_ESTIMATION, :20

// BORDER_REFLECT_101 index for i in [-n+1, 2n-2]
VSTAB_HD int reflect101(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return i;
}

// cv::COLOR_BGR2GRAY 8U: (3735 B + 19235 G + 9798 R + 16384) >> 15   
VSTAB_HD int luma_q15(int b, int g, int r) {
    return (3735 * b + 19235 * g + 9798 * r + 16384) >> 15;
}

// OpenCV pads every LK pyramid level by the window size; kLkPad >= 22 covers every tap the
// tracker may touch (window origins in [-21, cols-1], taps up to origin + 22).
constexpr int kLkPad = 24;

// BORDER_REFLECT_101 for any i (cv::borderInterpolate loops for small images), n >= 2
VSTAB_HD int reflect101_multi(int i, int n) {
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
}

// A device-resident gray pyramid of one frame.  Level l exists three times inside the frame's
// block: tight (w x h u8: pyrDown chain, corner detector, taps), padded by kLkPad with
// BORDER_REFLECT_101 (row pitch `pitch[l]`) as 4-byte bilinear "quads" (the 2x2 neighbourhood of
// every pixel in one word, so a bilinear tap is one load + two dp2a) and its Scharr derivative
// image, padded with zeros (short2 {dx, dy}, same geometry) -- the two buffers
// cv::calcOpticalFlowPyrLK tracks on.
struct PyrDesc {
    int w[kLkLevels];
    int h[kLkLevels];
    int pitch[kLkLevels];    // padded row pitch in pixels (multiple of 4)
    int nlev;                // levels OpenCV keeps: next level must be larger than the LK window in both dims
    size_t off[kLkLevels];   // byte offset of the tight level l inside one frame's pyramid block
    size_t poff[kLkLevels];  // byte offset of padded pixel (-kLkPad, -kLkPad) of level l
    size_t doff[kLkLevels];  // byte offset of the padded derivative image (4 bytes per pixel)
    size_t qoff[kLkLevels];  // byte offset of the padded "quad" image: {p(y,x), p(y,x+1), p(y+1,x), p(y+1,x+1)} per pixel
    size_t frame_bytes;      // bytes of one frame's pyramid block
};

template <typename T>
VSTAB_D T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

VSTAB_D long long warp_sum_ll(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace vstabk
