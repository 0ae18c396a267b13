// K8: image conditioning of the ORB / SIFT registration path
//   /root/reference/src/stabilizer.cpp:448-477
//     resize(INTER_NEAREST) of the full-resolution presentation frame -> working size   (:450-451)
//     cvtColor(BGR2GRAY)                                                               (:455)
//     medianBlur(5)                          BORDER_REPLICATE                          (:464)
//     filter2D [0 -1 0; -1 5 -1; 0 -1 0]     8U saturate, BORDER_REFLECT_101            (:466-470)
//     CLAHE(clipLimit 2.0, tiles 8x8)                                                  (:472-475)
//     medianBlur(5)                                                                    (:477)
// All stages are integer (or single-rounded float) and bit-exact against cv2 4.13.0.
// The sharpened image is never materialised: the histogram and the CLAHE apply kernels recompute
// it from the median image (5 loads per pixel).
#include <cmath>
#include <vector>
#include "kernels.h"

namespace vstabk {
namespace {

constexpr int MTX = 32, MTY = 8, MPX = 4;     // threads x, threads y, pixels per thread along x
constexpr int MTW = MTX * MPX;                 // 128-pixel tile rows
constexpr int MSW = MTW + 8;                   // staged row: columns x0-4 .. x0+MTW+3 (word-aligned start), 34 words

// 25-element median selection network on TWO pixels at once: every register holds the same window position of two
// horizontally adjacent pixels as u16x2, and a comparator is one VIMNMX.U16x2 pair.
VSTAB_D unsigned median25x2(unsigned* v) {
#define CS(a, b) { const unsigned lo_ = __vminu2(v[a], v[b]); v[b] = __vmaxu2(v[a], v[b]); v[a] = lo_; }
#include "median25.inc"
#undef CS
    return v[12];
}

__global__ void __launch_bounds__(256)
nn_gray_kernel(const uint8_t* __restrict__ frame, size_t pitch, const int* __restrict__ xofs,
               const int* __restrict__ yofs, int w, int h, uint8_t* __restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    const uint8_t* p = frame + (size_t)yofs[y] * pitch + (size_t)xofs[x] * 3;
    out[(size_t)y * w + x] = (uint8_t)luma_q15(p[0], p[1], p[2]);
}

// 5x5 median, BORDER_REPLICATE.  A CTA stages a (8+4) x (128+8) tile as words (aligned loads where the word lies inside
// the row, clamped bytes at the image edges); a thread produces 4 pixels of a row as two u16x2 pairs: per source row three
// words + two funnel shifts give the 8 bytes its windows span, five PRMT per pair spread them into u16x2 lanes.
__global__ void __launch_bounds__(MTX * MTY)
median5_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int w, int h) {
    __shared__ __align__(16) unsigned t[MTY + 4][MSW / 4];
    const int x0 = blockIdx.x * MTW, y0 = blockIdx.y * MTY;
    const int tid = threadIdx.y * MTX + threadIdx.x;
    const bool words_ok = (w & 3) == 0 && ((reinterpret_cast<uintptr_t>(in) & 3) == 0);
    for (int i = tid; i < (MTY + 4) * (MSW / 4); i += MTX * MTY) {
        const int r = i / (MSW / 4), c = i - r * (MSW / 4);
        const int yy = min(max(y0 + r - 2, 0), h - 1), xb = x0 - 4 + 4 * c;
        const uint8_t* row = in + (size_t)yy * w;
        unsigned v;
        if (words_ok && xb >= 0 && xb + 3 <= w - 1) {
            v = __ldg(reinterpret_cast<const unsigned*>(row + xb));
        } else {
            v = (unsigned)row[min(max(xb, 0), w - 1)] | ((unsigned)row[min(max(xb + 1, 0), w - 1)] << 8) |
                ((unsigned)row[min(max(xb + 2, 0), w - 1)] << 16) | ((unsigned)row[min(max(xb + 3, 0), w - 1)] << 24);
        }
        t[r][c] = v;
    }
    __syncthreads();
    const int x = x0 + threadIdx.x * MPX, y = y0 + threadIdx.y;
    if (x >= w || y >= h) return;
    // the windows of pixels x .. x+3 span columns x-2 .. x+5 = staged bytes 4 tx + 2 .. 4 tx + 9
    unsigned va[25], vb[25];
#pragma unroll
    for (int dy = 0; dy < 5; ++dy) {
        const unsigned* rp = &t[threadIdx.y + dy][threadIdx.x];
        const unsigned q0 = rp[0], q1 = rp[1], q2 = rp[2];
        const unsigned w0 = __funnelshift_r(q0, q1, 16), w1 = __funnelshift_r(q1, q2, 16);   // bytes 0..3, 4..7 of the span
        const unsigned wm = __funnelshift_r(w0, w1, 16);                                        // bytes 2..5
        // pair A = pixels (x, x+1): window position dx holds span bytes (dx, dx+1); pair B = (x+2, x+3): bytes (dx+2, dx+3)
        va[dy * 5 + 0] = __byte_perm(w0, 0u, 0x4140); va[dy * 5 + 1] = __byte_perm(w0, 0u, 0x4241);
        va[dy * 5 + 2] = __byte_perm(w0, 0u, 0x4342); va[dy * 5 + 3] = __byte_perm(wm, 0u, 0x4241);
        va[dy * 5 + 4] = __byte_perm(w1, 0u, 0x4140);
        vb[dy * 5 + 0] = __byte_perm(w0, 0u, 0x4342); vb[dy * 5 + 1] = __byte_perm(wm, 0u, 0x4241);
        vb[dy * 5 + 2] = __byte_perm(w1, 0u, 0x4140); vb[dy * 5 + 3] = __byte_perm(w1, 0u, 0x4241);
        vb[dy * 5 + 4] = __byte_perm(w1, 0u, 0x4342);
    }
    const unsigned ma = median25x2(va), mb = median25x2(vb);
    const unsigned res = __byte_perm(ma, mb, 0x6420);          // {A.lo, A.hi, B.lo, B.hi} low bytes
    uint8_t* o = out + (size_t)y * w + x;
    if (x + 3 < w && ((reinterpret_cast<uintptr_t>(o) & 3) == 0)) {
        *reinterpret_cast<unsigned*>(o) = res;
    } else {
        for (int k = 0; k < 4 && x + k < w; ++k) o[k] = (uint8_t)(res >> (8 * k));
    }
}

// sharpen at (x, y) of the median image `m` (BORDER_REFLECT_101), saturated to u8
VSTAB_D int sharpen_at(const uint8_t* __restrict__ m, int w, int h, int x, int y) {
    const int xl = reflect101(x - 1, w), xr = reflect101(x + 1, w);
    const int yu = reflect101(y - 1, h), yd = reflect101(y + 1, h);
    const int c = m[(size_t)y * w + x];
    const int s = 5 * c - m[(size_t)yu * w + x] - m[(size_t)yd * w + x] - m[(size_t)y * w + xl] - m[(size_t)y * w + xr];
    return min(255, max(0, s));
}

// per-tile 256-bin histogram of the sharpened image; when the image size is not a multiple of the
// tile grid OpenCV extends it to the right / bottom with BORDER_REFLECT_101 first
__global__ void __launch_bounds__(256)
clahe_hist_kernel(const uint8_t* __restrict__ m, int w, int h, int tw, int th, int tiles_x,
                  unsigned int* __restrict__ hist) {
    __shared__ unsigned int sh[256];
    const int tile = blockIdx.x, part = blockIdx.y, nparts = gridDim.y;
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    sh[threadIdx.x] = 0;
    __syncthreads();
    const int rows0 = (int)(((long long)part * th) / nparts), rows1 = (int)(((long long)(part + 1) * th) / nparts);
    const int n = (rows1 - rows0) * tw;
    for (int i = threadIdx.x; i < n; i += 256) {
        const int r = i / tw, c = i - r * tw;
        const int y = reflect101(ty * th + rows0 + r, h), x = reflect101(tx * tw + c, w);
        atomicAdd(&sh[sharpen_at(m, w, h, x, y)], 1u);
    }
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(&hist[tile * 256 + threadIdx.x], sh[threadIdx.x]);
}

// clip, redistribute, cumulative LUT (one CTA of 256 threads per tile)
__global__ void __launch_bounds__(256)
clahe_lut_kernel(unsigned int* __restrict__ hist, int tile_area, int clip_limit, uint8_t* __restrict__ lut) {
    __shared__ int hsh[256];
    __shared__ int wsum[8];
    __shared__ int s_clipped;
    const int tile = blockIdx.x, i = threadIdx.x;
    int hv = (int)hist[tile * 256 + i];
    hist[tile * 256 + i] = 0;                                  // ready for the next frame
    int excess = hv > clip_limit ? hv - clip_limit : 0;
    hv = min(hv, clip_limit);
    // total clipped
    int e = excess;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    if ((i & 31) == 0) wsum[i >> 5] = e;
    __syncthreads();
    if (i == 0) { int t = 0; for (int k = 0; k < 8; ++k) t += wsum[k]; s_clipped = t; }
    __syncthreads();
    const int clipped = s_clipped;
    const int batch = clipped / 256;
    int residual = clipped - batch * 256;
    hv += batch;
    if (residual != 0) {
        const int step = max(256 / residual, 1);
        // for (k = 0; k < 256 && residual > 0; k += step, residual--) hist[k]++
        if (i % step == 0 && i / step < residual) hv += 1;
    }
    hsh[i] = hv;
    __syncthreads();
    // inclusive prefix sum (Hillis-Steele in shared memory)
    for (int o = 1; o < 256; o <<= 1) {
        const int t = i >= o ? hsh[i - o] : 0;
        __syncthreads();
        hsh[i] += t;
        __syncthreads();
    }
    const float lut_scale = (float)(256 - 1) / (float)tile_area;
    const int v = __float2int_rn(__fmul_rn((float)hsh[i], lut_scale));
    lut[tile * 256 + i] = (uint8_t)min(255, max(0, v));
}

__global__ void __launch_bounds__(256)
clahe_apply_kernel(const uint8_t* __restrict__ m, int w, int h, int tw, int th, int tiles_x, int tiles_y,
                   const uint8_t* __restrict__ lut, uint8_t* __restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    const float inv_tw = __fdiv_rn(1.0f, (float)tw), inv_th = __fdiv_rn(1.0f, (float)th);
    const float tyf = __fsub_rn(__fmul_rn((float)y, inv_th), 0.5f);
    int ty1 = (int)floorf(tyf);
    int ty2 = ty1 + 1;
    const float ya = __fsub_rn(tyf, (float)ty1), ya1 = __fsub_rn(1.0f, ya);
    ty1 = max(ty1, 0); ty2 = min(ty2, tiles_y - 1);
    const float txf = __fsub_rn(__fmul_rn((float)x, inv_tw), 0.5f);
    int tx1 = (int)floorf(txf);
    int tx2 = tx1 + 1;
    const float xa = __fsub_rn(txf, (float)tx1), xa1 = __fsub_rn(1.0f, xa);
    tx1 = max(tx1, 0); tx2 = min(tx2, tiles_x - 1);
    const int v = sharpen_at(m, w, h, x, y);
    const float l11 = (float)lut[(ty1 * tiles_x + tx1) * 256 + v], l12 = (float)lut[(ty1 * tiles_x + tx2) * 256 + v];
    const float l21 = (float)lut[(ty2 * tiles_x + tx1) * 256 + v], l22 = (float)lut[(ty2 * tiles_x + tx2) * 256 + v];
    const float top = __fadd_rn(__fmul_rn(l11, xa1), __fmul_rn(l12, xa));
    const float bot = __fadd_rn(__fmul_rn(l21, xa1), __fmul_rn(l22, xa));
    const float res = __fadd_rn(__fmul_rn(top, ya1), __fmul_rn(bot, ya));
    out[(size_t)y * w + x] = (uint8_t)min(255, max(0, __float2int_rn(res)));
}

}  // namespace

void build_nn_table(int src, int dst, int* tab) {
    // cv::resize INTER_NEAREST: x_ofs[x] = min(cvFloor(x * ifx), src - 1), ifx = 1 / (dst / src) in double
    const double inv_scale = (double)dst / (double)src;
    const double ifx = 1.0 / inv_scale;
    for (int d = 0; d < dst; ++d) {
        int s = (int)std::floor(d * ifx);
        tab[d] = s < src - 1 ? s : src - 1;
    }
}

size_t featprep_workspace_bytes(int w, int h) {
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    return al((size_t)w * h) * 2 + al(64 * 256 * 4) + al(64 * 256);
}

void launch_featprep(const uint8_t* frame, size_t pitch, const int* xofs, const int* yofs, int w, int h,
                     void* workspace, uint8_t* out, cudaStream_t st) {
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    uint8_t* a = (uint8_t*)workspace;
    uint8_t* b = a + al((size_t)w * h);
    unsigned int* hist = (unsigned int*)(b + al((size_t)w * h));
    uint8_t* lut = (uint8_t*)hist + al(64 * 256 * 4);
    const int tiles = 8;
    // CLAHE tile size from the (possibly extended) image
    // (when either dimension is not a multiple of 8 OpenCV pads BOTH by 8 - (dim % 8), i.e. a full
    // extra 8 pixels on an already divisible axis -- CLAHE_Impl::apply)
    const bool divisible = (w % tiles) == 0 && (h % tiles) == 0;
    const int ew = divisible ? w : w + tiles - (w % tiles), eh = divisible ? h : h + tiles - (h % tiles);
    const int tw = ew / tiles, th = eh / tiles;
    const int tile_area = tw * th;
    int clip = (int)(2.0 * tile_area / 256);
    if (clip < 1) clip = 1;
    count_launch(6);
    nn_gray_kernel<<<dim3((w + 255) / 256, h), 256, 0, st>>>(frame, pitch, xofs, yofs, w, h, a);
    dim3 mg((w + MTW - 1) / MTW, (h + MTY - 1) / MTY), mb(MTX, MTY);
    median5_kernel<<<mg, mb, 0, st>>>(a, b, w, h);
    cudaMemsetAsync(hist, 0, 64 * 256 * 4, st);
    clahe_hist_kernel<<<dim3(64, 8), 256, 0, st>>>(b, w, h, tw, th, tiles, hist);
    clahe_lut_kernel<<<64, 256, 0, st>>>(hist, tile_area, clip, lut);
    clahe_apply_kernel<<<dim3((w + 255) / 256, h), 256, 0, st>>>(b, w, h, tw, th, tiles, tiles, lut, a);
    median5_kernel<<<mg, mb, 0, st>>>(a, out, w, h);
}

}  // namespace vstabk
