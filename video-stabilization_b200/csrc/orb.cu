// K9 / K10: ORB keypoints + rBRIEF descriptors and brute-force Hamming 2-NN matching
//   cv::ORB::create(2500, 1.2f, 12, 31, 0, 2, ORB::FAST_SCORE, 31, 20)->detectAndCompute
//       /root/reference/src/stabilizer.cpp:483-491, :537-545, :557-559, :605-606
//   filterKeypointByRelativeSize(ratio 0.10)                     :290-309, :608-612
//   cv::BFMatcher(NORM_HAMMING).knnMatch(ref, cur, 2) + Lowe ratio 0.6   :647-673
// Integer work, bit-exact against cv2 4.13.0:
//   pyramid   level l = INTER_LINEAR_EXACT resize of level l-1 (Q8 coefficients from exact
//             rational source coordinates, computed on the host in double)
//   FAST-9/16 threshold 20: score = max over the 16 arcs of 9 of min(v - ring) / min(ring - v), - 1;
//             3x3 non-maximum suppression (strict), 31-px border, per-level retainBest with
//             ties kept (score histogram -> threshold)
//   angle     intensity centroid over the radius-15 disc, cv::fastAtan2 polynomial (float)
//   blur      7x7 sigma 2 Gaussian in float (separable, FMA chains), rounded half-even to u8
//   rBRIEF    256 comparisons on the steered pattern, one warp per keypoint (lane = byte)
//   matching  __popc over 8 x u32, running best-2 per reference row, strict '<' in ascending
//             train order (ties -> lowest index), ratio test on the integer distances in float
// Keypoints are emitted level-major and row-major inside a level (OpenCV's order inside a level
// is the unspecified permutation left by std::nth_element; the set is identical).
#include <cub/device/device_radix_sort.cuh>
#include <cmath>
#include <cstring>
#include <vector>
#include "kernels.h"

namespace vstabk {
namespace {

__constant__ signed char c_pattern[256 * 4] = {
#include "orb_pattern.inc"
};
__constant__ int c_umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};

constexpr int kEdge = 31;                 // edgeThreshold
constexpr int kFastThr = 20;

// ---------------------------------------------------------------- pyramid (INTER_LINEAR_EXACT)
__global__ void __launch_bounds__(256)
orb_resize_kernel(const uint8_t* __restrict__ src, int sw, int sh, uint8_t* __restrict__ dst, int dw, int dh,
                  const int2* __restrict__ xtab, const int2* __restrict__ ytab) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= dw) return;
    const int2 tx = xtab[x], ty = ytab[y];             // {index, alpha (0..256)}
    const uint8_t* r0 = src + (size_t)ty.x * sw;
    const uint8_t* r1 = src + (size_t)min(ty.x + 1, sh - 1) * sw;
    const int x1 = min(tx.x + 1, sw - 1);
    const unsigned t0 = r0[tx.x] * (256 - tx.y) + r0[x1] * tx.y;       // Q8.8
    const unsigned t1 = r1[tx.x] * (256 - tx.y) + r1[x1] * tx.y;
    dst[(size_t)y * dw + x] = (uint8_t)((t0 * (256 - ty.y) + t1 * ty.y + 32768u) >> 16);
}

struct OrbLevels {
    int w[kOrbLevels], h[kOrbLevels];
    unsigned off[kOrbLevels];            // byte offset of level l in the image / blur / score buffers
    int first_tile[kOrbLevels + 1];      // 32 x 8 tiles, levels back to back
    int tiles_x[kOrbLevels];
    int quota[kOrbLevels];
    float scale[kOrbLevels];
    int nlevels_used;                    // levels that survive the relative-size filter
};

VSTAB_D int level_of_tile(const OrbLevels& L, int tile) {
    int l = 0;
#pragma unroll
    for (int i = 1; i < kOrbLevels; ++i) l += (tile >= L.first_tile[i]);
    return l;
}

// ---------------------------------------------------------------- FAST score map
constexpr int FTX = 32, FTY = 8;
__global__ void __launch_bounds__(FTX * FTY)
fast_score_kernel(const uint8_t* __restrict__ pyr, uint8_t* __restrict__ score, OrbLevels L) {
    __shared__ uint8_t t[FTY + 6][FTX + 8];
    const int l = level_of_tile(L, blockIdx.x);
    const int w = L.w[l], h = L.h[l];
    const int ti = blockIdx.x - L.first_tile[l];
    const int ty = ti / L.tiles_x[l], tx = ti - ty * L.tiles_x[l];
    const int x0 = tx * FTX, y0 = ty * FTY;
    const uint8_t* img = pyr + L.off[l];
    const int tid = threadIdx.y * FTX + threadIdx.x;
    for (int i = tid; i < (FTY + 6) * (FTX + 6); i += FTX * FTY) {
        const int r = i / (FTX + 6), c = i - r * (FTX + 6);
        const int yy = min(max(y0 + r - 3, 0), h - 1), xx = min(max(x0 + c - 3, 0), w - 1);
        t[r][c] = img[(size_t)yy * w + xx];
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x >= w || y >= h) return;
    int s = 0;
    if (x >= 3 && x < w - 3 && y >= 3 && y < h - 3) {
        const int r = threadIdx.y + 3, c = threadIdx.x + 3;
        const int v = t[r][c];
        // ring offsets (dx, dy), k = 0..15
        const int d0 = v - t[r + 3][c], d4 = v - t[r][c + 3], d8 = v - t[r - 3][c], d12 = v - t[r][c - 3];
        // high-speed test: a 9-arc contains at least 2 of the 4 compass points... with the same sign
        const int nb = (d0 > kFastThr) + (d4 > kFastThr) + (d8 > kFastThr) + (d12 > kFastThr);
        const int nd = (d0 < -kFastThr) + (d4 < -kFastThr) + (d8 < -kFastThr) + (d12 < -kFastThr);
        if (nb >= 2 || nd >= 2) {
            int d[16];
            d[0] = d0; d[4] = d4; d[8] = d8; d[12] = d12;
            d[1] = v - t[r + 3][c + 1]; d[2] = v - t[r + 2][c + 2]; d[3] = v - t[r + 1][c + 3];
            d[5] = v - t[r - 1][c + 3]; d[6] = v - t[r - 2][c + 2]; d[7] = v - t[r - 3][c + 1];
            d[9] = v - t[r - 3][c - 1]; d[10] = v - t[r - 2][c - 2]; d[11] = v - t[r - 1][c - 3];
            d[13] = v - t[r + 1][c - 3]; d[14] = v - t[r + 2][c - 2]; d[15] = v - t[r + 3][c - 1];
            // min / max over every circular window of 9: m2 -> m4 -> m8 -> m9
            int mn[16], mx[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) { mn[k] = min(d[k], d[(k + 1) & 15]); mx[k] = max(d[k], d[(k + 1) & 15]); }
            int mn4[16], mx4[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) { mn4[k] = min(mn[k], mn[(k + 2) & 15]); mx4[k] = max(mx[k], mx[(k + 2) & 15]); }
            int best = -256, worst = 256;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int a = min(min(mn4[k], mn4[(k + 4) & 15]), d[(k + 8) & 15]);     // min of d[k..k+8]
                const int b = max(max(mx4[k], mx4[(k + 4) & 15]), d[(k + 8) & 15]);     // max of d[k..k+8]
                best = max(best, a);
                worst = min(worst, b);
            }
            const int sc = max(best, -worst);       // largest t' with 9 contiguous |diff| >= t' of one sign
            if (sc > kFastThr) s = sc - 1;
        }
    }
    score[L.off[l] + (size_t)y * w + x] = (uint8_t)s;
}

// ---------------------------------------------------------------- NMS + border filter -> candidates + score histogram
__global__ void __launch_bounds__(256)
fast_nms_kernel(const uint8_t* __restrict__ score, OrbLevels L, unsigned int* __restrict__ hist /* [levels][256] */,
                unsigned long long* __restrict__ cand, int* __restrict__ ncand, int cap) {
    const int l = blockIdx.y;
    if (l >= L.nlevels_used) return;
    const int w = L.w[l], h = L.h[l];
    const int iw = w - 2 * kEdge, ih = h - 2 * kEdge;
    if (iw <= 0 || ih <= 0) return;
    const uint8_t* sc = score + L.off[l];
    const int n = iw * ih;
    for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        const int i = base + threadIdx.x;
        bool keep = false;
        int x = 0, y = 0, s = 0;
        if (i < n) {
            y = kEdge + i / iw;
            x = kEdge + (i - (y - kEdge) * iw);
            s = sc[(size_t)y * w + x];
            if (s > 0) {
                keep = true;
#pragma unroll
                for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                    for (int dx = -1; dx <= 1; ++dx)
                        if ((dx | dy) != 0 && sc[(size_t)(y + dy) * w + (x + dx)] >= s) keep = false;
            }
        }
        if (keep) atomicAdd(&hist[l * 256 + s], 1u);
        const unsigned ball = __ballot_sync(0xffffffffu, keep);
        if (ball) {
            const int lane = threadIdx.x & 31;
            int pos0 = 0;
            if (lane == 0) pos0 = atomicAdd(ncand, __popc(ball));
            pos0 = __shfl_sync(0xffffffffu, pos0, 0);
            if (keep) {
                const int pos = pos0 + __popc(ball & ((1u << lane) - 1u));
                // key: level | y | x | score  (ascending sort = level-major, row-major)
                if (pos < cap)
                    cand[pos] = ((unsigned long long)l << 56) | ((unsigned long long)y << 36) | ((unsigned long long)x << 16) |
                                (unsigned long long)s;
            }
        }
    }
}

// retainBest per level: threshold = score of the quota-th best; everything >= threshold stays (ties kept).
// Candidates below their level's threshold get the key ~0 (sorted to the end).
__global__ void __launch_bounds__(256)
orb_threshold_kernel(const unsigned int* __restrict__ hist, OrbLevels L, int* __restrict__ thr) {
    const int l = threadIdx.x;
    if (l >= kOrbLevels) return;
    int t = 0, acc = 0;
    if (l < L.nlevels_used) {
        const int quota = L.quota[l];
        t = 1;
        for (int s = 255; s >= 1; --s) {
            acc += (int)hist[l * 256 + s];
            if (acc >= quota) { t = s; break; }
        }
        if (quota <= 0) t = 256;          // KeyPointsFilter::retainBest(0) clears
    } else {
        t = 256;
    }
    thr[l] = t;
}

// keys beyond the candidate count sort to the end
__global__ void __launch_bounds__(256)
orb_tail_kernel(unsigned long long* __restrict__ cand, const int* __restrict__ ncand, int cap) {
    const int n = min(*ncand, cap);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n && i < cap) cand[i] = ~0ull;
}

// Stable compaction of the sorted candidate list (level-major, row-major) to the keypoints that
// survive retainBest: score >= threshold of their level.  Single CTA.
__global__ void __launch_bounds__(1024)
orb_compact_kernel(const unsigned long long* __restrict__ sorted, const int* __restrict__ ncand, int cap,
                   const int* __restrict__ thr, unsigned long long* __restrict__ kept, int* __restrict__ nkept, int max_kp) {
    __shared__ int wsum[32];
    __shared__ int s_base;
    const int n = min(*ncand, cap);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        unsigned long long k = 0;
        bool keep = false;
        if (i < n) { k = sorted[i]; keep = (int)(k & 0xffffu) >= thr[(int)(k >> 56)]; }
        const unsigned ball = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) wsum[wid] = __popc(ball);
        __syncthreads();
        int off = s_base;
        for (int w = 0; w < wid; ++w) off += wsum[w];
        if (keep) {
            const int pos = off + __popc(ball & ((1u << lane) - 1u));
            if (pos < max_kp) kept[pos] = k;
        }
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 32; ++w) t += wsum[w]; s_base += t; }
        __syncthreads();
    }
    if (threadIdx.x == 0) *nkept = s_base;
}

// ---- OpenCV's keypoint ORDER inside a level (reference capture only) -----------------------------
// KeyPointsFilter::retainBest leaves the survivors in the permutation produced by
//   std::nth_element(begin, begin + quota - 1, end, response >)  followed by
//   std::partition(begin + quota, end, response >= response[quota - 1])
// (libstdc++: introselect with median-of-3 to *first, unguarded partition, final insertion sort).
// The order of the REFERENCE keypoints fixes the order of the match list and hence the RANSAC
// sample sequence of estimateAffinePartial2D, so the reference set is put into exactly that order:
// one thread per level replays the algorithm on (score, row-major rank) words in shared memory.
// The current frame's keypoints never need it (their order does not reach the estimator).
struct Sel {
    unsigned* a;
    __device__ bool gt(int i, int j) const { return (a[i] >> 24) > (a[j] >> 24); }          // KeypointResponseGreater
    __device__ void swp(int i, int j) { const unsigned t = a[i]; a[i] = a[j]; a[j] = t; }
};

__device__ void sel_move_median_to_first(Sel& S, int result, int a, int b, int c) {
    if (S.gt(a, b)) {
        if (S.gt(b, c)) S.swp(result, b);
        else if (S.gt(a, c)) S.swp(result, c);
        else S.swp(result, a);
    } else if (S.gt(a, c)) S.swp(result, a);
    else if (S.gt(b, c)) S.swp(result, c);
    else S.swp(result, b);
}

__device__ int sel_unguarded_partition(Sel& S, int first, int last, int pivot) {
    while (true) {
        while (S.gt(first, pivot)) ++first;
        --last;
        while (S.gt(pivot, last)) --last;
        if (!(first < last)) return first;
        S.swp(first, last);
        ++first;
    }
}

__device__ void sel_insertion_sort(Sel& S, int first, int last) {
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
        const unsigned val = S.a[i];
        if ((val >> 24) > (S.a[first] >> 24)) {
            for (int j = i; j > first; --j) S.a[j] = S.a[j - 1];        // move_backward
            S.a[first] = val;
        } else {
            int j = i;                                                   // unguarded linear insert
            while ((val >> 24) > (S.a[j - 1] >> 24)) { S.a[j] = S.a[j - 1]; --j; }
            S.a[j] = val;
        }
    }
}

// returns false when introselect would have fallen back to heap_select (depth limit): not replayed
__device__ bool sel_nth_element(Sel& S, int first, int nth, int last) {
    if (first == last || nth == last) return true;
    int depth = 0;
    for (int n = last - first; n > 1; n >>= 1) ++depth;                  // __lg(n)
    depth *= 2;
    while (last - first > 3) {
        if (depth == 0) return false;
        --depth;
        const int mid = first + (last - first) / 2;
        sel_move_median_to_first(S, first, first + 1, mid, last - 1);
        const int cut = sel_unguarded_partition(S, first + 1, last, first);
        if (cut <= nth) first = cut; else last = cut;
    }
    sel_insertion_sort(S, first, last);
    return true;
}

__global__ void __launch_bounds__(32)
orb_reference_order_kernel(const unsigned long long* __restrict__ sorted, const int* __restrict__ ncand, int cap,
                           const unsigned int* __restrict__ hist, OrbLevels L, int smem_words,
                           unsigned long long* __restrict__ kept, const int* __restrict__ nkept, int max_kp,
                           int* __restrict__ replay_ok) {
    extern __shared__ unsigned sel_buf[];
    const int l = blockIdx.x;
    if (l >= L.nlevels_used) return;
    // segment of level l in the sorted list and in the kept list
    int seg0 = 0, kept0 = 0, cnt = 0, kcnt = 0;
    for (int k = 0; k <= l; ++k) {
        int c = 0;
        for (int s2 = threadIdx.x; s2 < 256; s2 += 32) c += (int)hist[k * 256 + s2];
        c = __reduce_add_sync(0xffffffffu, c);
        int thr_k = 1, acc = 0;
        if (L.quota[k] <= 0) thr_k = 256;
        else for (int s2 = 255; s2 >= 1; --s2) { acc += (int)hist[k * 256 + s2]; if (acc >= L.quota[k]) { thr_k = s2; break; } }
        int kc = 0;
        for (int s2 = thr_k; s2 < 256; ++s2) kc += (int)hist[k * 256 + s2];
        if (k < l) { seg0 += c; kept0 += kc; } else { cnt = c; kcnt = kc; }
    }
    const int quota = L.quota[l];
    if (cnt <= quota || quota <= 0) return;             // retainBest leaves the (row-major) order untouched
    if (seg0 + cnt > min(*ncand, cap) || cnt > smem_words || kept0 + kcnt > max_kp) {
        if (threadIdx.x == 0) atomicExch(replay_ok, 0);
        return;
    }
    for (int i = threadIdx.x; i < cnt; i += 32) sel_buf[i] = ((unsigned)(sorted[seg0 + i] & 0xffu) << 24) | (unsigned)i;
    __syncwarp();
    if (threadIdx.x == 0) {
        Sel S{sel_buf};
        bool ok = sel_nth_element(S, 0, quota - 1, cnt);
        int new_end = quota;
        if (ok) {
            const unsigned amb = S.a[quota - 1] >> 24;
            // std::partition(begin + quota, end, response >= ambiguous)   (bidirectional version)
            int first = quota, last = cnt;
            while (true) {
                while (true) { if (first == last) goto done; else if ((S.a[first] >> 24) >= amb) ++first; else break; }
                --last;
                while (true) { if (first == last) goto done; else if (!((S.a[last] >> 24) >= amb)) --last; else break; }
                S.swp(first, last);
                ++first;
            }
        done:
            new_end = first;
            if (new_end != kcnt) ok = false;
        }
        if (!ok) atomicExch(replay_ok, 0);
        sel_buf[smem_words] = ok ? 1u : 0u;
    }
    __syncwarp();
    if (sel_buf[smem_words]) {
        for (int i = threadIdx.x; i < kcnt; i += 32) kept[kept0 + i] = sorted[seg0 + (sel_buf[i] & 0xffffffu)];
    }
}

// ---------------------------------------------------------------- 7x7 sigma-2 Gaussian blur (float, separable)
__global__ void __launch_bounds__(FTX * FTY)
orb_blur_kernel(const uint8_t* __restrict__ pyr, uint8_t* __restrict__ blur, OrbLevels L) {
    __shared__ uint8_t t[FTY + 6][FTX + 8];
    __shared__ float rowf[FTY + 6][FTX + 1];
    const int l = level_of_tile(L, blockIdx.x);
    const int w = L.w[l], h = L.h[l];
    const int ti = blockIdx.x - L.first_tile[l];
    const int ty = ti / L.tiles_x[l], tx = ti - ty * L.tiles_x[l];
    const int x0 = tx * FTX, y0 = ty * FTY;
    const uint8_t* img = pyr + L.off[l];
    const int tid = threadIdx.y * FTX + threadIdx.x;
    for (int i = tid; i < (FTY + 6) * (FTX + 6); i += FTX * FTY) {
        const int r = i / (FTX + 6), c = i - r * (FTX + 6);
        const int yy = reflect101_multi(y0 + r - 3, h), xx = reflect101_multi(x0 + c - 3, w);
        t[r][c] = img[(size_t)yy * w + xx];
    }
    __syncthreads();
    // getGaussianKernel(7, 2, CV_32F)
    const float k0 = 0.07015932f, k1 = 0.13107488f, k2 = 0.19071282f, k3 = 0.21610594f;
    for (int i = tid; i < (FTY + 6) * FTX; i += FTX * FTY) {
        const int r = i / FTX, c = i - r * FTX;
        // row filter: FMA chain over the 7 taps in tap order
        float s = __fmul_rn((float)t[r][c], k0);
        s = __fmaf_rn((float)t[r][c + 1], k1, s);
        s = __fmaf_rn((float)t[r][c + 2], k2, s);
        s = __fmaf_rn((float)t[r][c + 3], k3, s);
        s = __fmaf_rn((float)t[r][c + 4], k2, s);
        s = __fmaf_rn((float)t[r][c + 5], k1, s);
        s = __fmaf_rn((float)t[r][c + 6], k0, s);
        rowf[r][c] = s;
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x >= w || y >= h) return;
    const int r = threadIdx.y + 3, c = threadIdx.x;
    // symmetric column filter: k3*S0 + k2*(S1+S-1) + k1*(S2+S-2) + k0*(S3+S-3)
    float s = __fmul_rn(rowf[r][c], k3);
    s = __fmaf_rn(__fadd_rn(rowf[r + 1][c], rowf[r - 1][c]), k2, s);
    s = __fmaf_rn(__fadd_rn(rowf[r + 2][c], rowf[r - 2][c]), k1, s);
    s = __fmaf_rn(__fadd_rn(rowf[r + 3][c], rowf[r - 3][c]), k0, s);
    blur[L.off[l] + (size_t)y * w + x] = (uint8_t)min(255, max(0, __float2int_rn(s)));
}

// cv::fastAtan2(y, x) in degrees, every float operation separately rounded
VSTAB_D float fast_atan2_deg(float y, float x) {
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = __fmul_rn(0.9997878412794807f, scale), p3 = __fmul_rn(-0.3258083974640975f, scale);
    const float p5 = __fmul_rn(0.1555786518463281f, scale), p7 = __fmul_rn(-0.04432655554792128f, scale);
    const float eps = (float)2.2204460492503131e-16;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

// ---------------------------------------------------------------- orientation + rBRIEF, one warp per keypoint
__global__ void __launch_bounds__(128)
orb_describe_kernel(const unsigned long long* __restrict__ keys, const int* __restrict__ nkept, int max_kp,
                    const uint8_t* __restrict__ pyr, const uint8_t* __restrict__ blur, OrbLevels L,
                    OrbKeypoint* __restrict__ kps, uint8_t* __restrict__ desc) {
    const int n = min(*nkept, max_kp);
    const int kp = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (kp >= n) return;
    const unsigned long long key = keys[kp];
    const int l = (int)(key >> 56), y = (int)((key >> 36) & 0xfffffu), x = (int)((key >> 16) & 0xfffffu), s = (int)(key & 0xffffu);
    const int w = L.w[l];
    const uint8_t* img = pyr + L.off[l];
    // intensity centroid (ICAngles): row v of the disc per lane, lanes 0..30 <-> v = -15..15
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        const int v = lane - 15;
        const int d = c_umax[v < 0 ? -v : v];
        const uint8_t* row = img + (size_t)(y + v) * w + x;
        int sum = 0;
        for (int u = -d; u <= d; ++u) { const int p = row[u]; m10 += u * p; sum += p; }
        m01 = v * sum;
    }
    m10 = __reduce_add_sync(0xffffffffu, m10);
    m01 = __reduce_add_sync(0xffffffffu, m01);
    const float angle = fast_atan2_deg((float)m01, (float)m10);
    // steered pattern: float products rounded separately (no FMA), cvRound
    const float rad = __fmul_rn(angle, (float)(3.14159265358979323846 / 180.0));
    const float a = (float)cos((double)rad), b = (float)sin((double)rad);
    const uint8_t* bimg = blur + L.off[l] + (size_t)y * w + x;
    unsigned byte = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const signed char* p = c_pattern + (lane * 8 + j) * 4;
        const float x0 = (float)p[0], y0 = (float)p[1], x1 = (float)p[2], y1 = (float)p[3];
        const int ix0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
        const int iy0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
        const int ix1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
        const int iy1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
        const int t0 = bimg[iy0 * w + ix0], t1 = bimg[iy1 * w + ix1];
        byte |= (unsigned)(t0 < t1) << j;
    }
    desc[(size_t)kp * 32 + lane] = (uint8_t)byte;
    if (lane == 0) {
        OrbKeypoint k;
        const float sc = L.scale[l];
        k.x = __fmul_rn((float)x, sc); k.y = __fmul_rn((float)y, sc);
        k.size = __fmul_rn(31.f, sc);
        k.angle = angle;
        k.response = (float)s;
        k.octave = l;
        kps[kp] = k;
    }
}

// ---------------------------------------------------------------- Hamming 2-NN + ratio test
// One warp per query row: lane l scans train rows l, l+32, ... (in index order, so ties keep the lowest index), then
// the 32 partial (best, index, second) triples are merged; `second` is the second order statistic of the distances.
constexpr int HQ = 4;           // query rows (warps) per CTA
__global__ void __launch_bounds__(32 * HQ)
hamming_knn2_kernel(const uint8_t* __restrict__ qdesc, const int* __restrict__ nq_p, int nq_max,
                    const uint8_t* __restrict__ tdesc, const int* __restrict__ nt_p, int nt_max, float ratio,
                    int* __restrict__ best_idx, int* __restrict__ best_d, int* __restrict__ second_d,
                    uint8_t* __restrict__ good) {
    const int nq = min(*nq_p, nq_max), nt = min(*nt_p, nt_max);
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * HQ + (threadIdx.x >> 5);
    if (q >= nq) return;
    const uint4* qp = reinterpret_cast<const uint4*>(qdesc + (size_t)q * 32);
    const uint4 a0 = __ldg(qp), a1 = __ldg(qp + 1);
    constexpr int kNone = 1 << 30;
    int b0 = kNone, b1 = kNone, bi = 0x7fffffff;
    for (int j = lane; j < nt; j += 32) {
        const uint4* tp = reinterpret_cast<const uint4*>(tdesc + (size_t)j * 32);
        const uint4 c0 = __ldg(tp), c1 = __ldg(tp + 1);
        const int d = __popc(a0.x ^ c0.x) + __popc(a0.y ^ c0.y) + __popc(a0.z ^ c0.z) + __popc(a0.w ^ c0.w) +
                      __popc(a1.x ^ c1.x) + __popc(a1.y ^ c1.y) + __popc(a1.z ^ c1.z) + __popc(a1.w ^ c1.w);
        if (d < b0) { b1 = b0; b0 = d; bi = j; }
        else if (d < b1) { b1 = d; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int ob0 = __shfl_xor_sync(0xffffffffu, b0, o), ob1 = __shfl_xor_sync(0xffffffffu, b1, o),
                  obi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob0 < b0 || (ob0 == b0 && obi < bi)) { b1 = min(ob1, b0); b0 = ob0; bi = obi; }   // the other side wins
        else { b1 = min(b1, ob0); }
    }
    if (lane == 0) {
        best_idx[q] = b0 == kNone ? -1 : bi; best_d[q] = b0; second_d[q] = b1;
        // knnMatch returns 2 neighbours only when the train set has >= 2 rows; Lowe ratio in float (:661-662)
        good[q] = (nt >= 2 && (float)b0 < __fmul_rn(ratio, (float)b1)) ? 1 : 0;
    }
}

// gather the matched point pairs (reference order) for the similarity fit
__global__ void __launch_bounds__(256)
match_gather_kernel(const OrbKeypoint* __restrict__ ref_kps, const int* __restrict__ nref_p, int nref_max,
                    const OrbKeypoint* __restrict__ cur_kps, const int* __restrict__ best_idx,
                    const uint8_t* __restrict__ good, float2* __restrict__ ref_pts, float2* __restrict__ cur_pts,
                    uint8_t* __restrict__ status, int* __restrict__ nmatch) {
    // single CTA: stable compaction in reference order
    __shared__ int wsum[8];
    __shared__ int s_base;
    const int nref = min(*nref_p, nref_max);
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int base = 0; base < nref; base += 256) {
        const int i = base + threadIdx.x;
        const bool g = i < nref && good[i];
        const unsigned ball = __ballot_sync(0xffffffffu, g);
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        if (lane == 0) wsum[wid] = __popc(ball);
        __syncthreads();
        int off = s_base;
        for (int k = 0; k < wid; ++k) off += wsum[k];
        if (g) {
            const int pos = off + __popc(ball & ((1u << lane) - 1u));
            const OrbKeypoint a = ref_kps[i], b = cur_kps[best_idx[i]];
            ref_pts[pos] = make_float2(a.x, a.y);
            cur_pts[pos] = make_float2(b.x, b.y);
            status[pos] = 1;
        }
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int k = 0; k < 8; ++k) t += wsum[k]; s_base += t; }
        __syncthreads();
    }
    if (threadIdx.x == 0) *nmatch = s_base;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static size_t al256(size_t v) { return (v + 255) & ~(size_t)255; }

// coefficients of cv::resize(INTER_LINEAR_EXACT) for one axis
static void build_exact_table(int src, int dst, std::vector<int2>& tab) {
    tab.resize(dst);
    const double scale = (double)src / (double)dst;
    for (int d = 0; d < dst; ++d) {
        const double f = (d + 0.5) * scale - 0.5;
        int i; int alpha;
        if (f < 0) { i = 0; alpha = 0; }
        else {
            i = (int)std::floor(f);
            if (i >= src - 1) { i = src - 1; alpha = 0; }
            else alpha = (int)std::floor((f - i) * 256.0 + 0.5);
        }
        tab[d] = make_int2(i, alpha);
    }
}

OrbPlan* orb_plan_create(int w, int h, double size_ratio, int max_keypoints, std::string* err) {
    OrbPlan* P = new OrbPlan();
    P->w = w; P->h = h; P->max_kp = max_keypoints;
    OrbLevels& L = *new OrbLevels();
    P->levels = &L;
    // per-level quota (ORB_Impl::detectAndCompute): float arithmetic as in OpenCV
    const int nfeatures = 2500, nlevels = kOrbLevels;
    const double scaleFactor = (double)1.2f;
    const float factor = (float)(1.0 / scaleFactor);
    float ndesired = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; ++l) {
        L.quota[l] = (int)std::rintf(ndesired);
        sum += L.quota[l];
        ndesired *= factor;
    }
    L.quota[nlevels - 1] = nfeatures - sum > 0 ? nfeatures - sum : 0;
    size_t off = 0;
    int tiles = 0;
    L.nlevels_used = 0;
    for (int l = 0; l < nlevels; ++l) {
        const float sc = (float)std::pow(scaleFactor, (double)l);
        L.scale[l] = sc;
        L.w[l] = (int)std::rintf((float)w / sc);
        L.h[l] = (int)std::rintf((float)h / sc);
        L.off[l] = (unsigned)off;
        off += al256((size_t)L.w[l] * L.h[l]);
        // filterKeypointByRelativeSize (:290-309): keep size < ratio * rows; size = 31 * scale
        const bool used = size_ratio <= 0.0 || (double)(31.f * sc) < size_ratio * (double)h;
        if (used && L.nlevels_used == l) L.nlevels_used = l + 1;
    }
    for (int l = 0; l < nlevels; ++l) {
        L.first_tile[l] = tiles;
        L.tiles_x[l] = (L.w[l] + FTX - 1) / FTX;
        if (l < L.nlevels_used) tiles += L.tiles_x[l] * ((L.h[l] + FTY - 1) / FTY);
    }
    L.first_tile[nlevels] = tiles;
    P->pyr_bytes = off;
    P->ntiles = tiles;
    P->cap = (int)al256((size_t)w * h / 8 + 4096);
    // resize tables
    std::vector<int2> all;
    P->tab_off.resize(nlevels * 2, 0);
    for (int l = 1; l < L.nlevels_used; ++l) {
        std::vector<int2> tx, ty;
        build_exact_table(L.w[l - 1], L.w[l], tx);
        build_exact_table(L.h[l - 1], L.h[l], ty);
        P->tab_off[2 * l] = all.size(); all.insert(all.end(), tx.begin(), tx.end());
        P->tab_off[2 * l + 1] = all.size(); all.insert(all.end(), ty.begin(), ty.end());
    }
    size_t temp = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, temp, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, P->cap);
    P->cub_bytes = temp;
    const size_t bytes = al256(sizeof(int2) * (all.size() + 1)) + 3 * al256(P->pyr_bytes) + 2 * al256(sizeof(unsigned long long) * P->cap) +
                         al256(kOrbLevels * 256 * 4) + al256(256) + al256(temp);
    if (cudaMalloc(&P->mem, bytes) != cudaSuccess) { if (err) *err = "cudaMalloc(ORB workspace) failed"; delete &L; delete P; return nullptr; }
    char* b = (char*)P->mem;
    P->tabs = (int2*)b; b += al256(sizeof(int2) * (all.size() + 1));
    P->pyr = (uint8_t*)b; b += al256(P->pyr_bytes);
    P->blur = (uint8_t*)b; b += al256(P->pyr_bytes);
    P->score = (uint8_t*)b; b += al256(P->pyr_bytes);
    P->cand = (unsigned long long*)b; b += al256(sizeof(unsigned long long) * P->cap);
    P->cand_sorted = (unsigned long long*)b; b += al256(sizeof(unsigned long long) * P->cap);
    P->hist = (unsigned int*)b; b += al256(kOrbLevels * 256 * 4);
    P->counters = (int*)b; b += al256(256);            // [0] ncand, [1] nkept, [4..15] thresholds
    P->cub_temp = b;
    if (!all.empty()) cudaMemcpy(P->tabs, all.data(), sizeof(int2) * all.size(), cudaMemcpyHostToDevice);
    return P;
}

void orb_plan_destroy(OrbPlan* P) {
    if (!P) return;
    if (P->mem) cudaFree(P->mem);
    delete (OrbLevels*)P->levels;
    delete P;
}

int orb_levels_used(const OrbPlan* P) { return ((const OrbLevels*)P->levels)->nlevels_used; }

// gray: device u8 w x h (tight).  Outputs: kps[max_kp], desc[max_kp][32], *count (device int) = number of keypoints.
// reference_order: additionally put the keypoints of every level into OpenCV's retainBest order.
static void launch_orb_body(OrbPlan* P, const uint8_t* gray, OrbKeypoint* kps, uint8_t* desc, int* count, bool reference_order,
                            cudaStream_t st);

void launch_orb(OrbPlan* P, const uint8_t* gray, OrbKeypoint* kps, uint8_t* desc, int* count, bool reference_order,
                cudaStream_t st) {
    // the reference capture (once per mode switch) uploads a host flag: plain launches; everything else replays a graph
    if (reference_order) { launch_orb_body(P, gray, kps, desc, count, true, st); return; }
    const unsigned long long key[5] = {(unsigned long long)gray, (unsigned long long)kps, (unsigned long long)desc,
                                       (unsigned long long)count, (unsigned long long)st};
    run_graphed(P->graphs, key, st, [&] { launch_orb_body(P, gray, kps, desc, count, false, st); });
}

static void launch_orb_body(OrbPlan* P, const uint8_t* gray, OrbKeypoint* kps, uint8_t* desc, int* count, bool reference_order,
                            cudaStream_t st) {
    OrbLevels& L = *(OrbLevels*)P->levels;
    count_launch(8 + (L.nlevels_used - 1) + (reference_order ? 1 : 0));
    cudaMemcpyAsync(P->pyr, gray, (size_t)P->w * P->h, cudaMemcpyDeviceToDevice, st);
    for (int l = 1; l < L.nlevels_used; ++l)
        orb_resize_kernel<<<dim3((L.w[l] + 255) / 256, L.h[l]), 256, 0, st>>>(
            P->pyr + L.off[l - 1], L.w[l - 1], L.h[l - 1], P->pyr + L.off[l], L.w[l], L.h[l], P->tabs + P->tab_off[2 * l],
            P->tabs + P->tab_off[2 * l + 1]);
    cudaMemsetAsync(P->hist, 0, kOrbLevels * 256 * 4, st);
    cudaMemsetAsync(P->counters, 0, 256, st);
    cudaMemsetAsync(count, 0, sizeof(int), st);
    if (P->ntiles > 0) {
        fast_score_kernel<<<P->ntiles, dim3(FTX, FTY), 0, st>>>(P->pyr, P->score, L);
        orb_blur_kernel<<<P->ntiles, dim3(FTX, FTY), 0, st>>>(P->pyr, P->blur, L);
        fast_nms_kernel<<<dim3(148, L.nlevels_used), 256, 0, st>>>(P->score, L, P->hist, P->cand, P->counters, P->cap);
    }
    orb_threshold_kernel<<<1, 32, 0, st>>>(P->hist, L, P->counters + 4);
    orb_tail_kernel<<<(P->cap + 255) / 256, 256, 0, st>>>(P->cand, P->counters, P->cap);
    size_t temp = P->cub_bytes;
    cub::DeviceRadixSort::SortKeys(P->cub_temp, temp, (const unsigned long long*)P->cand, P->cand_sorted, P->cap, 0, 64, st);
    // the kept keys reuse the (now free) unsorted candidate buffer
    orb_compact_kernel<<<1, 1024, 0, st>>>(P->cand_sorted, P->counters, P->cap, P->counters + 4, P->cand, count, P->max_kp);
    if (reference_order && L.nlevels_used > 0) {
        const int words = 50 * 1024;
        static PerDeviceOnce once;
        once.run([&] { cudaFuncSetAttribute(orb_reference_order_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (words + 1) * 4); });
        int one = 1;
        cudaMemcpyAsync(P->counters + 2, &one, sizeof(int), cudaMemcpyHostToDevice, st);
        orb_reference_order_kernel<<<L.nlevels_used, 32, (words + 1) * 4, st>>>(P->cand_sorted, P->counters, P->cap, P->hist, L,
                                                                                words, P->cand, count, P->max_kp, P->counters + 2);
    }
    orb_describe_kernel<<<(P->max_kp + 3) / 4, 128, 0, st>>>(P->cand, count, P->max_kp, P->pyr, P->blur, L, kps, desc);
}

void launch_hamming_match(const uint8_t* ref_desc, const int* nref, const OrbKeypoint* ref_kps, const uint8_t* cur_desc,
                          const int* ncur, const OrbKeypoint* cur_kps, int max_kp, float ratio, int* best_idx, int* best_d,
                          int* second_d, uint8_t* good, float2* ref_pts, float2* cur_pts, uint8_t* status, int* nmatch,
                          cudaStream_t st) {
    count_launch(2);
    hamming_knn2_kernel<<<(max_kp + HQ - 1) / HQ, 32 * HQ, 0, st>>>(ref_desc, nref, max_kp, cur_desc, ncur, max_kp, ratio, best_idx,
                                                                 best_d, second_d, good);
    match_gather_kernel<<<1, 256, 0, st>>>(ref_kps, nref, max_kp, cur_kps, best_idx, good, ref_pts, cur_pts, status, nmatch);
}

}  // namespace vstabk
