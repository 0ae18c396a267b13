// K3: Shi-Tomasi corners = cv::goodFeaturesToTrack(gray, 1300, 0.01, minDist, noMask, 3, 3, false, 0.04)
// (/root/reference/src/stabilizer.cpp:931-980; the reference calls it twice with identical
// output, :949 and :961 -- it runs once here).  SURVEY A.5:
//   pass A  cornerMinEigenVal: Sobel3 (OpenCV's FMA order) -> products -> 3x3 box (f64 sums)
//           -> min eigenvalue (f32), global max, and -- in the same kernel -- every 3x3 local
//           maximum as a 64-bit sort key (value desc, address desc).  OpenCV thresholds at
//           0.01*max *before* its dilate test, but a pixel above the threshold can only lose to
//           a neighbour that is itself above it, so "local maximum" does not depend on the
//           threshold: the cut is applied after the sort, where it is a prefix of the list.
//           The min-eigenvalue map is not written to HBM (only when a tap asks for it).
//   sort    segmented radix sort of the candidate keys (one segment per frame);
//   pass C  greedy min-distance suppression in sorted order on a cell grid held in shared
//           memory, one warp per frame, 32 candidates per step with ballot/shuffle conflict
//           resolution -- the accepted set and its order equal the sequential OpenCV loop.
#include <cub/device/device_segmented_radix_sort.cuh>
#include "kernels.h"

namespace vstabk {
namespace {

constexpr int ETX = 32, ETY = 16;                 // candidate tile
constexpr int EW = ETX + 2, EH = ETY + 2;         // eig footprint (halo 1)
constexpr int PW = ETX + 4, PH = ETY + 4;         // product footprint (halo 2)
constexpr int GW = ETX + 6, GH = ETY + 6;         // gray footprint (halo 3)
constexpr unsigned short kEmpty16 = 0xffffu;
constexpr int kCellCap = 4;                       // u16 slots per cell ({dy,dx} relative to the cell)
constexpr int kGreedySmemMax = 200 * 1024;

__global__ void __launch_bounds__(ETX * ETY)
eig_kernel(const uint8_t* __restrict__ gray, size_t gray_stride, int w, int h,
           float* __restrict__ eig_out, unsigned int* __restrict__ maxbits,
           unsigned long long* __restrict__ keys, int* __restrict__ seg_end, int cap, double quality) {
    __shared__ float g[GH][GW + 1];
    __shared__ float pxx[PH][PW + 1], pxy[PH][PW + 1], pyy[PH][PW + 1];
    __shared__ float se[EH][EW + 1];
    __shared__ float wmax[ETX * ETY / 32];
    const int frame = blockIdx.z;
    const uint8_t* src = gray + (size_t)frame * gray_stride;
    const int x0 = blockIdx.x * ETX, y0 = blockIdx.y * ETY;
    const int tid = threadIdx.y * ETX + threadIdx.x;

    for (int i = tid; i < GH * GW; i += ETX * ETY) {
        const int r = i / GW, c = i - r * GW;
        int yy = min(max(y0 - 3 + r, -(h - 1)), 2 * h - 2);
        int xx = min(max(x0 - 3 + c, -(w - 1)), 2 * w - 2);
        yy = reflect101(yy, h);
        xx = reflect101(xx, w);
        g[r][c] = (float)src[(size_t)yy * w + xx];
    }
    __syncthreads();

    // scale = 1 / (2^(ksize-1) * blockSize * 255) ; k1 = float(scale), k0 = float(2*scale)
    const float k1 = (float)(1.0 / (4.0 * 3.0 * 255.0));
    const float k0 = (float)(2.0 / (4.0 * 3.0 * 255.0));
    for (int i = tid; i < PH * PW; i += ETX * ETY) {
        const int r = i / PW, c = i - r * PW;
        // image position of this product sample, reflected into the image (box BORDER_REFLECT_101)
        int py = min(max(y0 - 2 + r, -(h - 1)), 2 * h - 2);
        int px = min(max(x0 - 2 + c, -(w - 1)), 2 * w - 2);
        py = reflect101(py, h);
        px = reflect101(px, w);
        // tile coordinates of that position (clamped: out-of-tile only for unused samples)
        const int tr = min(max(py - (y0 - 3), 1), GH - 2);
        const int tc = min(max(px - (x0 - 3), 1), GW - 2);
        // Dx = fma(S[y-1] + S[y+1], k1, S[y]*k0),  S[y] = p[y][x+1] - p[y][x-1]
        const float sm = g[tr - 1][tc + 1] - g[tr - 1][tc - 1];
        const float s0 = g[tr][tc + 1] - g[tr][tc - 1];
        const float sp = g[tr + 1][tc + 1] - g[tr + 1][tc - 1];
        const float dx = __fmaf_rn(__fadd_rn(sm, sp), k1, __fmul_rn(s0, k0));
        // Dy = R[y+1] - R[y-1],  R[y] = fma(p[x+1], k1, fma(p[x], k0, p[x-1]*k1))
        // OpenCV's row filter runs its FMA vector body over the first floor(w/32)*32 columns and a
        // scalar, un-contracted tail over the rest ([probe] cv2 4.13.0 on AVX-512 hosts; all
        // production widths 640/1280/1920/3840 are multiples of 32 and never reach the tail).
        float rm, rp;
        if (px < (w & ~31)) {
            rm = __fmaf_rn(g[tr - 1][tc + 1], k1, __fmaf_rn(g[tr - 1][tc], k0, __fmul_rn(g[tr - 1][tc - 1], k1)));
            rp = __fmaf_rn(g[tr + 1][tc + 1], k1, __fmaf_rn(g[tr + 1][tc], k0, __fmul_rn(g[tr + 1][tc - 1], k1)));
        } else {
            rm = __fadd_rn(__fadd_rn(__fmul_rn(g[tr - 1][tc - 1], k1), __fmul_rn(g[tr - 1][tc], k0)),
                           __fmul_rn(g[tr - 1][tc + 1], k1));
            rp = __fadd_rn(__fadd_rn(__fmul_rn(g[tr + 1][tc - 1], k1), __fmul_rn(g[tr + 1][tc], k0)),
                           __fmul_rn(g[tr + 1][tc + 1], k1));
        }
        const float dy = __fsub_rn(rp, rm);
        pxx[r][c] = __fmul_rn(dx, dx);
        pxy[r][c] = __fmul_rn(dx, dy);
        pyy[r][c] = __fmul_rn(dy, dy);
    }
    __syncthreads();

    // min eigenvalue on the (ETX+2) x (ETY+2) footprint; positions outside the image stay 0
    // (never compared: candidates exclude the 1-px image border)
    float m = 0.f;
    for (int i = tid; i < EH * EW; i += ETX * ETY) {
        const int r = i / EW, c = i - r * EW;           // eig (r,c) <-> image (y0-1+r, x0-1+c); products r..r+2
        const int x = x0 - 1 + c, y = y0 - 1 + r;
        float e = 0.f;
        if (x >= 0 && x < w && y >= 0 && y < h) {
            double sxx = 0.0, sxy = 0.0, syy = 0.0;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
                const double rxx = __dadd_rn(__dadd_rn((double)pxx[r + dy][c], (double)pxx[r + dy][c + 1]),
                                             (double)pxx[r + dy][c + 2]);
                const double rxy = __dadd_rn(__dadd_rn((double)pxy[r + dy][c], (double)pxy[r + dy][c + 1]),
                                             (double)pxy[r + dy][c + 2]);
                const double ryy = __dadd_rn(__dadd_rn((double)pyy[r + dy][c], (double)pyy[r + dy][c + 1]),
                                             (double)pyy[r + dy][c + 2]);
                sxx = __dadd_rn(sxx, rxx);
                sxy = __dadd_rn(sxy, rxy);
                syy = __dadd_rn(syy, ryy);
            }
            const float a = __fmul_rn((float)sxx, 0.5f);
            const float b = (float)sxy;
            const float cc = __fmul_rn((float)syy, 0.5f);
            const float d = __fsub_rn(a, cc);
            const float rad = __fsqrt_rn(__fadd_rn(__fmul_rn(d, d), __fmul_rn(b, b)));
            e = __fsub_rn(__fadd_rn(a, cc), rad);
            const bool own = r >= 1 && r <= ETY && c >= 1 && c <= ETX;   // pixel of this tile (not halo)
            if (own) {
                m = fmaxf(m, e);
                if (eig_out) eig_out[(size_t)frame * w * h + (size_t)y * w + x] = e;
            }
        }
        se[r][c] = e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((tid & 31) == 0) wmax[tid >> 5] = m;
    __syncthreads();
    if (tid < 32) {
        float v = tid < ETX * ETY / 32 ? wmax[tid] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
        if (tid == 0 && v > 0.f) atomicMax(maxbits + frame, __float_as_uint(v));
    }

    // 3x3 local maxima (ties kept, like eig == dilate(eig)) inside the 1-px image border
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    const int r = threadIdx.y + 1, c = threadIdx.x + 1;
    const float v = se[r][c];
    bool is_cand = x >= 1 && x < w - 1 && y >= 1 && y < h - 1 && v > 0.f;
    if (is_cand) {
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx)
                if (se[r + dy][c + dx] > v) is_cand = false;
    }
    // conservative early cut: the running maximum only grows, so anything at or below
    // 0.01 * (maximum so far) is certainly below the final threshold
    if (is_cand) {
        const float lb = __uint_as_float(*(volatile unsigned int*)(maxbits + frame));
        if (v <= (float)((double)lb * quality) * 0.999f) is_cand = false;
    }
    const unsigned ball = __ballot_sync(0xffffffffu, is_cand);
    if (ball) {
        const int lane = tid & 31;
        int pos0 = 0;
        if (lane == 0) pos0 = atomicAdd(seg_end + frame, __popc(ball));
        pos0 = __shfl_sync(0xffffffffu, pos0, 0);
        if (is_cand) {
            const int pos = pos0 + __popc(ball & ((1u << lane) - 1u));
            if (pos < (frame + 1) * cap)
                keys[pos] = ((unsigned long long)__float_as_uint(v) << 32) | (unsigned long long)(unsigned)(y * w + x);
        }
    }
}

__global__ void gftt_reset_kernel(unsigned int* maxbits, int* seg_begin, int* seg_end, int cap, int nframes) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f < nframes) {
        maxbits[f] = 0u;
        seg_begin[f] = f * cap;
        seg_end[f] = f * cap;
    }
}

__global__ void clamp_segments_kernel(int* seg_end, int cap, int nframes) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f < nframes) seg_end[f] = min(seg_end[f], (f + 1) * cap);
}

// Greedy minimum-distance selection, one warp per frame.  The cell grid (4 u16 slots per
// cell, {dy,dx} relative to the cell origin) lives in shared memory when it fits, else in
// the global scratch `grid_glob`.
__global__ void __launch_bounds__(32)
greedy_kernel(const unsigned long long* __restrict__ keys, const int* __restrict__ seg_begin,
              const int* __restrict__ seg_end, const unsigned int* __restrict__ maxbits, double quality,
              int w, int min_distance, int cell, int grid_w, int grid_h,
              unsigned short* grid_glob, int use_smem, int max_corners,
              float2* __restrict__ pts, int* __restrict__ counts) {
    extern __shared__ unsigned short grid_sm[];
    const int frame = blockIdx.x;
    const int lane = threadIdx.x;
    const unsigned long long* k = keys + seg_begin[frame];
    const int n = seg_end[frame] - seg_begin[frame];
    const int ncells = grid_w * grid_h;
    unsigned short* gfr = use_smem ? grid_sm : grid_glob + (size_t)frame * ncells * kCellCap;
    {
        uint2* g2 = reinterpret_cast<uint2*>(gfr);
        for (int i = lane; i < ncells; i += 32) g2[i] = make_uint2(0xffffffffu, 0xffffffffu);
    }
    __syncwarp();
    float2* out = pts + (size_t)frame * kMaxCorners;
    const int md2 = min_distance * min_distance;
    const float maxv = __uint_as_float(maxbits[frame]);
    const float thr = (float)((double)maxv * quality);       // cv::threshold takes float(thresh)
    int accepted = 0;

    for (int base = 0; base < n && accepted < max_corners; base += 32) {
        const int rank = base + lane;
        bool valid = rank < n;
        int x = 0, y = 0;
        if (valid) {
            const unsigned long long key = k[rank];
            valid = __uint_as_float((unsigned)(key >> 32)) > thr;     // THRESH_TOZERO cut: a prefix of the sorted list
            const unsigned idx = (unsigned)(key & 0xffffffffull);
            y = idx / w;
            x = idx - y * w;
        }
        if (!__any_sync(0xffffffffu, valid)) break;
        bool ok = valid;
        int cx = 0, cy = 0;
        if (min_distance >= 1) {
            cx = x / cell;
            cy = y / cell;
            if (valid) {
                const int x1 = max(cx - 1, 0), x2 = min(cx + 1, grid_w - 1);
                const int y1 = max(cy - 1, 0), y2 = min(cy + 1, grid_h - 1);
                for (int gy = y1; gy <= y2 && ok; ++gy)
                    for (int gx = x1; gx <= x2 && ok; ++gx) {
                        const uint2 sl = *reinterpret_cast<const uint2*>(gfr + ((size_t)gy * grid_w + gx) * kCellCap);
                        const unsigned s[4] = {sl.x & 0xffffu, sl.x >> 16, sl.y & 0xffffu, sl.y >> 16};
#pragma unroll
                        for (int q = 0; q < kCellCap; ++q) {
                            if (s[q] != kEmpty16) {
                                const int dx = x - (gx * cell + (int)(s[q] & 0xffu));
                                const int dy = y - (gy * cell + (int)(s[q] >> 8));
                                if (dx * dx + dy * dy < md2) ok = false;
                            }
                        }
                    }
            }
            // resolve conflicts inside the batch in rank order
            unsigned pending = __ballot_sync(0xffffffffu, ok);
            while (pending) {
                const int kk = __ffs(pending) - 1;             // lowest-rank still-ok lane: accepted
                const int bx = __shfl_sync(0xffffffffu, x, kk);
                const int by = __shfl_sync(0xffffffffu, y, kk);
                if (lane > kk && ok) {
                    const int dx = x - bx, dy = y - by;
                    if (dx * dx + dy * dy < md2) ok = false;
                }
                pending = __ballot_sync(0xffffffffu, ok) & ~((2u << kk) - 1u);
            }
        }
        const unsigned acc = __ballot_sync(0xffffffffu, ok);
        const int my = accepted + __popc(acc & ((1u << lane) - 1u));
        if (ok && my < max_corners) {
            out[my] = make_float2((float)x, (float)y);
            if (min_distance >= 1) {
                unsigned short* cellp = gfr + ((size_t)cy * grid_w + cx) * kCellCap;
                const unsigned short val = (unsigned short)(((y - cy * cell) << 8) | (x - cx * cell));
                for (int q = 0; q < kCellCap; ++q)
                    if (atomicCAS(cellp + q, kEmpty16, val) == kEmpty16) break;
            }
        }
        accepted += __popc(acc);
        __syncwarp();
    }
    if (lane == 0) counts[frame] = min(accepted, max_corners);
}

}  // namespace

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

size_t gftt_workspace_bytes(int w, int h, int min_distance, int max_frames, GfttWorkspace* L) {
    GfttWorkspace ws{};
    ws.max_frames = max_frames;
    ws.cap = (int)align_up((size_t)w * h / 4 + 1024, 256);
    ws.cell = min_distance >= 1 ? min_distance : 1;
    ws.grid_w = (w + ws.cell - 1) / ws.cell;
    ws.grid_h = (h + ws.cell - 1) / ws.cell;
    ws.ncells = ws.grid_w * ws.grid_h;
    ws.grid_in_smem = (size_t)ws.ncells * kCellCap * sizeof(unsigned short) <= (size_t)kGreedySmemMax ? 1 : 0;
    size_t temp = 0;
    cub::DeviceSegmentedRadixSort::SortKeysDescending(
        nullptr, temp, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
        (int)((size_t)ws.cap * max_frames), max_frames, (const int*)nullptr, (const int*)nullptr, 0, 64);
    ws.cub_temp_bytes = temp;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    // offsets are stored in the pointer fields and rebased by gftt_bind_workspace
    ws.eig = (float*)take((size_t)w * h * sizeof(float));            // one frame: parity tap only
    ws.maxbits = (unsigned int*)take((size_t)max_frames * sizeof(unsigned int));
    ws.keys = (unsigned long long*)take((size_t)max_frames * ws.cap * sizeof(unsigned long long));
    ws.keys_alt = (unsigned long long*)take((size_t)max_frames * ws.cap * sizeof(unsigned long long));
    ws.seg_begin = (int*)take((size_t)max_frames * sizeof(int));
    ws.seg_end = (int*)take((size_t)max_frames * sizeof(int));
    ws.grid = (unsigned short*)take(ws.grid_in_smem ? 16 : (size_t)max_frames * ws.ncells * kCellCap * sizeof(unsigned short));
    ws.cub_temp = (void*)take(temp);
    *L = ws;
    return off;
}

void gftt_bind_workspace(void* base, GfttWorkspace* ws) {
    char* b = (char*)base;
    ws->eig = (float*)(b + (size_t)ws->eig);
    ws->maxbits = (unsigned int*)(b + (size_t)ws->maxbits);
    ws->keys = (unsigned long long*)(b + (size_t)ws->keys);
    ws->keys_alt = (unsigned long long*)(b + (size_t)ws->keys_alt);
    ws->seg_begin = (int*)(b + (size_t)ws->seg_begin);
    ws->seg_end = (int*)(b + (size_t)ws->seg_end);
    ws->grid = (unsigned short*)(b + (size_t)ws->grid);
    ws->cub_temp = (void*)(b + (size_t)ws->cub_temp);
}

void launch_gftt(const uint8_t* gray, size_t gray_frame_stride, int w, int h, int nframes,
                 double quality, int min_distance, int max_corners, GfttWorkspace& ws,
                 float2* pts, int* counts, float* eig_out, cudaStream_t st) {
    if (nframes <= 0) return;
    if (max_corners > kMaxCorners) max_corners = kMaxCorners;
    count_launch(4);   // reset, eig+candidates, clamp, greedy (+ cub's radix-sort passes, not counted)
    gftt_reset_kernel<<<(nframes + 127) / 128, 128, 0, st>>>(ws.maxbits, ws.seg_begin, ws.seg_end, ws.cap, nframes);
    {
        dim3 grid((w + ETX - 1) / ETX, (h + ETY - 1) / ETY, nframes);
        dim3 block(ETX, ETY);
        eig_kernel<<<grid, block, 0, st>>>(gray, gray_frame_stride, w, h, eig_out, ws.maxbits, ws.keys, ws.seg_end, ws.cap, quality);
    }
    clamp_segments_kernel<<<(nframes + 127) / 128, 128, 0, st>>>(ws.seg_end, ws.cap, nframes);
    size_t temp = ws.cub_temp_bytes;
    cub::DeviceSegmentedRadixSort::SortKeysDescending(
        ws.cub_temp, temp, (const unsigned long long*)ws.keys, ws.keys_alt, (int)((size_t)ws.cap * nframes),
        nframes, (const int*)ws.seg_begin, (const int*)ws.seg_end, 0, 63, st);
    const size_t smem = ws.grid_in_smem ? (size_t)ws.ncells * kCellCap * sizeof(unsigned short) : 0;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(greedy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGreedySmemMax);
        attr_set = true;
    }
    greedy_kernel<<<nframes, 32, smem, st>>>(ws.keys_alt, ws.seg_begin, ws.seg_end, ws.maxbits, quality, w, min_distance,
                                             ws.cell, ws.grid_w, ws.grid_h, ws.grid, ws.grid_in_smem, max_corners,
                                             pts, counts);
}

}  // namespace vstabk
