// K1 ingest: one pass over a full-resolution BGR frame producing
//   (a) the working-resolution gray image  = cvtColor(resize(frame, INTER_LINEAR), BGR2GRAY)
//       -- /root/reference/src/stabilizer.cpp:1169-1175 (bit-exact), and
//   (b) the per-channel byte sums that cv::mean needs for the border colour
//       -- /root/reference/src/stabilizer.cpp:1309 (the reference re-reads the frame for this).
// HBM-bound: every source byte is read exactly once with 16-byte loads; the resize taps
// of a band come from the rows the same CTA streams (L1/L2 hits).
#include <cmath>
#include "kernels.h"

namespace vstabk {

// Per-axis source index and Q11 coefficients of cv::resize(INTER_LINEAR):
// fx = float((d+0.5)*scale - 0.5); s = floor(fx); fx -= s; clamp; c1 = rint(fx*2048), c0 = rint((1-fx)*2048)
void build_ingest_tables(int src, int dst, int mode, int4* tab) {
    for (int d = 0; d < dst; ++d) {
        int4 e;
        if (mode == 0) {
            e = make_int4(d, d, 2048, 0);
        } else if (mode == 1) {
            e = make_int4(2 * d, 2 * d + 1, 1024, 1024);
        } else {
            double scale = (double)src / (double)dst;
            float fx = (float)((d + 0.5) * scale - 0.5);
            int s = (int)floorf(fx);
            fx -= (float)s;
            if (s < 0) { s = 0; fx = 0.f; }
            if (s >= src - 1) { s = src - 1; fx = 0.f; }
            int c1 = (int)rintf(fx * 2048.f);
            int c0 = (int)rintf((1.f - fx) * 2048.f);
            int s1 = s + 1 < src ? s + 1 : src - 1;
            e = make_int4(s, s1, c0, c1);
        }
        tab[d] = e;
    }
}

namespace {

constexpr int kIngestThreads = 256;

__device__ __forceinline__ uint4 ld_stream16(const uint8_t* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::evict_last.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// Adds the B,G,R byte sums of 48 bytes (16 pixels, phase-aligned) held in 12 words.
__device__ __forceinline__ void sum48(const uint4& a, const uint4& b, const uint4& c,
                                      unsigned& sb, unsigned& sg, unsigned& sr) {
    // word k holds bytes 4k..4k+3; channel of byte i is i % 3 (0=B,1=G,2=R)
    // k%3==0: B G R B   k%3==1: G R B G   k%3==2: R B G R
    const unsigned m0 = 0x01000001u, m1 = 0x00000100u, m2 = 0x00010000u;
#define ACC(w0, w1, w2)                                                     \
    sb = __dp4a(w0, m0, sb); sg = __dp4a(w0, m1, sg); sr = __dp4a(w0, m2, sr); \
    sg = __dp4a(w1, m0, sg); sr = __dp4a(w1, m1, sr); sb = __dp4a(w1, m2, sb); \
    sr = __dp4a(w2, m0, sr); sb = __dp4a(w2, m1, sb); sg = __dp4a(w2, m2, sg);
    ACC(a.x, a.y, a.z)
    ACC(a.w, b.x, b.y)
    ACC(b.z, b.w, c.x)
    ACC(c.y, c.z, c.w)
#undef ACC
}

__global__ void __launch_bounds__(kIngestThreads)
ingest_kernel(IngestPlan plan, const uint8_t* __restrict__ frames, size_t pitch, size_t frame_stride,
              uint8_t* __restrict__ gray, size_t gray_frame_stride,
              unsigned long long* __restrict__ sums) {
    const int band = blockIdx.x;
    const int frame = blockIdx.y;
    const uint8_t* src = frames + (size_t)frame * frame_stride;
    uint8_t* dst = gray + (size_t)frame * gray_frame_stride;

    const int dy0 = band * plan.rows_per_band;
    const int dy1 = min(plan.dst_h, dy0 + plan.rows_per_band);
    // source rows owned by this band for the channel sums: an exact partition of [0, src_h)
    const int sy0 = (int)(((long long)dy0 * plan.src_h) / plan.dst_h);
    const int sy1 = (dy1 >= plan.dst_h) ? plan.src_h : (int)(((long long)dy1 * plan.src_h) / plan.dst_h);

    // ---- (b) channel sums over the owned source rows, 48-byte groups per thread ----------
    unsigned sb = 0, sg = 0, sr = 0;
    const int row_bytes = plan.src_w * 3;
    const int groups_per_row = row_bytes / 48;
    const int ngroups = groups_per_row * (sy1 - sy0);
    for (int g = threadIdx.x; g < ngroups; g += kIngestThreads) {
        const int r = g / groups_per_row;
        const int c = g - r * groups_per_row;
        const uint8_t* p = src + (size_t)(sy0 + r) * pitch + (size_t)c * 48;
        const uint4 a = ld_stream16(p);
        const uint4 b = ld_stream16(p + 16);
        const uint4 cc = ld_stream16(p + 32);
        sum48(a, b, cc, sb, sg, sr);
    }
    // tail pixels of each row (row_bytes not a multiple of 48)
    const int tail0 = groups_per_row * 16;                 // first tail pixel
    const int tailn = plan.src_w - tail0;
    if (tailn > 0) {
        const int nt = tailn * (sy1 - sy0);
        for (int t = threadIdx.x; t < nt; t += kIngestThreads) {
            const int r = t / tailn;
            const int x = tail0 + (t - r * tailn);
            const uint8_t* p = src + (size_t)(sy0 + r) * pitch + (size_t)x * 3;
            sb += p[0]; sg += p[1]; sr += p[2];
        }
    }

    // ---- (a) gray pixels of this band ---------------------------------------------------
    const int npx = (dy1 - dy0) * plan.dst_w;
    for (int i = threadIdx.x; i < npx; i += kIngestThreads) {
        const int yy = i / plan.dst_w;
        const int x = i - yy * plan.dst_w;
        const int y = dy0 + yy;
        int b, g, r;
        if (plan.mode == 0) {
            const uint8_t* p = src + (size_t)y * pitch + (size_t)x * 3;
            b = p[0]; g = p[1]; r = p[2];
        } else if (plan.mode == 1) {
            const uint8_t* p0 = src + (size_t)(2 * y) * pitch + (size_t)(2 * x) * 3;
            const uint8_t* p1 = p0 + pitch;
            b = (p0[0] + p0[3] + p1[0] + p1[3] + 2) >> 2;
            g = (p0[1] + p0[4] + p1[1] + p1[4] + 2) >> 2;
            r = (p0[2] + p0[5] + p1[2] + p1[5] + 2) >> 2;
        } else {
            const int4 tx = __ldg(plan.xtab + x);
            const int4 ty = __ldg(plan.ytab + y);
            const uint8_t* r0 = src + (size_t)ty.x * pitch;
            const uint8_t* r1 = src + (size_t)ty.y * pitch;
            const uint8_t* a0 = r0 + (size_t)tx.x * 3;
            const uint8_t* a1 = r0 + (size_t)tx.y * 3;
            const uint8_t* b0 = r1 + (size_t)tx.x * 3;
            const uint8_t* b1 = r1 + (size_t)tx.y * 3;
            int v[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int S0 = a0[c] * tx.z + a1[c] * tx.w;
                const int S1 = b0[c] * tx.z + b1[c] * tx.w;
                v[c] = ((((ty.z * (S0 >> 4)) >> 16) + ((ty.w * (S1 >> 4)) >> 16) + 2) >> 2);
            }
            b = v[0]; g = v[1]; r = v[2];
        }
        dst[(size_t)y * plan.dst_w + x] = (uint8_t)luma_q15(b, g, r);
    }

    // ---- block reduction of the sums, one 64-bit atomic per channel per CTA ---------------
    sb = warp_sum(sb); sg = warp_sum(sg); sr = warp_sum(sr);
    __shared__ unsigned red[3][kIngestThreads / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { red[0][wid] = sb; red[1][wid] = sg; red[2][wid] = sr; }
    __syncthreads();
    if (threadIdx.x < 3) {
        unsigned long long t = 0;
#pragma unroll
        for (int w = 0; w < kIngestThreads / 32; ++w) t += red[threadIdx.x][w];
        atomicAdd(sums + (size_t)frame * 3 + threadIdx.x, t);
    }
}

}  // namespace

void launch_ingest(const IngestPlan& plan, const uint8_t* frames, size_t pitch, size_t frame_stride,
                   int nframes, uint8_t* gray, size_t gray_frame_stride,
                   unsigned long long* sums, cudaStream_t st) {
    if (nframes <= 0) return;
    dim3 grid(plan.nbands, nframes);
    count_launch(1);
    ingest_kernel<<<grid, kIngestThreads, 0, st>>>(plan, frames, pitch, frame_stride, gray,
                                                   gray_frame_stride, sums);
}

}  // namespace vstabk
