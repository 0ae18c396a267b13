// K5: Stabilizer::estimateMotion  (/root/reference/src/stabilizer.cpp:211-275)
//   keep status==1 pairs (:203-208)  ->  < 10 points: identity (:215)
//   cv::estimateAffinePartial2D(prev, cur, noArray(), RANSAC)            (:224-225)
//       OpenCV's RANSAC loop restated exactly (same RNG stream, same 2-point models, same f32
//       error test, same adaptive iteration bound => bit-identical consensus mask), each
//       iteration scored by the whole CTA; then the closed-form least-squares similarity on
//       the consensus set -- the fixed point of OpenCV's LM refinement (<= 3e-11 px)
//   NaN guard -> identity (:241); embed 2x3 in 3x3 (:244-251)
//   decompose about the working-image centre, force s = 1, recompose (:261-272).
// One CTA per frame pair.
#include "homography.cuh"
#include "kernels.h"

namespace vstabk {
namespace {

constexpr int kFitThreads = 256;

VSTAB_D unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

template <typename T>
VSTAB_D T block_sum(T v, T* scratch /* >= 8 */) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    T t = 0;
#pragma unroll
    for (int w = 0; w < kFitThreads / 32; ++w) t += scratch[w];
    return t;
}

template <int CAP>
__global__ void __launch_bounds__(kFitThreads)
fit_kernel(const float2* __restrict__ prev_pts, const float2* __restrict__ next_pts,
           const uint8_t* __restrict__ status, const int* __restrict__ counts,
           double thresh, double cx, double cy, double* __restrict__ T, double* __restrict__ M,
           int* __restrict__ fit_counts, long frame_id0) {
    extern __shared__ float fit_smem[];
    float* spx = fit_smem; float* spy = spx + CAP; float* sqx = spy + CAP; float* sqy = sqx + CAP;
    __shared__ int wcount[kFitThreads / 32];
    __shared__ int wcount2[2][kFitThreads / 32];
    __shared__ double dscratch[kFitThreads / 32];
    __shared__ int s_m;

    const int frame = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n = min(counts[frame], CAP);
    const float2* P = prev_pts + (size_t)frame * CAP;
    const float2* Q = next_pts + (size_t)frame * CAP;
    const uint8_t* S = status + (size_t)frame * CAP;

    // ---- stable compaction of the tracked pairs -----------------------------------------
    constexpr int kPer = (CAP + kFitThreads - 1) / kFitThreads;           // consecutive items per thread
    const int i0 = tid * kPer;
    int mine = 0;
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
        const int i = i0 + j;
        if (i < n && S[i] == 1) ++mine;
    }
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wcount[wid] = incl;
    __syncthreads();
    int wbase = 0;
    for (int w = 0; w < wid; ++w) wbase += wcount[w];
    int pos = wbase + incl - mine;
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
        const int i = i0 + j;
        if (i < n && S[i] == 1) {
            const float2 p = P[i], q = Q[i];
            spx[pos] = p.x; spy[pos] = p.y; sqx[pos] = q.x; sqy[pos] = q.y;
            ++pos;
        }
    }
    if (tid == kFitThreads - 1) s_m = wbase + incl;
    __syncthreads();
    const int m = s_m;

    double* Tout = T + (size_t)frame * 9;
    if (m < kMinPointsForMotion) {                      // stabilizer.cpp:215
        if (tid == 0) {
            eye3(Tout);
            if (M) { double* Mo = M + (size_t)frame * 6; Mo[0] = 1; Mo[1] = 0; Mo[2] = 0; Mo[3] = 0; Mo[4] = 1; Mo[5] = 0; }
            if (fit_counts) { fit_counts[frame * 2] = m; fit_counts[frame * 2 + 1] = 0; }
        }
        return;
    }

    // ---- RANSACPointSetRegistrator::run, restated exactly (oracle/cv_restate.py::ransac_similarity):
    // cv::RNG seeded with (uint64)-1, 2 distinct uniform indices per iteration, 2-point model in
    // f64, errors in f32 with the model cast to float, "goodCount > max(maxGood, 1)" update and
    // the adaptive iteration bound.  The loop is sequential by definition (each draw depends on
    // the RNG state, the bound on the best count so far); every iteration is evaluated by the
    // whole CTA (thread t owns points t, t+256, ...), so the consensus mask is bit-identical
    // to OpenCV's and the loop typically ends after 2-5 iterations.
    const float thr2 = (float)(thresh * thresh);
    unsigned long long rng = 0xFFFFFFFFFFFFFFFFull;
    auto rng_uniform = [&](unsigned n_) -> int {
        rng = (unsigned long long)(unsigned)rng * 4164903690ull + (unsigned)(rng >> 32);
        return (int)((unsigned)rng % n_);
    };
    constexpr int kOwn = (CAP + kFitThreads - 1) / kFitThreads;           // points per thread (mask bits)
    unsigned best_bits = 0;
    int max_good = 0;
    int niters = 2000;
    for (int iter = 0; iter < niters; ++iter) {
        const int i0 = rng_uniform((unsigned)m);
        int i1;
        do { i1 = rng_uniform((unsigned)m); } while (i1 == i0);
        // AffinePartial2DEstimatorCallback::runKernel
        const double x1 = spx[i0], y1 = spy[i0], x2 = spx[i1], y2 = spy[i1];
        const double X1 = sqx[i0], Y1 = sqy[i0], X2 = sqx[i1], Y2 = sqy[i1];
        const double d = 1. / ((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2));
        const double S0 = d * ((X1 - X2) * (x1 - x2) + (Y1 - Y2) * (y1 - y2));
        const double S1 = d * ((Y1 - Y2) * (x1 - x2) - (X1 - X2) * (y1 - y2));
        const double S2 = d * ((Y1 - Y2) * (x1 * y2 - x2 * y1) - (X1 * y2 - X2 * y1) * (y1 - y2) - (X1 * x2 - X2 * x1) * (x1 - x2));
        const double S3 = d * (-(X1 - X2) * (x1 * y2 - x2 * y1) - (Y1 * x2 - Y2 * x1) * (x1 - x2) - (Y1 * y2 - Y2 * y1) * (y1 - y2));
        // Affine2DEstimatorCallback::computeError: float model, float arithmetic
        const float F0 = (float)S0, F1 = (float)(-S1), F2 = (float)S2, F3 = (float)S1, F4 = (float)S0, F5 = (float)S3;
        unsigned bits = 0;
        int cnt = 0;
#pragma unroll
        for (int k = 0; k < kOwn; ++k) {
            const int p = tid + k * kFitThreads;
            if (p < m) {
                const float fx = spx[p], fy = spy[p];
                const float ea = __fsub_rn(__fadd_rn(__fadd_rn(__fmul_rn(F0, fx), __fmul_rn(F1, fy)), F2), sqx[p]);
                const float eb = __fsub_rn(__fadd_rn(__fadd_rn(__fmul_rn(F3, fx), __fmul_rn(F4, fy)), F5), sqy[p]);
                const float err = __fadd_rn(__fmul_rn(ea, ea), __fmul_rn(eb, eb));
                if (err <= thr2) { bits |= 1u << k; ++cnt; }
            }
        }
        cnt = warp_sum(cnt);
        int* slot = wcount2[iter & 1];
        if (lane == 0) slot[wid] = cnt;
        __syncthreads();
        int good = 0;
#pragma unroll
        for (int w = 0; w < kFitThreads / 32; ++w) good += slot[w];
        if (good > max(max_good, 1)) {
            best_bits = bits;
            max_good = good;
            // RANSACUpdateNumIters(confidence 0.99, ep, modelPoints 2, niters)
            double ep = (double)(m - good) / (double)m;
            ep = ep < 0. ? 0. : (ep > 1. ? 1. : ep);
            const double num0 = 1. - 0.99 > 2.2250738585072014e-308 ? 1. - 0.99 : 2.2250738585072014e-308;
            const double den0 = 1. - pow(1. - ep, 2.0);
            if (den0 < 2.2250738585072014e-308) {
                niters = 0;
            } else {
                const double num = log(num0), den = log(den0);
                niters = (den >= 0 || -num >= niters * (-den)) ? niters : __double2int_rn(num / den);
            }
        }
    }
    if (max_good <= 0) {                                // no model: M empty -> identity (:241)
        if (tid == 0) {
            eye3(Tout);
            if (M) { double* Mo = M + (size_t)frame * 6; Mo[0] = 1; Mo[1] = 0; Mo[2] = 0; Mo[3] = 0; Mo[4] = 1; Mo[5] = 0; }
            if (fit_counts) { fit_counts[frame * 2] = m; fit_counts[frame * 2 + 1] = 0; }
        }
        return;
    }

    // ---- LM refinement on the consensus set == closed-form LS similarity ---------
    // pass 1: count + means
    double s1 = 0, sx = 0, sy = 0, sX = 0, sY = 0;
#pragma unroll
    for (int k = 0; k < kOwn; ++k) {
        const int p = tid + k * kFitThreads;
        if (p < m && ((best_bits >> k) & 1u)) {
            s1 += 1.0; sx += spx[p]; sy += spy[p]; sX += sqx[p]; sY += sqy[p];
        }
    }
    s1 = block_sum(s1, dscratch);
    sx = block_sum(sx, dscratch); sy = block_sum(sy, dscratch);
    sX = block_sum(sX, dscratch); sY = block_sum(sY, dscratch);
    const double mpx = sx / s1, mpy = sy / s1, mqx = sX / s1, mqy = sY / s1;
    // pass 2: centred second moments
    double den = 0, dot = 0, crs = 0;
#pragma unroll
    for (int k = 0; k < kOwn; ++k) {
        const int p = tid + k * kFitThreads;
        if (p < m && ((best_bits >> k) & 1u)) {
            const double px = spx[p] - mpx, py = spy[p] - mpy, qx = sqx[p] - mqx, qy = sqy[p] - mqy;
            den += px * px + py * py;
            dot += px * qx + py * qy;
            crs += px * qy - py * qx;
        }
    }
    den = block_sum(den, dscratch);
    dot = block_sum(dot, dscratch);
    crs = block_sum(crs, dscratch);

    if (tid == 0) {
        const double la = dot / den, lb = crs / den;
        const double ltx = mqx - (la * mpx - lb * mpy);
        const double lty = mqy - (lb * mpx + la * mpy);
        double H[9] = {la, -lb, ltx, lb, la, lty, 0.0, 0.0, 1.0};
        bool ok = isfinite(la) && isfinite(lb) && isfinite(ltx) && isfinite(lty);   // checkRange(M), :241
        if (M) {
            double* Mo = M + (size_t)frame * 6;
            Mo[0] = la; Mo[1] = -lb; Mo[2] = ltx; Mo[3] = lb; Mo[4] = la; Mo[5] = lty;
        }
        bool valid = false;
        if (ok) {
            HParams hp;
            if (decompose_h(H, cx, cy, &hp)) {           // :261-266
                hp.s = 1.0;
                compose_h(&hp, cx, cy, Tout);
                valid = true;
            } else {
                eye3(Tout);                              // :268-272
            }
        } else {
            eye3(Tout);
        }
        // second count: inliers of the consensus set, 0 when no valid transform came out (the ORB /
        // SIFT registration then keeps its previous matrix, :738-749)
        if (fit_counts) { fit_counts[frame * 2] = m; fit_counts[frame * 2 + 1] = valid ? (int)s1 : 0; }
    }
}

}  // namespace

void launch_fit(const float2* prev_pts, const float2* next_pts, const uint8_t* status,
                const int* counts, int nframes, double thresh, double cx, double cy,
                double* T, double* M, int* fit_counts, const long* /*frame_ids*/, long frame_id0,
                cudaStream_t st) {
    if (nframes <= 0) return;
    count_launch(1);
    fit_kernel<kMaxCorners><<<nframes, kFitThreads, sizeof(float) * 4 * kMaxCorners, st>>>(
        prev_pts, next_pts, status, counts, thresh, cx, cy, T, M, fit_counts, frame_id0);
}

void launch_fit_large(const float2* ref_pts, const float2* cur_pts, const uint8_t* status, const int* count,
                      double thresh, double cx, double cy, double* T, double* M, int* fit_counts, cudaStream_t st) {
    static PerDeviceOnce once;
    once.run([] { cudaFuncSetAttribute(fit_kernel<kOrbMaxKp>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(float) * 4 * kOrbMaxKp)); });
    count_launch(1);
    fit_kernel<kOrbMaxKp><<<1, kFitThreads, sizeof(float) * 4 * kOrbMaxKp, st>>>(ref_pts, cur_pts, status, count, thresh, cx,
                                                                                 cy, T, M, fit_counts, 0);
}

}  // namespace vstabk
