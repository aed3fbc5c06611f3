// fd_align.cu — 5-point similarity estimate (LMedS) + fixed-point bilinear warp to the ArcFace crop, batched.
//
// Replaces FaceAlignment::call (face_alignment.rs:27-141): cv::estimateAffinePartial2D(landmarks, template, LMEDS,
// 3.0, 2000, 0.99, 10) (:50-59) and cv::warpAffine(img, M, (112,112), INTER_LINEAR, BORDER_CONSTANT, 0) (:119-126).
//
// estimate_kernel: one thread per face, fp64, the operation order of OpenCV's calib3d/ptsetreg.cpp (RNG reseeded
// per call, 2-point exact similarity per LMedS iteration, float32 squared errors, median, inlier threshold) followed
// by the least-squares similarity over the inliers that OpenCV's LM refinement converges to.
// warp_kernel: OpenCV's imgwarp.cpp fixed-point scheme (inverse matrix in fp64, 10-bit coordinates, 5-bit sub-pixel
// position, 15-bit weights, per-tap BORDER_CONSTANT) so crops are bit-identical, not "within a few grey levels".
#include <algorithm>
#include <cmath>
#include <cfloat>
#include "fd_internal.cuh"

namespace fd {

struct EstArgs {
    const float *from;     // (F,5,2)
    const float *to;       // (F,5,2) or nullptr -> tmpl
    float tmpl[10];
    const int *count_dev;  // optional device-side F
    int F;
    int niters;            // <= 16
    int8_t pair0[16], pair1[16];  // LMedS sample pairs per iteration (host-precomputed from OpenCV's RNG)
    double *M12;           // (F,12): M (6) then inverse (6)
    double *M_out;         // optional (F,6)
    uint8_t *ok;           // (F)
    uint8_t *ok_out;       // optional (F)
};

__device__ __forceinline__ unsigned rng_next(unsigned long long &state) {
    state = (unsigned long long)(unsigned)state * 4164903690ull + (unsigned)(state >> 32);
    return (unsigned)state;
}

__device__ __forceinline__ void fit2(const float *f, const float *t, int i0, int i1, double *M) {
    double x1 = f[2 * i0], y1 = f[2 * i0 + 1], x2 = f[2 * i1], y2 = f[2 * i1 + 1];
    double X1 = t[2 * i0], Y1 = t[2 * i0 + 1], X2 = t[2 * i1], Y2 = t[2 * i1 + 1];
    double d = 1. / ((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2));
    double S0 = d * ((X1 - X2) * (x1 - x2) + (Y1 - Y2) * (y1 - y2));
    double S1 = d * ((Y1 - Y2) * (x1 - x2) - (X1 - X2) * (y1 - y2));
    double S2 = d * ((Y1 - Y2) * (x1 * y2 - x2 * y1) - (X1 * y2 - X2 * y1) * (y1 - y2) - (X1 * x2 - X2 * x1) * (x1 - x2));
    double S3 = d * (-(X1 - X2) * (x1 * y2 - x2 * y1) - (Y1 * x2 - Y2 * x1) * (x1 - x2) - (Y1 * y2 - Y2 * y1) * (y1 - y2));
    M[0] = S0; M[1] = -S1; M[2] = S2; M[3] = S1; M[4] = S0; M[5] = S3;
}
__device__ __forceinline__ void affine_err5(const float *f, const float *t, const double *M, float *err) {
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        double a = M[0] * f[2 * i] + M[1] * f[2 * i + 1] + M[2] - t[2 * i];
        double b = M[3] * f[2 * i] + M[4] * f[2 * i + 1] + M[5] - t[2 * i + 1];
        err[i] = (float)(a * a + b * b);
    }
}
__device__ __forceinline__ void invert_affine(const double *M, double *iM) {
    double D = M[0] * M[4] - M[1] * M[3];
    D = D != 0 ? 1. / D : 0;
    double A11 = M[4] * D, A22 = M[0] * D;
    iM[0] = A11; iM[1] = M[1] * (-D); iM[3] = M[3] * (-D); iM[4] = A22;
    iM[2] = -iM[0] * M[2] - iM[1] * M[5];
    iM[5] = -iM[3] * M[2] - iM[4] * M[5];
}

// 16 lanes per face: lane k < niters evaluates LMedS iteration k (the sample pairs depend only on the RNG, which OpenCV
// reseeds per call, so they are precomputed on the host); a (median, iteration) lexicographic min over the lanes picks
// the model the sequential loop would have kept (first strictly smaller median wins); lane 0 finishes.
constexpr int EST_LANES = 16;
__global__ void estimate_kernel(EstArgs a) {
    const int F = a.count_dev ? min(*a.count_dev, a.F) : a.F;
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    const int f = gt / EST_LANES, sub = gt % EST_LANES;
    const bool live = f < F;
    const int fc = live ? f : 0;
    float from[10], to[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) {
        from[k] = live ? a.from[(size_t)fc * 10 + k] : 0.0f;
        to[k] = a.to ? (live ? a.to[(size_t)fc * 10 + k] : 0.0f) : a.tmpl[k];
    }
    const int n = 5;
    double model[6] = {0, 0, 0, 0, 0, 0};
    double median = DBL_MAX;
    float err[5];
    if (live && sub < a.niters) {
        fit2(from, to, a.pair0[sub], a.pair1[sub], model);
        bool finite = true;
#pragma unroll
        for (int k = 0; k < 6; ++k) finite &= isfinite(model[k]);
        if (finite) {
            affine_err5(from, to, model, err);
            float s[5] = {err[0], err[1], err[2], err[3], err[4]};
#pragma unroll
            for (int i = 1; i < 5; ++i) {  // insertion sort of 5
                float v = s[i];
                int j = i - 1;
                while (j >= 0 && s[j] > v) { s[j + 1] = s[j]; --j; }
                s[j + 1] = v;
            }
            const double med = (double)s[2];
            if (med < DBL_MAX) median = med;  // NaN / inf medians never win (median < minMedian is false)
        }
    }
    // lexicographic (median, lane) min within each group of 16 lanes
    int win = sub;
    double best_med = median;
#pragma unroll
    for (int o = EST_LANES / 2; o > 0; o >>= 1) {
        const double om = __shfl_xor_sync(0xffffffffu, best_med, o);
        const int ow = __shfl_xor_sync(0xffffffffu, win, o);
        if (om < best_med || (om == best_med && ow < win)) { best_med = om; win = ow; }
    }
    const int base_lane = (threadIdx.x & 31) & ~(EST_LANES - 1);
    double best[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) best[k] = __shfl_sync(0xffffffffu, model[k], base_lane + win);
    if (!live || sub != 0) return;
    const double minMedian = best_med;
    bool ok = minMedian < DBL_MAX;
    double M[6] = {0, 0, 0, 0, 0, 0};
    if (ok) {
        double sigma = 2.5 * 1.4826 * (1 + 5. / (n - 2)) * sqrt(minMedian);
        if (sigma < 0.001) sigma = 0.001;
        affine_err5(from, to, best, err);
        const float thr = (float)(sigma * sigma);
        bool mask[5];
        int good = 0;
#pragma unroll
        for (int i = 0; i < 5; ++i) { mask[i] = err[i] <= thr; good += mask[i]; }
        ok = good >= 2;
        if (ok) {
            // least squares over the inliers (fixed operation order; the test oracle performs the identical sequence)
            double sx = 0, sy = 0, sX = 0, sY = 0;
#pragma unroll
            for (int i = 0; i < 5; ++i) if (mask[i]) { sx += from[2 * i]; sy += from[2 * i + 1]; sX += to[2 * i]; sY += to[2 * i + 1]; }
            double mx = sx / good, my = sy / good, mX = sX / good, mY = sY / good;
            double num_a = 0, num_b = 0, den = 0;
#pragma unroll
            for (int i = 0; i < 5; ++i) if (mask[i]) {
                double dx = from[2 * i] - mx, dy = from[2 * i + 1] - my, dX = to[2 * i] - mX, dY = to[2 * i + 1] - mY;
                num_a += dx * dX + dy * dY;
                num_b += dx * dY - dy * dX;
                den += dx * dx + dy * dy;
            }
            double sa = num_a / den, sb = num_b / den;
            M[0] = sa; M[1] = -sb; M[2] = mX - (sa * mx - sb * my);
            M[3] = sb; M[4] = sa;  M[5] = mY - (sb * mx + sa * my);
        }
    }
    double iM[6];
    invert_affine(M, iM);
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        a.M12[(size_t)f * 12 + k] = M[k];
        a.M12[(size_t)f * 12 + 6 + k] = iM[k];
        if (a.M_out) a.M_out[(size_t)f * 6 + k] = M[k];
    }
    a.ok[f] = ok ? 1 : 0;
    if (a.ok_out) a.ok_out[f] = ok ? 1 : 0;
}

// M given by the caller (fd_warp_affine): fill M12 with M and its inverse
__global__ void invert_kernel(const double *M, int F, double *M12, uint8_t *ok) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    double m[6], iM[6];
    for (int k = 0; k < 6; ++k) m[k] = M[(size_t)f * 6 + k];
    invert_affine(m, iM);
    for (int k = 0; k < 6; ++k) {
        M12[(size_t)f * 12 + k] = m[k];
        M12[(size_t)f * 12 + 6 + k] = iM[k];
    }
    ok[f] = 1;
}

constexpr int WARP_BAND = 16;     // output rows per CTA
constexpr int WARP_TX = 128, WARP_TY = 2;
constexpr int WARP_THREADS = WARP_TX * WARP_TY;

struct WarpArgs {
    const FrameDev *frames;
    const int *frame_idx;  // (F) or nullptr -> frame 0
    const double *M12;
    const uint8_t *ok;
    const int *count_dev;  // optional
    int F;
    uint8_t *crops;        // (F, ch, cw, 3)
    int cw, ch;
};

// The 6 bytes of two horizontally adjacent BGR pixels starting at p, through two aligned 64-bit loads and 32-bit
// funnel shifts (lo = bytes 0..3, hi = bytes 4..5 in its low half) instead of six byte loads.  Caller guarantees
// that [p & ~7, (p & ~7) + 16) is readable.
__device__ __forceinline__ void load6_fast(const uint8_t *p, unsigned &lo, unsigned &hi) {
    const uintptr_t ad = reinterpret_cast<uintptr_t>(p);
    const uint2 *al = reinterpret_cast<const uint2 *>(ad & ~(uintptr_t)7);
    const uint2 q0 = __ldg(al), q1 = __ldg(al + 1);
    const unsigned sh = (unsigned)(ad & 7);
    const bool up = sh >= 4;
    const unsigned a = up ? q0.y : q0.x, b = up ? q1.x : q0.y, c = up ? q1.y : q1.x;
    const unsigned sft = (sh & 3) * 8;
    lo = __funnelshift_r(a, b, sft);
    hi = __funnelshift_r(b, c, sft);
}
// same, from an 8-byte aligned base and a 32-bit byte offset (interior path: one 64-bit add per load pair)
__device__ __forceinline__ void load6_off(const uint8_t *base8, unsigned off, unsigned &lo, unsigned &hi) {
    const uint2 *al = reinterpret_cast<const uint2 *>(base8 + (off & ~7u));
    const uint2 q0 = __ldg(al), q1 = __ldg(al + 1);
    const bool up = (off & 4u) != 0;
    const unsigned a = up ? q0.y : q0.x, b = up ? q1.x : q0.y, c = up ? q1.y : q1.x;
    const unsigned sft = (off & 3u) * 8;
    lo = __funnelshift_r(a, b, sft);
    hi = __funnelshift_r(b, c, sft);
}
__device__ __forceinline__ void extract6(uint2 q0, uint2 q1, unsigned off, unsigned &lo, unsigned &hi) {
    const bool up = (off & 4u) != 0;
    const unsigned a = up ? q0.y : q0.x, b = up ? q1.x : q0.y, c = up ? q1.y : q1.x;
    const unsigned sft = (off & 3u) * 8;
    lo = __funnelshift_r(a, b, sft);
    hi = __funnelshift_r(b, c, sft);
}
// guarded variant for the frame's first/last bytes and for single-column border taps
__device__ __forceinline__ void load6_safe(const uint8_t *p, const uint8_t *lo_lim, const uint8_t *hi_lim, unsigned &lo, unsigned &hi) {
    const uintptr_t ad = reinterpret_cast<uintptr_t>(p);
    const uint8_t *al = reinterpret_cast<const uint8_t *>(ad & ~(uintptr_t)7);
    if (al >= lo_lim && al + 16 <= hi_lim) {
        load6_fast(p, lo, hi);
        return;
    }
    lo = hi = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) lo |= (unsigned)__ldg(p + k) << (8 * k);
    hi = (unsigned)__ldg(p + 4) | ((unsigned)__ldg(p + 5) << 8);
}

// OpenCV: out = (sum_ij v_ij * W_ij + 2^14) >> 15 with W_ij = a_i * b_j * 32 (a = {32-ay, ay}, b = {32-ax, ax}; the
// 15-bit table entries are exact multiples of 32).  Integer arithmetic is exact, so the sum is evaluated separably:
// horizontal 2-tap dot products with 6-bit weights (dp4a on the packed byte pair), then the vertical pair; and
// (32*S + 2^14) >> 15 == (S + 512) >> 10.
__device__ __forceinline__ void blend_store(unsigned lo0, unsigned hi0, unsigned lo1, unsigned hi1, int ax, int ay, uint8_t *o) {
    const unsigned wx = (unsigned)(32 - ax) | ((unsigned)ax << 8);   // bytes: [32-ax, ax, 0, 0]
    const int wy0 = 32 - ay, wy1 = ay;
    // byte pairs (tap0, tap1) per channel: b = bytes (0,3), g = (1,4), r = (2,5) of [lo | hi]
    const int hb0 = (int)__dp4a(__byte_perm(lo0, hi0, 0x7730), wx, 0u), hb1 = (int)__dp4a(__byte_perm(lo1, hi1, 0x7730), wx, 0u);
    const int hg0 = (int)__dp4a(__byte_perm(lo0, hi0, 0x7741), wx, 0u), hg1 = (int)__dp4a(__byte_perm(lo1, hi1, 0x7741), wx, 0u);
    const int hr0 = (int)__dp4a(__byte_perm(lo0, hi0, 0x7752), wx, 0u), hr1 = (int)__dp4a(__byte_perm(lo1, hi1, 0x7752), wx, 0u);
    o[0] = (uint8_t)((hb0 * wy0 + hb1 * wy1 + 512) >> 10);
    o[1] = (uint8_t)((hg0 * wy0 + hg1 * wy1 + 512) >> 10);
    o[2] = (uint8_t)((hr0 * wy0 + hr1 * wy1 + 512) >> 10);
}

__global__ void __launch_bounds__(WARP_THREADS) warp_kernel(WarpArgs a) {
    extern __shared__ int wsm[];
    int *adelta = wsm;             // [cw]
    int *bdelta = adelta + a.cw;   // [cw]
    int *X0s = bdelta + a.cw;      // [WARP_BAND]
    int *Y0s = X0s + WARP_BAND;    // [WARP_BAND]
    const int tid = threadIdx.y * WARP_TX + threadIdx.x;
    const int F = a.count_dev ? min(*a.count_dev, a.F) : a.F;
    const int band_y0 = blockIdx.x * WARP_BAND;
    const int rows = min(WARP_BAND, a.ch - band_y0);
    for (int f = blockIdx.y; f < F; f += gridDim.y) {
        uint8_t *crop = a.crops + (size_t)f * a.ch * a.cw * 3;
        if (!a.ok[f]) {
            for (int p = tid; p < rows * a.cw * 3; p += WARP_THREADS) crop[(size_t)band_y0 * a.cw * 3 + p] = 0;
            continue;
        }
        const FrameDev fr = a.frames[a.frame_idx ? a.frame_idx[f] : 0];
        const double *iM = a.M12 + (size_t)f * 12 + 6;
        const double i0 = iM[0], i1 = iM[1], i2 = iM[2], i3 = iM[3], i4 = iM[4], i5 = iM[5];
        __syncthreads();  // previous face's tables are no longer read
        for (int x = tid; x < a.cw; x += WARP_THREADS) {
            adelta[x] = __double2int_rn(i0 * x * 1024);  // saturate_cast<int>(M[0]*x*AB_SCALE)
            bdelta[x] = __double2int_rn(i3 * x * 1024);
        }
        if (tid < rows) {
            const int y = band_y0 + tid;
            X0s[tid] = __double2int_rn((i1 * y + i2) * 1024) + 16;  // + round_delta
            Y0s[tid] = __double2int_rn((i4 * y + i5) * 1024) + 16;
        }
        __syncthreads();
        // The map is affine, so the tap coordinates of the band are extremal at its 4 corners: if all 4 corner taps
        // (and their +1 neighbours) are interior, no pixel of the band needs a border test.
        bool interior = (reinterpret_cast<uintptr_t>(fr.data) & 7) == 0 && (fr.pitch & 7) == 0 && fr.pitch >= 16;
        {
            const int xs[2] = {0, a.cw - 1}, ys[2] = {0, rows - 1};
#pragma unroll
            for (int cy = 0; cy < 2; ++cy)
#pragma unroll
                for (int cx = 0; cx < 2; ++cx) {
                    const long long Xl = (long long)X0s[ys[cy]] + adelta[xs[cx]], Yl = (long long)Y0s[ys[cy]] + bdelta[xs[cx]];
                    const long long sx = Xl >> 10, sy = Yl >> 10;
                    interior = interior && sx >= 0 && sx + 1 < fr.w && sy >= 0 && sy + 1 < fr.h - 1;  // not the last row: 16-byte over-read
                }
        }
        const uint8_t *lo_lim = fr.data, *hi_lim = fr.data + (size_t)(fr.h - 1) * fr.pitch + (size_t)fr.w * 3;  // end of valid pixel bytes
        // lanes = consecutive output pixels of one row, so the lanes of one load instruction share cache lines
        const bool small_frame = (size_t)fr.h * fr.pitch < 0x7fffffffull;
        for (int x = threadIdx.x; x < a.cw; x += WARP_TX) {
            const unsigned adx = (unsigned)adelta[x], bdx = (unsigned)bdelta[x];
            uint8_t *o = crop + ((size_t)(band_y0 + threadIdx.y) * a.cw + x) * 3;
            const size_t ostep = (size_t)WARP_TY * a.cw * 3;
            if (interior && small_frame) {
                const unsigned pitch = (unsigned)fr.pitch;
                // two output rows per step: their 8 aligned 64-bit loads are issued before any is consumed
                for (int ry = threadIdx.y; ry < rows; ry += 2 * WARP_TY, o += 2 * ostep) {
                    const int ryb = min(ry + WARP_TY, rows - 1);
                    const bool two = ry + WARP_TY < rows;
                    const int Xa = (int)((unsigned)X0s[ry] + adx) >> 5, Ya = (int)((unsigned)Y0s[ry] + bdx) >> 5;
                    const int Xb = (int)((unsigned)X0s[ryb] + adx) >> 5, Yb = (int)((unsigned)Y0s[ryb] + bdx) >> 5;
                    const unsigned offa = (unsigned)(Ya >> 5) * pitch + (unsigned)(Xa >> 5) * 3u;
                    const unsigned offb = (unsigned)(Yb >> 5) * pitch + (unsigned)(Xb >> 5) * 3u;
                    const uint2 *pa0 = reinterpret_cast<const uint2 *>(fr.data + (offa & ~7u));
                    const uint2 *pa1 = reinterpret_cast<const uint2 *>(fr.data + ((offa + pitch) & ~7u));
                    const uint2 *pb0 = reinterpret_cast<const uint2 *>(fr.data + (offb & ~7u));
                    const uint2 *pb1 = reinterpret_cast<const uint2 *>(fr.data + ((offb + pitch) & ~7u));
                    const uint2 qa0 = __ldg(pa0), qa1 = __ldg(pa0 + 1), qa2 = __ldg(pa1), qa3 = __ldg(pa1 + 1);
                    const uint2 qb0 = __ldg(pb0), qb1 = __ldg(pb0 + 1), qb2 = __ldg(pb1), qb3 = __ldg(pb1 + 1);
                    unsigned lo0, hi0, lo1, hi1;
                    extract6(qa0, qa1, offa, lo0, hi0);
                    extract6(qa2, qa3, offa + pitch, lo1, hi1);
                    blend_store(lo0, hi0, lo1, hi1, Xa & 31, Ya & 31, o);
                    if (two) {
                        extract6(qb0, qb1, offb, lo0, hi0);
                        extract6(qb2, qb3, offb + pitch, lo1, hi1);
                        blend_store(lo0, hi0, lo1, hi1, Xb & 31, Yb & 31, o + ostep);
                    }
                }
                continue;
            }
            for (int ry = threadIdx.y; ry < rows; ry += WARP_TY, o += ostep) {
                const int X = (int)((unsigned)X0s[ry] + adx) >> 5, Y = (int)((unsigned)Y0s[ry] + bdx) >> 5;
                const int ax = X & 31, ay = Y & 31;
                unsigned lo0 = 0, hi0 = 0, lo1 = 0, hi1 = 0;
                const int sx = max(-32768, min(32767, X >> 5)), sy = max(-32768, min(32767, Y >> 5));
                const bool inx0 = sx >= 0 && sx < fr.w, inx1 = sx + 1 >= 0 && sx + 1 < fr.w;
                const bool iny0 = sy >= 0 && sy < fr.h, iny1 = sy + 1 >= 0 && sy + 1 < fr.h;
                if (inx0 && inx1) {
                    const uint8_t *p0 = fr.data + (ptrdiff_t)sy * fr.pitch + (ptrdiff_t)sx * 3;
                    if (iny0) load6_safe(p0, lo_lim, hi_lim, lo0, hi0);
                    if (iny1) load6_safe(p0 + fr.pitch, lo_lim, hi_lim, lo1, hi1);
                } else if (inx0 || inx1) {  // one tap column outside the image (BORDER_CONSTANT 0)
                    const int sxv = inx0 ? sx : sx + 1;
                    unsigned t0 = 0, t1 = 0;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        if (iny0) t0 |= (unsigned)__ldg(fr.data + (ptrdiff_t)sy * fr.pitch + sxv * 3 + c) << (8 * c);
                        if (iny1) t1 |= (unsigned)__ldg(fr.data + (ptrdiff_t)(sy + 1) * fr.pitch + sxv * 3 + c) << (8 * c);
                    }
                    if (inx0) { lo0 = t0; lo1 = t1; }              // tap 0 valid, tap 1 (bytes 3..5) zero
                    else { lo0 = t0 << 24; hi0 = t0 >> 8; lo1 = t1 << 24; hi1 = t1 >> 8; }  // tap 1 valid
                }
                blend_store(lo0, hi0, lo1, hi1, ax, ay, o);
            }
        }
    }
}

static int lmeds_niters(double p, double ep, int modelPoints, int maxIters) {  // RANSACUpdateNumIters
    p = std::max(p, 0.); p = std::min(p, 1.);
    ep = std::max(ep, 0.); ep = std::min(ep, 1.);
    double num = std::max(1. - p, DBL_MIN);
    double denom = 1. - std::pow(1. - ep, modelPoints);
    if (denom < DBL_MIN) return 0;
    num = std::log(num);
    denom = std::log(denom);
    return denom >= 0 || -num >= maxIters * (-denom) ? maxIters : (int)std::lrint(num / denom);
}

// from_dev (F,10); to_dev (F,10) or nullptr (ctx template).  count_dev optional device-side face count (<= F_cap).
int estimate_launch(fd_ctx *ctx, const float *from_dev, const float *to_dev, const int *count_dev, int F_cap,
                    double *M12_dev, double *M_out_dev, uint8_t *ok_dev, uint8_t *ok_out_dev) {
    if (F_cap <= 0) return FD_OK;
    EstArgs a;
    a.from = from_dev;
    a.to = to_dev;
    for (int i = 0; i < 10; ++i) a.tmpl[i] = (&ctx->cfg.template_landmarks[0][0])[i];
    a.count_dev = count_dev;
    a.F = F_cap;
    a.niters = std::max(lmeds_niters(0.99, 0.45, 2, 2000), 3);
    if (a.niters > 16) return fail(FD_ERR_INVALID, "estimate: LMedS iteration count exceeds the lane group");
    {   // cv::RNG(-1): state = (uint32)state * 4164903690 + (state >> 32); uniform(0,n) = next() % n
        unsigned long long st = 0xFFFFFFFFFFFFFFFFull;
        auto next = [&]() { st = (unsigned long long)(unsigned)st * 4164903690ull + (unsigned)(st >> 32); return (unsigned)st; };
        for (int it = 0; it < 16; ++it) {
            int i0 = (int)(next() % 5u), i1;
            do { i1 = (int)(next() % 5u); } while (i1 == i0);
            a.pair0[it] = (int8_t)i0;
            a.pair1[it] = (int8_t)i1;
        }
    }
    a.M12 = M12_dev;
    a.M_out = M_out_dev;
    a.ok = ok_dev;
    a.ok_out = ok_out_dev;
    estimate_kernel<<<(F_cap * EST_LANES + 127) / 128, 128, 0, ctx->stream>>>(a);
    FD_LAUNCH_CHECK(ctx);
    return FD_OK;
}

int invert_launch(fd_ctx *ctx, const double *M_dev, int F, double *M12_dev, uint8_t *ok_dev) {
    if (F <= 0) return FD_OK;
    invert_kernel<<<(F + 127) / 128, 128, 0, ctx->stream>>>(M_dev, F, M12_dev, ok_dev);
    FD_LAUNCH_CHECK(ctx);
    return FD_OK;
}

int warp_launch(fd_ctx *ctx, const FrameDev *frames_dev, const int32_t *frame_idx_dev, const double *M12_dev,
                const uint8_t *ok_dev, const int *count_dev, int F_cap, uint8_t *crops_dev, int cw, int ch) {
    if (F_cap <= 0) return FD_OK;
    WarpArgs a;
    a.frames = frames_dev;
    a.frame_idx = frame_idx_dev;
    a.M12 = M12_dev;
    a.ok = ok_dev;
    a.count_dev = count_dev;
    a.F = F_cap;
    a.crops = crops_dev;
    a.cw = cw;
    a.ch = ch;
    const int bands = (ch + WARP_BAND - 1) / WARP_BAND;
    // device-side counts: a bounded grid that strides over the faces
    int gy = count_dev ? std::min(F_cap, std::max(1, ctx->num_sms * 8 / bands)) : std::min(F_cap, 65535);
    size_t smem = sizeof(int) * (2 * (size_t)cw + 2 * WARP_BAND);
    warp_kernel<<<dim3(bands, gy), dim3(WARP_TX, WARP_TY), smem, ctx->stream>>>(a);
    FD_LAUNCH_CHECK(ctx);
    return FD_OK;
}

}  // namespace fd
