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
constexpr int WARP_TX = 32, WARP_TY = 8;
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

// The 6 bytes of two horizontally adjacent BGR pixels starting at p, through at most two aligned 64-bit loads (one
// when the 6 bytes sit inside one 8-byte word) instead of six byte loads: the warp is L1-wavefront bound.
__device__ __forceinline__ unsigned long long load6(const uint8_t *p, const uint8_t *lo_lim, const uint8_t *hi_lim) {
    const uintptr_t ad = reinterpret_cast<uintptr_t>(p);
    const uint8_t *al = reinterpret_cast<const uint8_t *>(ad & ~(uintptr_t)7);
    const int sh = (int)(ad & 7);
    if (al >= lo_lim && al + 16 <= hi_lim) {
        unsigned long long lo = __ldg(reinterpret_cast<const unsigned long long *>(al));
        if (sh <= 2) return lo >> (8 * sh);
        unsigned long long hi = __ldg(reinterpret_cast<const unsigned long long *>(al + 8));
        return (lo >> (8 * sh)) | (hi << (64 - 8 * sh));
    }
    unsigned long long v = 0;  // first / last bytes of the frame buffer: plain byte loads
#pragma unroll
    for (int k = 0; k < 6; ++k) v |= (unsigned long long)__ldg(p + k) << (8 * k);
    return v;
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
    const int ngroups = (a.cw + 3) >> 2;
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
        const uint8_t *lo_lim = fr.data, *hi_lim = fr.data + (size_t)(fr.h - 1) * fr.pitch + (size_t)fr.w * 3;  // end of valid pixel bytes
        // consecutive lanes = consecutive output pixels, so the lanes of one load instruction share cache lines
        for (int p = tid; p < rows * a.cw; p += WARP_THREADS) {
            const int ry = p / a.cw, x = p - ry * a.cw;
            const int X = (int)((unsigned)X0s[ry] + (unsigned)adelta[x]) >> 5;
            const int Y = (int)((unsigned)Y0s[ry] + (unsigned)bdelta[x]) >> 5;
            const int sx = max(-32768, min(32767, X >> 5)), sy = max(-32768, min(32767, Y >> 5));
            const int ax = X & 31, ay = Y & 31;
            const int w00 = (32 - ay) * (32 - ax) * 32, w01 = (32 - ay) * ax * 32;
            const int w10 = ay * (32 - ax) * 32, w11 = ay * ax * 32;
            const bool inx0 = sx >= 0 && sx < fr.w, inx1 = sx + 1 >= 0 && sx + 1 < fr.w;
            const bool iny0 = sy >= 0 && sy < fr.h, iny1 = sy + 1 >= 0 && sy + 1 < fr.h;
            unsigned long long t0 = 0, t1 = 0;  // rows sy, sy+1: bytes [b0 g0 r0 b1 g1 r1]
            if (inx0 && inx1) {
                const uint8_t *p0 = fr.data + (ptrdiff_t)sy * fr.pitch + (ptrdiff_t)sx * 3;
                if (iny0) t0 = load6(p0, lo_lim, hi_lim);
                if (iny1) t1 = load6(p0 + fr.pitch, lo_lim, hi_lim);
            } else if (inx0 || inx1) {  // one tap column outside the image (BORDER_CONSTANT 0)
                const int sxv = inx0 ? sx : sx + 1;
                const int shl = inx0 ? 0 : 24;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    if (iny0) t0 |= (unsigned long long)__ldg(fr.data + (ptrdiff_t)sy * fr.pitch + sxv * 3 + c) << (8 * c + shl);
                    if (iny1) t1 |= (unsigned long long)__ldg(fr.data + (ptrdiff_t)(sy + 1) * fr.pitch + sxv * 3 + c) << (8 * c + shl);
                }
            }
            uint8_t *o = crop + ((size_t)(band_y0 + ry) * a.cw + x) * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int v00 = (int)((t0 >> (8 * c)) & 0xff), v01 = (int)((t0 >> (8 * c + 24)) & 0xff);
                const int v10 = (int)((t1 >> (8 * c)) & 0xff), v11 = (int)((t1 >> (8 * c + 24)) & 0xff);
                o[c] = (uint8_t)((v00 * w00 + v01 * w01 + v10 * w10 + v11 * w11 + (1 << 14)) >> 15);
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
