// fd_preprocess.cu — letterbox resize + zero pad + BGR->RGB + normalise + HWC->NCHW fp32, one launch per batch.
//
// Replaces RetinaFaceDetection::_preprocess (face_detection.rs:131-198: cv::resize INTER_LINEAR into the top-left
// corner of a zero canvas) and the per-pixel tensor loop (face_detection.rs:220-230).  The resize reproduces OpenCV's
// 8UC3 fixed-point scheme bit for bit (imgproc/resize.cpp: 11-bit coefficients, int32 horizontal pass,
// ((b*(T>>4))>>16 ... +2)>>2 vertical pass); a float bilinear would be off by up to 6 grey levels.
//
// Data movement: a CTA owns a band of output rows of one image.  The (one or two) source rows an output row needs are
// staged into shared memory with 128-bit coalesced loads — the second row is skipped when its vertical weight is zero
// (exact integer down-scales, e.g. 1080p -> 640x360 reads one source row in three) — the taps are then gathered from
// shared memory, and each thread writes 4 consecutive pixels of the three channel planes as 128-bit stores.
// Normalisation is a 3x256-entry table of the reference's exact (p/scale - mean)/std expression (two IEEE divides per
// entry instead of per pixel).
#include "fd_internal.cuh"

namespace fd {

constexpr int PRE_ROWS = 8;  // output rows per CTA

struct PreArgs {
    const FrameDev *frames;
    float *out;
    int out_w, out_h;
    float pixel_scale, means[3], stds[3];
    int row_buf_bytes;  // per staged row (multiple of 16, >= max(w*3)+32)
    int vec_store;      // out_w % 4 == 0
};

// stage nbytes of a source row into shared memory keeping the source's 16-byte phase; returns the pointer to byte 0
__device__ __forceinline__ const uint8_t *stage_row(uint8_t *buf, const uint8_t *__restrict__ src, int nbytes) {
    const int mis = (int)(reinterpret_cast<uintptr_t>(src) & 15);
    uint8_t *dst = buf + mis;
    const int head = mis ? min(nbytes, 16 - mis) : 0;
    for (int i = threadIdx.x; i < head; i += blockDim.x) dst[i] = __ldg(src + i);
    const int nvec = (nbytes - head) >> 4;
    const int4 *s4 = reinterpret_cast<const int4 *>(src + head);
    int4 *d4 = reinterpret_cast<int4 *>(dst + head);
    for (int v = threadIdx.x; v < nvec; v += blockDim.x) d4[v] = __ldg(s4 + v);
    const int done = head + (nvec << 4);
    for (int i = done + threadIdx.x; i < nbytes; i += blockDim.x) dst[i] = __ldg(src + i);
    return dst;
}

// horizontal + vertical fixed-point taps for one channel
__device__ __forceinline__ int resize_px(const uint8_t *r0, const uint8_t *r1, int x0, int x1, int a0, int a1, int b0, int b1) {
    int t0 = r0[x0] * a0 + r0[x1] * a1;
    int v = (b0 * (t0 >> 4)) >> 16;
    if (b1 != 0) {
        int t1 = r1[x0] * a0 + r1[x1] * a1;
        v += (b1 * (t1 >> 4)) >> 16;
    }
    return (v + 2) >> 2;
}

__device__ __forceinline__ short sat_short_rn(float v) {
    int r = __float2int_rn(v);  // round-half-even, as cvRound
    return (short)max(-32768, min(32767, r));
}

// x tap table for one destination column (cv::resize, INTER_LINEAR): byte offsets of the two taps and their weights
__device__ __forceinline__ void x_taps(int dx, double scale_x, int sw, int *o0, int *o1, short *a0, short *a1) {
    float fx = (float)(((double)dx + 0.5) * scale_x - 0.5);
    int sx = (int)floorf(fx);
    fx = __fsub_rn(fx, (float)sx);
    if (sx < 0) { fx = 0.0f; sx = 0; }
    if (sx >= sw - 1) { fx = 0.0f; sx = sw - 1; }
    *o0 = sx * 3;
    *o1 = min(sx + 1, sw - 1) * 3;
    *a0 = sat_short_rn(__fmul_rn(__fsub_rn(1.0f, fx), 2048.0f));
    *a1 = sat_short_rn(__fmul_rn(fx, 2048.0f));
}
__device__ __forceinline__ void y_taps(int dy, double scale_y, int sh, int *y0, int *y1, int *b0, int *b1) {
    float fy = (float)(((double)dy + 0.5) * scale_y - 0.5);
    int sy = (int)floorf(fy);
    fy = __fsub_rn(fy, (float)sy);
    *b0 = sat_short_rn(__fmul_rn(__fsub_rn(1.0f, fy), 2048.0f));
    *b1 = sat_short_rn(__fmul_rn(fy, 2048.0f));
    *y0 = min(max(sy, 0), sh - 1);
    *y1 = min(max(sy + 1, 0), sh - 1);
}

__global__ void preprocess_kernel(PreArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const FrameDev f = a.frames[blockIdx.y];
    const int ow = a.out_w;
    float *lut = reinterpret_cast<float *>(smem);                  // [3][256], indexed by BGR channel
    int *xo0 = reinterpret_cast<int *>(lut + 768);                 // [ow]
    int *xo1 = xo0 + ow;                                           // [ow]
    short *xa0 = reinterpret_cast<short *>(xo1 + ow);              // [ow]
    short *xa1 = xa0 + ow;                                         // [ow]
    size_t tab_bytes = 768 * 4 + (size_t)ow * 12;
    tab_bytes = (tab_bytes + 15) & ~(size_t)15;
    uint8_t *rowbuf0 = smem + tab_bytes;
    uint8_t *rowbuf1 = rowbuf0 + a.row_buf_bytes;

    for (int i = threadIdx.x; i < 768; i += blockDim.x) {
        int c = i >> 8, v = i & 255;
        lut[i] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)v, a.pixel_scale), a.means[c]), a.stds[c]);  // face_detection.rs:227
    }
    for (int x = threadIdx.x; x < f.new_w; x += blockDim.x) x_taps(x, f.scale_x, f.w, &xo0[x], &xo1[x], &xa0[x], &xa1[x]);
    __syncthreads();
    const float pad_r = lut[2 * 256], pad_g = lut[1 * 256], pad_b = lut[0];
    const size_t plane = (size_t)a.out_h * ow;
    float *out_img = a.out + (size_t)blockIdx.y * 3 * plane;
    const int ngroups = (ow + 3) >> 2;
    const int row_begin = blockIdx.x * PRE_ROWS;
    const int row_end = min(row_begin + PRE_ROWS, a.out_h);
    const int src_row_bytes = f.w * 3;

    for (int dy = row_begin; dy < row_end; ++dy) {
        float *o_r = out_img + (size_t)dy * ow;  // plane 0 = R (BGR channel 2), face_detection.rs:226-227
        float *o_g = o_r + plane;
        float *o_b = o_g + plane;
        const bool content = dy < f.new_h;
        const uint8_t *r0 = nullptr, *r1 = nullptr;
        int b0 = 0, b1 = 0;
        if (content) {
            int y0, y1;
            y_taps(dy, f.scale_y, f.h, &y0, &y1, &b0, &b1);
            r0 = stage_row(rowbuf0, f.data + (size_t)y0 * f.pitch, src_row_bytes);
            r1 = b1 != 0 ? stage_row(rowbuf1, f.data + (size_t)y1 * f.pitch, src_row_bytes) : r0;
            __syncthreads();
        }
        for (int g = threadIdx.x; g < ngroups; g += blockDim.x) {
            float vr[4], vg[4], vb[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int x = 4 * g + j;
                if (content && x < f.new_w) {
                    const int x0 = xo0[x], x1 = xo1[x], a0 = xa0[x], a1 = xa1[x];
                    vb[j] = lut[resize_px(r0, r1, x0, x1, a0, a1, b0, b1)];
                    vg[j] = lut[256 + resize_px(r0, r1, x0 + 1, x1 + 1, a0, a1, b0, b1)];
                    vr[j] = lut[512 + resize_px(r0, r1, x0 + 2, x1 + 2, a0, a1, b0, b1)];
                } else {
                    vb[j] = pad_b; vg[j] = pad_g; vr[j] = pad_r;
                }
            }
            if (a.vec_store) {
                *reinterpret_cast<float4 *>(o_r + 4 * g) = make_float4(vr[0], vr[1], vr[2], vr[3]);
                *reinterpret_cast<float4 *>(o_g + 4 * g) = make_float4(vg[0], vg[1], vg[2], vg[3]);
                *reinterpret_cast<float4 *>(o_b + 4 * g) = make_float4(vb[0], vb[1], vb[2], vb[3]);
            } else {
                for (int j = 0; j < 4 && 4 * g + j < ow; ++j) {
                    o_r[4 * g + j] = vr[j];
                    o_g[4 * g + j] = vg[j];
                    o_b[4 * g + j] = vb[j];
                }
            }
        }
        if (content) __syncthreads();  // row buffers are overwritten by the next row
    }
}

// plain cv::resize to u8 HWC (face_detection.rs:156 on its own; also the FaceAlignment fallback :98-105)
__global__ void resize_u8_kernel(FrameDev f, uint8_t *__restrict__ out, int out_h, int out_w) {
    const int dy = blockIdx.y;
    int y0, y1, b0, b1;
    y_taps(dy, f.scale_y, f.h, &y0, &y1, &b0, &b1);
    const uint8_t *r0 = f.data + (size_t)y0 * f.pitch, *r1 = f.data + (size_t)y1 * f.pitch;
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < out_w; x += gridDim.x * blockDim.x) {
        int x0, x1;
        short a0, a1;
        x_taps(x, f.scale_x, f.w, &x0, &x1, &a0, &a1);
        uint8_t *o = out + ((size_t)dy * out_w + x) * 3;
        if (f.h == out_h && f.w == out_w) {  // cv::resize copies when the size is unchanged
            o[0] = r0[x * 3]; o[1] = r0[x * 3 + 1]; o[2] = r0[x * 3 + 2];
        } else {
            o[0] = (uint8_t)resize_px(r0, r1, x0, x1, a0, a1, b0, b1);
            o[1] = (uint8_t)resize_px(r0, r1, x0 + 1, x1 + 1, a0, a1, b0, b1);
            o[2] = (uint8_t)resize_px(r0, r1, x0 + 2, x1 + 2, a0, a1, b0, b1);
        }
    }
}

int preprocess_launch(fd_ctx *ctx, const FrameDev *frames_dev, int B, float *out_nchw_dev, int max_row_bytes) {
    PreArgs a;
    a.frames = frames_dev;
    a.out = out_nchw_dev;
    a.out_w = ctx->cfg.image_w;
    a.out_h = ctx->cfg.image_h;
    a.pixel_scale = ctx->cfg.pixel_scale;
    for (int i = 0; i < 3; ++i) {
        a.means[i] = ctx->cfg.pixel_means[i];
        a.stds[i] = ctx->cfg.pixel_stds[i];
    }
    a.row_buf_bytes = ((max_row_bytes + 32) + 15) & ~15;
    a.vec_store = (a.out_w % 4 == 0) && ((reinterpret_cast<uintptr_t>(out_nchw_dev) & 15) == 0);
    size_t tab_bytes = 768 * 4 + (size_t)a.out_w * 12;
    tab_bytes = (tab_bytes + 15) & ~(size_t)15;
    size_t smem = tab_bytes + 2 * (size_t)a.row_buf_bytes;
    if (smem > (size_t)ctx->max_smem_optin) return fail(FD_ERR_INVALID, "fd_preprocess: source row too wide for shared-memory staging");
    FD_CUDA(cudaFuncSetAttribute(preprocess_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int threads = std::min(1024, std::max(64, (((a.out_w + 3) / 4 + 31) / 32) * 32));
    dim3 grid((a.out_h + PRE_ROWS - 1) / PRE_ROWS, B);
    preprocess_kernel<<<grid, threads, smem, ctx->stream>>>(a);
    FD_LAUNCH_CHECK(ctx);
    return FD_OK;
}

int resize_launch(fd_ctx *ctx, const FrameDev &frame, uint8_t *out_dev, int out_h, int out_w) {
    dim3 grid(std::max(1, std::min(64, (out_w + 127) / 128)), out_h);
    resize_u8_kernel<<<grid, 128, 0, ctx->stream>>>(frame, out_dev, out_h, out_w);
    FD_LAUNCH_CHECK(ctx);
    return FD_OK;
}

}  // namespace fd
