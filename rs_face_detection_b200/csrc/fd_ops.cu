// fd_ops.cu — drop-in single operators with HOST pointers (blocking): the functions the reference exports from
// `processing` and `rcnn`, computed on the GPU.  Each uploads its operands to ctx scratch, launches, downloads.
#include <algorithm>
#include <cstring>
#include <mutex>
#include "fd_internal.cuh"

namespace fd {

int argsort_device(fd_ctx *ctx, const float *scores_as_dets, int n, int stride, int32_t *order_dev, int32_t *flag_dev);

// ---- elementwise kernels ------------------------------------------------------------------------------------
// rcnn::anchors::anchors (anchors.rs:3-21)
__global__ void anchors_plane_kernel(int H, int W, int stride, const float *__restrict__ base, int A, float *__restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= H * W * A) return;
    int k = i % A, p = i / A;
    int iw = p % W, ih = p / W;
    float sw = (float)(iw * stride), sh = (float)(ih * stride);
    float4 o;
    o.x = __fadd_rn(base[k * 4 + 0], sw);
    o.y = __fadd_rn(base[k * 4 + 1], sh);
    o.z = __fadd_rn(base[k * 4 + 2], sw);
    o.w = __fadd_rn(base[k * 4 + 3], sh);
    reinterpret_cast<float4 *>(out)[i] = o;
}

struct BoxGeom {
    float w, h, cx, cy;
};
__device__ __forceinline__ BoxGeom box_geom(const float *b) {  // face_detection.rs:522-525
    BoxGeom g;
    g.w = __fadd_rn(__fsub_rn(b[2], b[0]), 1.0f);
    g.h = __fadd_rn(__fsub_rn(b[3], b[1]), 1.0f);
    g.cx = __fadd_rn(b[0], __fmul_rn(0.5f, __fsub_rn(g.w, 1.0f)));
    g.cy = __fadd_rn(b[1], __fmul_rn(0.5f, __fsub_rn(g.h, 1.0f)));
    return g;
}
// one thread per (row, group of 4 columns).  all_groups=0 -> bbox_pred (first group regressed, rest copied)
__global__ void bbox_pred_kernel(const float *__restrict__ boxes, const float *__restrict__ deltas, int n, int ncols,
                                 int all_groups, float *__restrict__ out) {
    const int ngroups = (ncols + 3) / 4;
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * ngroups) return;
    int i = t / ngroups, g = t % ngroups, j = 4 * g;
    const float *d = deltas + (size_t)i * ncols + j;
    float *o = out + (size_t)i * ncols + j;
    if (j + 3 < ncols && (all_groups || g == 0)) {
        BoxGeom b = box_geom(boxes + 4 * (size_t)i);
        float pcx = __fadd_rn(__fmul_rn(d[0], b.w), b.cx), pcy = __fadd_rn(__fmul_rn(d[1], b.h), b.cy);
        float pw = __fmul_rn((float)exp((double)d[2]), b.w), ph = __fmul_rn((float)exp((double)d[3]), b.h);
        float hx = __fmul_rn(0.5f, __fsub_rn(pw, 1.0f)), hy = __fmul_rn(0.5f, __fsub_rn(ph, 1.0f));
        o[0] = __fsub_rn(pcx, hx);
        o[1] = __fsub_rn(pcy, hy);
        o[2] = __fadd_rn(pcx, hx);
        o[3] = __fadd_rn(pcy, hy);
    } else {
        for (int k = 0; k < 4 && j + k < ncols; ++k) o[k] = all_groups ? 0.0f : d[k];
    }
}
__global__ void landmark_pred_kernel(const float *__restrict__ boxes, const float *__restrict__ deltas, int n, float *__restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    BoxGeom b = box_geom(boxes + 4 * (size_t)i);
#pragma unroll
    for (int p = 0; p < 5; ++p) {
        out[(size_t)i * 10 + 2 * p] = __fadd_rn(__fmul_rn(deltas[(size_t)i * 10 + 2 * p], b.w), b.cx);
        out[(size_t)i * 10 + 2 * p + 1] = __fadd_rn(__fmul_rn(deltas[(size_t)i * 10 + 2 * p + 1], b.h), b.cy);
    }
}
// group = 4 -> clip_boxes (x,y,x,y); group = 10 -> clip_points (x,y,...)   (bbox_transform.rs:27-65)
__global__ void clip_kernel(float *__restrict__ v, int rows, int cols, int group, float width, float height) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * cols) return;
    int j = t % cols;
    if (j >= (cols / group) * group) return;  // trailing partial group untouched
    float lim = ((j % group) & 1) ? height : width;
    v[t] = fmaxf(fminf(v[t], lim), 0.0f);
}
__global__ void iou_pred_kernel(const float *__restrict__ boxes, const float *__restrict__ deltas, int n, int ncols,
                                int num_classes, float *__restrict__ out) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * ncols) return;
    int i = t / ncols, j = t % ncols;
    out[t] = j < 4 * num_classes ? __fadd_rn(deltas[t], boxes[4 * (size_t)i + (j & 3)]) : 0.0f;
}
__global__ void nonlinear_transform_kernel(const float *__restrict__ ex, const float *__restrict__ gt, int n, float *__restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    BoxGeom e = box_geom(ex + 4 * (size_t)i), g = box_geom(gt + 4 * (size_t)i);
    out[4 * (size_t)i + 0] = __fdiv_rn(__fsub_rn(g.cx, e.cx), __fadd_rn(e.w, 1e-14f));
    out[4 * (size_t)i + 1] = __fdiv_rn(__fsub_rn(g.cy, e.cy), __fadd_rn(e.h, 1e-14f));
    out[4 * (size_t)i + 2] = (float)log((double)__fdiv_rn(g.w, e.w));
    out[4 * (size_t)i + 3] = (float)log((double)__fdiv_rn(g.h, e.h));
}
// rcnn::bbox::bbox_overlaps (bbox.rs:4-30)
__global__ void bbox_overlaps_kernel(const float *__restrict__ boxes, int n, const float *__restrict__ query, int k, float *__restrict__ out) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * k) return;
    int i = t / k, q = t % k;
    const float *b = boxes + 4 * (size_t)i, *qb = query + 4 * (size_t)q;
    float box_area = __fmul_rn(__fadd_rn(__fsub_rn(qb[2], qb[0]), 1.0f), __fadd_rn(__fsub_rn(qb[3], qb[1]), 1.0f));
    float r = 0.0f;
    float iw = __fadd_rn(__fsub_rn(fminf(b[2], qb[2]), fmaxf(b[0], qb[0])), 1.0f);
    if (iw > 0.0f) {
        float ih = __fadd_rn(__fsub_rn(fminf(b[3], qb[3]), fmaxf(b[1], qb[1])), 1.0f);
        if (ih > 0.0f) {
            float ba = __fmul_rn(__fadd_rn(__fsub_rn(b[2], b[0]), 1.0f), __fadd_rn(__fsub_rn(b[3], b[1]), 1.0f));
            float inter = __fmul_rn(iw, ih);
            float ua = __fsub_rn(__fadd_rn(ba, box_area), inter);
            r = __fdiv_rn(inter, ua);
        }
    }
    out[t] = r;
}

// ---- upload / download helpers --------------------------------------------------------------------------------
static int up(fd_ctx *ctx, int slot, const void *host, size_t bytes, void **dev) {
    FD_TRY(ctx->scratch[slot].reserve(bytes));
    if (bytes) FD_CUDA(cudaMemcpyAsync(ctx->scratch[slot].p, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    *dev = ctx->scratch[slot].p;
    return FD_OK;
}
static int down(fd_ctx *ctx, void *host, const void *dev, size_t bytes) {
    if (bytes) FD_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    FD_CUDA(cudaStreamSynchronize(ctx->stream));
    return FD_OK;
}
static inline int blocks(size_t n, int t = 256) { return (int)((n + t - 1) / t); }

static int nms_host(fd_ctx *ctx, const float *dets, int K, int dim, float thresh, int mode, bool presorted, int32_t *keep,
                    int *num_keep) {
    FD_TRY(check_ctx(ctx));
    FD_REQUIRE(K >= 0 && num_keep && (K == 0 || (dets && keep)), "nms: bad arguments");
    FD_REQUIRE(dim >= (presorted ? 4 : 5), "nms: boxes_dim too small");
    *num_keep = 0;
    if (K == 0) return FD_OK;
    void *d_dets;
    FD_TRY(up(ctx, 0, dets, sizeof(float) * (size_t)K * dim, &d_dets));
    FD_TRY(ctx->scratch[1].reserve(sizeof(int32_t) * (size_t)K + 16));
    int32_t *d_keep = ctx->scratch[1].as<int32_t>();
    FD_TRY(ctx->scratch[2].reserve(16));
    int32_t *d_num = ctx->scratch[2].as<int32_t>();
    FD_TRY(nms_device(ctx, (const float *)d_dets, K, dim, thresh, mode, presorted, d_keep, d_num));
    int32_t res[2] = {0, 0};
    FD_TRY(down(ctx, res, d_num, sizeof(res)));
    if (res[1]) return fail(FD_ERR_NAN_SCORE, "nms: NaN score (the reference's ordering is undefined / panics)");
    FD_REQUIRE(res[0] >= 0 && res[0] <= K, "nms: internal count out of range");
    FD_TRY(down(ctx, keep, d_keep, sizeof(int32_t) * (size_t)res[0]));
    *num_keep = res[0];
    return FD_OK;
}

}  // namespace fd

using namespace fd;

FD_EXPORT int fd_nms(fd_ctx *ctx, const float *dets, int K, float thresh, int32_t *keep, int *num_keep) {
    return nms_host(ctx, dets, K, 5, thresh, 0, false, keep, num_keep);
}
FD_EXPORT int fd_nms_device(fd_ctx *ctx, const float *dets_dev, int K, float thresh, int32_t *keep_dev, int32_t *num_keep_dev) {
    FD_TRY(check_ctx(ctx));
    FD_REQUIRE(K >= 0 && num_keep_dev && (K == 0 || (dets_dev && keep_dev)), "fd_nms_device: bad arguments");
    if (K == 0) {
        FD_CUDA(cudaMemsetAsync(num_keep_dev, 0, 2 * sizeof(int32_t), ctx->stream));
        return FD_OK;
    }
    return nms_device(ctx, dets_dev, K, 5, thresh, 0, false, keep_dev, num_keep_dev);
}
FD_EXPORT int fd_nms_last_stats(fd_ctx *ctx, int32_t *out) {
    FD_TRY(check_ctx(ctx));
    FD_REQUIRE(out, "fd_nms_last_stats: null");
    return nms_last_stats(ctx, out);
}
FD_EXPORT int fd_cpu_nms(fd_ctx *ctx, const float *dets, int K, float thresh, int32_t *keep, int *num_keep) {
    return nms_host(ctx, dets, K, 5, thresh, 1, false, keep, num_keep);
}
FD_EXPORT int fd_nms_sorted(fd_ctx *ctx, const float *boxes, int n, int boxes_dim, float thresh, int32_t *keep, int *num_out) {
    return nms_host(ctx, boxes, n, boxes_dim, thresh, 0, true, keep, num_out);
}

// The reference's literal C symbols (gpu_nms.hpp:6-8).  Errors cannot be returned through this signature; like the
// reference (nms_kernel.cu:12-19) they are reported on stdout, and *num_out is set to 0.
static std::mutex g_legacy_mu;
static fd_ctx *g_legacy_ctx[64] = {nullptr};
static int g_legacy_device = 0;
FD_EXPORT void _set_device(int device_id) {
    std::lock_guard<std::mutex> lk(g_legacy_mu);
    g_legacy_device = device_id;
}
FD_EXPORT void _nms(int32_t *keep, int *num_out, const float *boxes_host, int boxes_num, int boxes_dim, float thresh, int device_id) {
    std::lock_guard<std::mutex> lk(g_legacy_mu);
    if (num_out) *num_out = 0;
    if (device_id < 0 || device_id >= 64) device_id = g_legacy_device;
    if (!g_legacy_ctx[device_id]) {
        if (fd_ctx_create(device_id, nullptr, &g_legacy_ctx[device_id]) != FD_OK) {
            printf("_nms: %s\n", fd_last_error());
            return;
        }
    }
    int n = 0;
    if (fd_nms_sorted(g_legacy_ctx[device_id], boxes_host, boxes_num, boxes_dim, thresh, keep, &n) != FD_OK) {
        printf("_nms: %s\n", fd_last_error());
        return;
    }
    if (num_out) *num_out = n;
}

FD_EXPORT int fd_argsort_descending(fd_ctx *ctx, const float *scores, int n, int32_t *order) {
    FD_TRY(check_ctx(ctx));
    FD_REQUIRE(n >= 0 && (n == 0 || (scores && order)), "fd_argsort_descending: bad arguments");
    if (n == 0) return FD_OK;
    // the device code reads the score of row i at base[i*stride + 4]; place the scores 4 floats into the buffer
    FD_TRY(ctx->scratch[0].reserve(sizeof(float) * ((size_t)n + 4)));
    float *d_sc = ctx->scratch[0].as<float>();
    FD_CUDA(cudaMemcpyAsync(d_sc + 4, scores, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    FD_TRY(ctx->scratch[1].reserve(sizeof(int32_t) * (size_t)n));
    FD_TRY(ctx->scratch[2].reserve(16));
    FD_TRY(argsort_device(ctx, d_sc, n, 1, ctx->scratch[1].as<int32_t>(), ctx->scratch[2].as<int32_t>()));
    int32_t flag = 0;
    FD_TRY(down(ctx, &flag, ctx->scratch[2].p, sizeof(flag)));
    if (flag) return fail(FD_ERR_NAN_SCORE, "fd_argsort_descending: NaN score (utils.rs:92 panics)");
    return down(ctx, order, ctx->scratch[1].p, sizeof(int32_t) * (size_t)n);
}

FD_EXPORT int fd_anchors_plane(fd_ctx *ctx, int height, int width, int stride, const float *base_anchors, int A, float *out) {
    FD_TRY(check_ctx(ctx));
    FD_REQUIRE(height >= 0 && width >= 0 && A > 0 && base_anchors && out, "fd_anchors_plane: bad arguments");
    size_t n = (size_t)height * width * A;
    if (n == 0) return FD_OK;
    void *d_base;
    FD_TRY(up(ctx, 0, base_anchors, sizeof(float) * 4 * A, &d_base));
    FD_TRY(ctx->scratch[1].reserve(sizeof(float) * 4 * n));
    anchors_plane_kernel<<<blocks(n), 256, 0, ctx->stream>>>(height, width, stride, (const float *)d_base, A, ctx->scratch[1].as<float>());
    FD_LAUNCH_CHECK(ctx);
    return down(ctx, out, ctx->scratch[1].p, sizeof(float) * 4 * n);
}

static int pred_host(fd_ctx *ctx, const float *boxes, const float *deltas, int n, int ncols, int all_groups, float *out) {
    FD_TRY(check_ctx(ctx));
    FD_REQUIRE(n >= 0 && ncols >= 4 && (n == 0 || (boxes && deltas && out)), "bbox_pred: bad arguments");
    if (n == 0) return FD_OK;  // reference returns zeros((0, ncols))
    void *d_b, *d_d;
    FD_TRY(up(ctx, 0, boxes, sizeof(float) * 4 * (size_t)n, &d_b));
    FD_TRY(up(ctx, 1, deltas, sizeof(float) * (size_t)n * ncols, &d_d));
    FD_TRY(ctx->scratch[2].reserve(sizeof(float) * (size_t)n * ncols));
    bbox_pred_kernel<<<blocks((size_t)n * ((ncols + 3) / 4)), 256, 0, ctx->stream>>>((const float *)d_b, (const float *)d_d, n, ncols,
                                                                                  all_groups, ctx->scratch[2].as<float>());
    FD_LAUNCH_CHECK(ctx);
    return down(ctx, out, ctx->scratch[2].p, sizeof(float) * (size_t)n * ncols);
}
FD_EXPORT int fd_bbox_pred(fd_ctx *ctx, const float *boxes, const float *deltas, int n, int ncols, float *out) {
    return pred_host(ctx, boxes, deltas, n, ncols, 0, out);
}
FD_EXPORT int fd_nonlinear_pred(fd_ctx *ctx, const float *boxes, const float *deltas, int n, int ncols, float *out) {
    return pred_host(ctx, boxes, deltas, n, ncols, 1, out);
}
FD_EXPORT int fd_landmark_pred(fd_ctx *ctx, const float *boxes, const float *deltas, int n, float *out) {
    FD_TRY(check_ctx(ctx));
    FD_REQUIRE(n >= 0 && (n == 0 || (boxes && deltas && out)), "fd_landmark_pred: bad arguments");
    if (n == 0) return FD_OK;
    void *d_b, *d_d;
    FD_TRY(up(ctx, 0, boxes, sizeof(float) * 4 * (size_t)n, &d_b));
    FD_TRY(up(ctx, 1, deltas, sizeof(float) * 10 * (size_t)n, &d_d));
    FD_TRY(ctx->scratch[2].reserve(sizeof(float) * 10 * (size_t)n));
    landmark_pred_kernel<<<blocks(n), 256, 0, ctx->stream>>>((const float *)d_b, (const float *)d_d, n, ctx->scratch[2].as<float>());
    FD_LAUNCH_CHECK(ctx);
    return down(ctx, out, ctx->scratch[2].p, sizeof(float) * 10 * (size_t)n);
}
static int clip_host(fd_ctx *ctx, float *v, int rows, int cols, int group, int im_h, int im_w) {
    FD_TRY(check_ctx(ctx));
    FD_REQUIRE(rows >= 0 && cols >= 0 && (rows * cols == 0 || v), "clip: bad arguments");
    size_t n = (size_t)rows * cols;
    if (n == 0) return FD_OK;
    void *d;
    FD_TRY(up(ctx, 0, v, sizeof(float) * n, &d));
    clip_kernel<<<blocks(n), 256, 0, ctx->stream>>>((float *)d, rows, cols, group, (float)im_w - 1.0f, (float)im_h - 1.0f);
    FD_LAUNCH_CHECK(ctx);
    return down(ctx, v, d, sizeof(float) * n);
}
FD_EXPORT int fd_clip_boxes(fd_ctx *ctx, float *boxes, int rows, int cols, int im_h, int im_w) {
    return clip_host(ctx, boxes, rows, cols, 4, im_h, im_w);
}
FD_EXPORT int fd_clip_points(fd_ctx *ctx, float *points, int rows, int cols, int im_h, int im_w) {
    return clip_host(ctx, points, rows, cols, 10, im_h, im_w);
}
FD_EXPORT int fd_iou_pred(fd_ctx *ctx, const float *boxes, const float *deltas, int n, int ncols, int num_classes, float *out) {
    FD_TRY(check_ctx(ctx));
    FD_REQUIRE(n >= 0 && ncols >= 4 * num_classes && num_classes >= 0 && (n == 0 || (boxes && deltas && out)), "fd_iou_pred: bad arguments");
    if (n == 0) return FD_OK;
    void *d_b, *d_d;
    FD_TRY(up(ctx, 0, boxes, sizeof(float) * 4 * (size_t)n, &d_b));
    FD_TRY(up(ctx, 1, deltas, sizeof(float) * (size_t)n * ncols, &d_d));
    FD_TRY(ctx->scratch[2].reserve(sizeof(float) * (size_t)n * ncols));
    iou_pred_kernel<<<blocks((size_t)n * ncols), 256, 0, ctx->stream>>>((const float *)d_b, (const float *)d_d, n, ncols, num_classes,
                                                                     ctx->scratch[2].as<float>());
    FD_LAUNCH_CHECK(ctx);
    return down(ctx, out, ctx->scratch[2].p, sizeof(float) * (size_t)n * ncols);
}
FD_EXPORT int fd_nonlinear_transform(fd_ctx *ctx, const float *ex, const float *gt, int n, float *out) {
    FD_TRY(check_ctx(ctx));
    FD_REQUIRE(n >= 0 && (n == 0 || (ex && gt && out)), "fd_nonlinear_transform: bad arguments");
    if (n == 0) return FD_OK;
    void *d_e, *d_g;
    FD_TRY(up(ctx, 0, ex, sizeof(float) * 4 * (size_t)n, &d_e));
    FD_TRY(up(ctx, 1, gt, sizeof(float) * 4 * (size_t)n, &d_g));
    FD_TRY(ctx->scratch[2].reserve(sizeof(float) * 4 * (size_t)n));
    nonlinear_transform_kernel<<<blocks(n), 256, 0, ctx->stream>>>((const float *)d_e, (const float *)d_g, n, ctx->scratch[2].as<float>());
    FD_LAUNCH_CHECK(ctx);
    return down(ctx, out, ctx->scratch[2].p, sizeof(float) * 4 * (size_t)n);
}
FD_EXPORT int fd_bbox_overlaps(fd_ctx *ctx, const float *boxes, int n, const float *query, int k, float *out) {
    FD_TRY(check_ctx(ctx));
    FD_REQUIRE(n >= 0 && k >= 0 && ((size_t)n * k == 0 || (boxes && query && out)), "fd_bbox_overlaps: bad arguments");
    if ((size_t)n * k == 0) return FD_OK;
    void *d_b, *d_q;
    FD_TRY(up(ctx, 0, boxes, sizeof(float) * 4 * (size_t)n, &d_b));
    FD_TRY(up(ctx, 1, query, sizeof(float) * 4 * (size_t)k, &d_q));
    FD_TRY(ctx->scratch[2].reserve(sizeof(float) * (size_t)n * k));
    bbox_overlaps_kernel<<<blocks((size_t)n * k), 256, 0, ctx->stream>>>((const float *)d_b, n, (const float *)d_q, k, ctx->scratch[2].as<float>());
    FD_LAUNCH_CHECK(ctx);
    return down(ctx, out, ctx->scratch[2].p, sizeof(float) * (size_t)n * k);
}
