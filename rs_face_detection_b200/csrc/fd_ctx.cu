// fd_ctx.cu — context, error plumbing, memory helpers, config defaults and the init-time anchor tables.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include "fd_internal.cuh"

namespace fd {

static thread_local std::string g_last_error;

void set_error(const std::string &msg) { g_last_error = msg; }
int fail(int code, const std::string &msg) {
    g_last_error = msg;
    return code;
}

int DevBuf::reserve(size_t bytes) {
    if (bytes <= cap) return FD_OK;
    if (p) {
        FD_CUDA(cudaFree(p));
        p = nullptr;
        cap = 0;
    }
    size_t want = std::max<size_t>(bytes, 256);
    FD_CUDA(cudaMalloc(&p, want));
    cap = want;
    return FD_OK;
}
void DevBuf::release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
}
int PinnedBuf::reserve(size_t bytes) {
    if (bytes <= cap) return FD_OK;
    if (p) {
        FD_CUDA(cudaFreeHost(p));
        p = nullptr;
        cap = 0;
    }
    size_t want = std::max<size_t>(bytes, 256);
    FD_CUDA(cudaMallocHost(&p, want));
    cap = want;
    return FD_OK;
}
void PinnedBuf::release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
}

// A mark = event recorded right after a launch; the time attributed to the launch is the gap since the previous
// mark on the same stream (kernel + its launch gap), which is what a serial step pays for it.
void trace_mark(fd_ctx *ctx, const char *file, int line, const char *name) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, ctx->stream);
    std::string key;
    if (name) key = name;
    else {
        const char *base = strrchr(file, '/');
        key = std::string(base ? base + 1 : file) + ":" + std::to_string(line);
    }
    ctx->trace.emplace_back(key, e);
}
static void trace_clear(fd_ctx *ctx) {
    for (auto &t : ctx->trace) cudaEventDestroy(t.second);
    ctx->trace.clear();
}
static void trace_dump(fd_ctx *ctx) {
    if (ctx->trace.size() < 2) return;
    fprintf(stderr, "[fd trace] %zu marks\n", ctx->trace.size());
    for (size_t i = 1; i < ctx->trace.size(); ++i) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->trace[i - 1].second, ctx->trace[i].second);
        fprintf(stderr, "[fd trace] %-28s +%8.1f us\n", ctx->trace[i].first.c_str(), ms * 1e3f);
    }
    trace_clear(ctx);
}

int check_ctx(const fd_ctx *ctx) {
    if (!ctx) return fail(FD_ERR_INVALID, "null fd_ctx");
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return fail(FD_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    return FD_OK;
}

// ---- init-time anchor tables (host arithmetic, fp32, same operation order as generate_anchors.rs) ------
struct WH {
    float w, h, cx, cy;
};
static WH whctrs(const float *a) {  // generate_anchors.rs:20-26
    WH r;
    r.w = a[2] - a[0] + 1.0f;
    r.h = a[3] - a[1] + 1.0f;
    r.cx = a[0] + 0.5f * (r.w - 1.0f);
    r.cy = a[1] + 0.5f * (r.h - 1.0f);
    return r;
}
static void mk(float ws, float hs, float cx, float cy, float *o) {  // :28-39
    o[0] = cx - 0.5f * (ws - 1.0f);
    o[1] = cy - 0.5f * (hs - 1.0f);
    o[2] = cx + 0.5f * (ws - 1.0f);
    o[3] = cy + 0.5f * (hs - 1.0f);
}
static int gen_anchors2(int base_size, const float *ratios, int nr, const float *scales, int ns, int stride, bool dense,
                        float *out) {
    const float base[4] = {0.0f, 0.0f, (float)base_size - 1.0f, (float)base_size - 1.0f};
    WH b = whctrs(base);
    const float size = b.w * b.h;
    int n = 0;
    for (int r = 0; r < nr; ++r) {
        // _ratio_enum (:141-148): ws rounded (half away from zero), hs = ws*ratio NOT rounded
        float ws = roundf(sqrtf(size / ratios[r]));
        float hs = ws * ratios[r];
        float ra[4];
        mk(ws, hs, b.cx, b.cy, ra);
        WH a = whctrs(ra);
        for (int s = 0; s < ns; ++s, ++n) mk(a.w * scales[s], a.h * scales[s], a.cx, a.cy, out + 4 * n);  // _scale_enum
    }
    if (dense) {  // :80-90
        for (int i = 0; i < 4 * n; ++i) out[4 * n + i] = out[i] + (float)stride / 2.0f;
        n *= 2;
    }
    return n;
}

}  // namespace fd

using namespace fd;

FD_EXPORT int fd_abi_version(void) { return FD_ABI_VERSION; }
FD_EXPORT const char *fd_last_error(void) { return g_last_error.c_str(); }

FD_EXPORT int fd_generate_anchors2(int base_size, const float *ratios, int n_ratios, const float *scales, int n_scales,
                                   int stride, int dense_anchor, float *out, int *n_out) {
    FD_REQUIRE(ratios && scales && out && n_ratios > 0 && n_scales > 0, "fd_generate_anchors2: bad arguments");
    int n = gen_anchors2(base_size, ratios, n_ratios, scales, n_scales, stride, dense_anchor != 0, out);
    if (n_out) *n_out = n;
    return FD_OK;
}
FD_EXPORT int fd_generate_anchors(int base_size, const float *ratios, int n_ratios, const float *scales, int n_scales,
                                  float *out, int *n_out) {
    return fd_generate_anchors2(base_size, ratios, n_ratios, scales, n_scales, 0, 0, out, n_out);
}
FD_EXPORT int fd_generate_anchors_fpn(const int *base_size, const float *ratios, const float *scales, int n_levels,
                                      float *out) {
    FD_REQUIRE(base_size && ratios && scales && out && n_levels > 0, "fd_generate_anchors_fpn: bad arguments");
    for (int i = 0; i < n_levels; ++i) gen_anchors2(base_size[i], ratios + i, 1, scales + i, 1, 0, false, out + 4 * i);
    return FD_OK;
}
FD_EXPORT int fd_generate_anchors_fpn2(int dense_anchor, const fd_anchor_cfg *cfg, int n_cfg, float *out,
                                       int *rows_per_stride, int *strides_sorted) {
    FD_REQUIRE(cfg && out && n_cfg > 0 && n_cfg <= FD_MAX_STRIDES, "fd_generate_anchors_fpn2: bad arguments");
    std::vector<int> order(n_cfg);
    for (int i = 0; i < n_cfg; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return cfg[a].stride > cfg[b].stride; });  // :123-124
    int row = 0;
    for (int k = 0; k < n_cfg; ++k) {
        const fd_anchor_cfg &c = cfg[order[k]];
        FD_REQUIRE(c.n_ratios > 0 && c.n_ratios <= 8 && c.n_scales > 0 && c.n_scales <= 8, "fd_generate_anchors_fpn2: bad cfg");
        int n = gen_anchors2(c.base_size, c.ratios, c.n_ratios, c.scales, c.n_scales, c.stride, dense_anchor != 0,
                             out + 4 * row);
        if (rows_per_stride) rows_per_stride[k] = n;
        if (strides_sorted) strides_sorted[k] = c.stride;
        row += n;
    }
    return FD_OK;
}

FD_EXPORT int fd_config_default(fd_config *cfg) {
    FD_REQUIRE(cfg, "fd_config_default: null cfg");
    memset(cfg, 0, sizeof(*cfg));
    cfg->image_w = 640;  // config.rs:27
    cfg->image_h = 640;
    cfg->conf_thr = 0.7f;   // config.rs:29
    cfg->iou_thr = 0.45f;   // config.rs:30
    cfg->n_strides = 3;     // face_detection.rs:52
    fd_anchor_cfg ac[3];
    memset(ac, 0, sizeof(ac));
    const int strides[3] = {32, 16, 8};
    const float scales[3][2] = {{32.0f, 16.0f}, {8.0f, 4.0f}, {2.0f, 1.0f}};  // face_detection.rs:56-80
    for (int i = 0; i < 3; ++i) {
        ac[i].stride = strides[i];
        ac[i].base_size = 16;
        ac[i].n_ratios = 1;
        ac[i].ratios[0] = 1.0f;
        ac[i].n_scales = 2;
        ac[i].scales[0] = scales[i][0];
        ac[i].scales[1] = scales[i][1];
        ac[i].allowed_border = 9999;
    }
    float anchors[3 * 2 * 4];
    int rows[3], sorted[3];
    FD_TRY(fd_generate_anchors_fpn2(0, ac, 3, anchors, rows, sorted));
    cfg->num_anchors = rows[0];
    for (int s = 0; s < 3; ++s) {
        cfg->strides[s] = sorted[s];
        for (int a = 0; a < 2; ++a)
            for (int k = 0; k < 4; ++k) cfg->base_anchors[s][a][k] = anchors[(s * 2 + a) * 4 + k];
    }
    for (int i = 0; i < 3; ++i) {  // face_detection.rs:105-107
        cfg->pixel_means[i] = 0.0f;
        cfg->pixel_stds[i] = 1.0f;
    }
    cfg->pixel_scale = 1.0f;
    for (int i = 0; i < 4; ++i) cfg->bbox_stds[i] = 1.0f;  // :91
    cfg->landmark_std = 1.0f;                              // :92
    cfg->crop_w = 112;                                     // config.rs:45
    cfg->crop_h = 112;
    const float tmpl[5][2] = {{38.2946f, 51.6963f}, {73.5318f, 51.5014f}, {56.0252f, 71.7366f},
                              {41.5493f, 92.3655f}, {70.7299f, 92.2041f}};  // config.rs:46-52
    memcpy(cfg->template_landmarks, tmpl, sizeof(tmpl));
    return FD_OK;
}

FD_EXPORT int fd_device_count(int *count) {
    FD_REQUIRE(count, "fd_device_count: null");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        return fail(FD_ERR_NO_DEVICE, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
    }
    *count = n;
    return FD_OK;
}

FD_EXPORT int fd_ctx_create(int device_id, const fd_config *cfg, fd_ctx **out) {
    FD_REQUIRE(out, "fd_ctx_create: null out");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(FD_ERR_NO_DEVICE, "no CUDA device: libfd_b200 has no CPU fallback (" +
                                          std::string(e != cudaSuccess ? cudaGetErrorString(e) : "0 devices") + ")");
    FD_REQUIRE(device_id >= 0 && device_id < n, "fd_ctx_create: device_id out of range");
    fd_config c;
    if (cfg) c = *cfg;
    else FD_TRY(fd_config_default(&c));
    FD_REQUIRE(c.image_w > 0 && c.image_h > 0 && c.n_strides > 0 && c.n_strides <= FD_MAX_STRIDES, "fd_ctx_create: bad geometry");
    FD_REQUIRE(c.num_anchors > 0 && c.num_anchors <= FD_MAX_ANCHORS, "fd_ctx_create: bad num_anchors");
    FD_REQUIRE(c.crop_w > 0 && c.crop_h > 0 && c.crop_w <= 1024 && c.crop_h <= 1024, "fd_ctx_create: bad crop size");
    for (int s = 0; s < c.n_strides; ++s) FD_REQUIRE(c.strides[s] > 0, "fd_ctx_create: bad stride");
    FD_CUDA(cudaSetDevice(device_id));
    // every failure below goes through the guard, so a half-built ctx (streams, events) is released, not leaked
    struct Guard {
        fd_ctx *p;
        ~Guard() { if (p) fd_ctx_destroy(p); }
    } guard{new fd_ctx()};
    fd_ctx *ctx = guard.p;
    ctx->device = device_id;
    ctx->cfg = c;
    ctx->trace_on = getenv("FD_TRACE") != nullptr && getenv("FD_TRACE")[0] == '1';
    cudaDeviceProp prop;
    FD_CUDA(cudaGetDeviceProperties(&prop, device_id));
    ctx->num_sms = prop.multiProcessorCount;
    ctx->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    FD_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    FD_CUDA(cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking));
    for (int i = 0; i < 4; ++i) FD_CUDA(cudaEventCreateWithFlags(&ctx->ev[i], cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) FD_CUDA(cudaEventCreateWithFlags(&ctx->ev_j[i], cudaEventDisableTiming));
    FD_CUDA(cudaEventCreateWithFlags(&ctx->ev_block, cudaEventDisableTiming | cudaEventBlockingSync));
    DecodeCfg &d = ctx->dcfg;
    memset(&d, 0, sizeof(d));
    d.n_strides = c.n_strides;
    d.A = c.num_anchors;
    for (int s = 0; s < c.n_strides; ++s) {
        d.stride[s] = c.strides[s];
        d.fh[s] = (c.image_h + c.strides[s] - 1) / c.strides[s];
        d.fw[s] = (c.image_w + c.strides[s] - 1) / c.strides[s];
        d.pos_off[s + 1] = d.pos_off[s] + d.fh[s] * d.fw[s];
        d.anchor_off[s + 1] = d.anchor_off[s] + d.fh[s] * d.fw[s] * d.A;
        for (int a = 0; a < d.A; ++a)
            for (int k = 0; k < 4; ++k) d.base[s][a][k] = c.base_anchors[s][a][k];
    }
    d.total_pos = d.pos_off[c.n_strides];
    d.total_anchors = d.anchor_off[c.n_strides];
    for (int k = 0; k < 4; ++k) d.bbox_stds[k] = c.bbox_stds[k];
    d.landmark_std = c.landmark_std;
    d.clip_w = (float)c.image_w - 1.0f;
    d.clip_h = (float)c.image_h - 1.0f;
    guard.p = nullptr;
    *out = ctx;
    return FD_OK;
}

FD_EXPORT void fd_ctx_destroy(fd_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->stream2) cudaStreamSynchronize(ctx->stream2);
    for (auto &b : ctx->scratch) b.release();
    for (auto &b : ctx->pinned) b.release();
    ctx->jpeg_coef_host.release();
    ctx->jpeg_desc_host.release();
    ctx->jpeg_aux_host.release();
    ctx->jpeg_flags_host.release();
    DevBuf *bufs[] = {&ctx->frames_dev, &ctx->det_scale_dev, &ctx->cand_count, &ctx->cand_keys, &ctx->cand_box,
                      &ctx->cand_lmk, &ctx->keep_src, &ctx->keep_count, &ctx->status_dev, &ctx->big_list,
                      &ctx->out_offsets, &ctx->out_det, &ctx->out_lmk, &ctx->out_frame_idx, &ctx->align_M,
                      &ctx->align_ok, &ctx->tickets, &ctx->scan_agg, &ctx->select_sel, &ctx->select_lmk, &ctx->select_fidx, &ctx->pipe_frames, &ctx->pipe_tensor, &ctx->pipe_crops, &ctx->pipe_mode, &ctx->jpeg_coef, &ctx->jpeg_planes, &ctx->jpeg_frames, &ctx->jpeg_desc, &ctx->jpeg_raw, &ctx->jpeg_stream, &ctx->jpeg_aux, &ctx->jpeg_sync, &ctx->jpeg_flags};
    for (auto *b : bufs) b->release();
    for (auto &b : ctx->nms_ws) b.release();
    for (auto &b : ctx->nms_ws_sp) b.release();
    for (auto &b : ctx->pipe_heads) b.release();
    for (int i = 0; i < 4; ++i)
        if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    for (int i = 0; i < 2; ++i)
        if (ctx->ev_j[i]) cudaEventDestroy(ctx->ev_j[i]);
    if (ctx->ev_block) cudaEventDestroy(ctx->ev_block);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    delete ctx;
}

FD_EXPORT int fd_ctx_get_config(const fd_ctx *ctx, fd_config *out) {
    FD_REQUIRE(ctx && out, "fd_ctx_get_config: null");
    *out = ctx->cfg;
    return FD_OK;
}
FD_EXPORT int fd_ctx_total_anchors(const fd_ctx *ctx, int32_t *out) {
    FD_REQUIRE(ctx && out, "fd_ctx_total_anchors: null");
    *out = ctx->dcfg.total_anchors;
    return FD_OK;
}
FD_EXPORT void *fd_ctx_stream(fd_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
FD_EXPORT int fd_ctx_synchronize(fd_ctx *ctx) {
    FD_TRY(check_ctx(ctx));
    FD_CUDA(cudaStreamSynchronize(ctx->stream));
    FD_CUDA(cudaStreamSynchronize(ctx->stream2));
    if (ctx->trace_on && !ctx->profile_on) trace_dump(ctx);
    return FD_OK;
}

// ---- per-kernel profile through the ABI (bench.py's per-kernel roofline lines) -----------------------------------------
FD_EXPORT int fd_ctx_profile(fd_ctx *ctx, int enable) {
    FD_TRY(check_ctx(ctx));
    FD_CUDA(cudaStreamSynchronize(ctx->stream));
    trace_clear(ctx);
    ctx->profile_on = enable != 0;
    ctx->trace_on = ctx->profile_on || (getenv("FD_TRACE") != nullptr && getenv("FD_TRACE")[0] == '1');
    if (ctx->profile_on) trace_mark(ctx, __FILE__, __LINE__, "(begin)");
    return FD_OK;
}
FD_EXPORT int fd_ctx_profile_fetch(fd_ctx *ctx, char *buf, size_t cap) {
    FD_TRY(check_ctx(ctx));
    FD_REQUIRE(buf && cap > 0, "fd_ctx_profile_fetch: bad buffer");
    FD_CUDA(cudaStreamSynchronize(ctx->stream));
    std::vector<std::string> names;
    std::vector<double> tot;
    std::vector<int> cnt;
    for (size_t i = 1; i < ctx->trace.size(); ++i) {
        float ms = 0.f;
        FD_CUDA(cudaEventElapsedTime(&ms, ctx->trace[i - 1].second, ctx->trace[i].second));
        const std::string &k = ctx->trace[i].first;
        if (k == "(begin)") continue;
        size_t j = 0;
        while (j < names.size() && names[j] != k) ++j;
        if (j == names.size()) { names.push_back(k); tot.push_back(0); cnt.push_back(0); }
        tot[j] += ms * 1e3;
        cnt[j] += 1;
    }
    std::string out;
    for (size_t j = 0; j < names.size(); ++j) {
        char line[256];
        snprintf(line, sizeof(line), "%s %d %.3f\n", names[j].c_str(), cnt[j], tot[j]);
        out += line;
    }
    trace_clear(ctx);
    if (ctx->profile_on) trace_mark(ctx, __FILE__, __LINE__, "(begin)");
    FD_REQUIRE(out.size() + 1 <= cap, "fd_ctx_profile_fetch: buffer too small");
    memcpy(buf, out.c_str(), out.size() + 1);
    return FD_OK;
}
FD_EXPORT int fd_ctx_set_sharing(fd_ctx *ctx, int contexts_in_flight) {
    FD_TRY(check_ctx(ctx));
    ctx->share_sms = contexts_in_flight > 1;
    return FD_OK;
}
FD_EXPORT int fd_ctx_launch_count(const fd_ctx *ctx, int64_t *out) {
    FD_REQUIRE(ctx && out, "fd_ctx_launch_count: null");
    *out = ctx->launches;
    return FD_OK;
}

FD_EXPORT int fd_dev_alloc(fd_ctx *ctx, size_t bytes, void **out) {
    FD_TRY(check_ctx(ctx));
    FD_REQUIRE(out, "fd_dev_alloc: null out");
    FD_CUDA(cudaMalloc(out, bytes ? bytes : 1));
    return FD_OK;
}
FD_EXPORT int fd_dev_free(fd_ctx *ctx, void *ptr) {
    FD_TRY(check_ctx(ctx));
    FD_CUDA(cudaFree(ptr));
    return FD_OK;
}
FD_EXPORT int fd_host_alloc_pinned(size_t bytes, void **out) {
    FD_REQUIRE(out, "fd_host_alloc_pinned: null out");
    FD_CUDA(cudaMallocHost(out, bytes ? bytes : 1));
    return FD_OK;
}
FD_EXPORT int fd_host_free_pinned(void *ptr) {
    FD_CUDA(cudaFreeHost(ptr));
    return FD_OK;
}
FD_EXPORT int fd_memcpy_h2d_async(fd_ctx *ctx, void *dst, const void *src, size_t bytes) {
    FD_TRY(check_ctx(ctx));
    FD_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return FD_OK;
}
FD_EXPORT int fd_memcpy_d2h_async(fd_ctx *ctx, void *dst, const void *src, size_t bytes) {
    FD_TRY(check_ctx(ctx));
    FD_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return FD_OK;
}
FD_EXPORT int fd_memcpy_h2d(fd_ctx *ctx, void *dst, const void *src, size_t bytes) {
    FD_TRY(fd_memcpy_h2d_async(ctx, dst, src, bytes));
    FD_CUDA(cudaStreamSynchronize(ctx->stream));
    return FD_OK;
}
FD_EXPORT int fd_memcpy_d2h(fd_ctx *ctx, void *dst, const void *src, size_t bytes) {
    FD_TRY(fd_memcpy_d2h_async(ctx, dst, src, bytes));
    FD_CUDA(cudaStreamSynchronize(ctx->stream));
    return FD_OK;
}
FD_EXPORT int fd_memset_dev(fd_ctx *ctx, void *dst, int value, size_t bytes) {
    FD_TRY(check_ctx(ctx));
    FD_CUDA(cudaMemsetAsync(dst, value, bytes, ctx->stream));
    return FD_OK;
}
