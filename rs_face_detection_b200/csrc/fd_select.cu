// fd_select.cu — FaceSelection::call (pipeline/module/face_selection.rs:72-189) on the device: picks ONE detection per image
// (centre / edge-margin / minimum-size heuristics, or the biggest face when enrolling) and the row whose key points go
// with it, so the reference's detect -> select -> align flow (face_pipeline/pipeline.rs:196-232) needs no host round trip
// between NMS and the warp.  SURVEY 8(f) row N3.
//
// One warp per image over its compact detection rows.  Every float operation is rounded separately in the reference's
// order (the library is built with -fmad=false); "first maximum wins" (strict `>`, :148-151 and :42-46) is a
// (value, lowest index) arg-max, and the key-point row is the FIRST row within 2 px of the selected box (:160-176).
#include "fd_internal.cuh"

namespace fd {

struct SelectArgs {
    const int *offsets;      // (B+1) rows of image b = [offsets[b], offsets[b+1])
    const float *det;        // (total,5)
    const float *lmk;        // (total,10) or nullptr (key_points == None)
    const FrameDev *frames;  // (B) image sizes
    int B;
    float mcl_ratio, mcr_ratio, me_ratio, min_ratio;
    int enroll;
    int *sel;                // (B,2): global row of the selected box, global row of its key points (-1 = None)
    float *sel_lmk;          // optional (B,10): key points of the selection, NaN when there is none
    int *sel_frame_idx;      // optional (B)
};

__device__ __forceinline__ void warp_argmax_first(float &v, int &i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, i, o);
        if (oi >= 0 && (i < 0 || ov > v || (ov == v && oi < i))) { v = ov; i = oi; }
    }
}

__global__ void __launch_bounds__(128) select_kernel(SelectArgs a) {
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= a.B) return;
    const int r0 = a.offsets[b], M = a.offsets[b + 1] - r0;
    const float *det = a.det + (size_t)r0 * 5;
    const float w = (float)a.frames[b].w, h = (float)a.frames[b].h;
    int box = -1, kp = -1;
    if (a.enroll) {   // get_biggest_area_face (:28-53): nothing without key points
        if (a.lmk) {
            float best = 0.0f;
            int bi = -1;
            for (int i = lane; i < M; i += 32) {
                const float *d = det + 5 * i;
                const float area = (d[2] - d[0]) * (d[3] - d[1]);
                if (area > best) { best = area; bi = i; }   // strict: a lane keeps its first maximum
            }
            warp_argmax_first(best, bi);
            box = kp = bi;
        }
    } else {
        const float mcl = a.mcl_ratio * w, mcr = a.mcr_ratio * w;          // :103-104
        const float me = fminf(50.0f, a.me_ratio * w);                      // :105-106
        const float x_cen = w / 2.0f, hw = h * w, w_me = w - me, h_me = h - me;
        int n_valid = 0, n_center = 0;
        for (int i = lane; i < M; i += 32) {                                // :110-135
            const float *d = det + 5 * i;
            const float dx = d[2] - d[0];
            const float area = dx * dx;                                     // the reference squares the width (:115)
            const float cx = (d[0] + d[2]) / 2.0f, cy = (d[1] + d[3]) / 2.0f;
            if (cx >= me && cx <= w_me && cy >= me && cy <= h_me && area / hw >= a.min_ratio) {
                ++n_valid;
                const float dc = cx - x_cen;
                if (-mcl <= dc && dc <= mcr) ++n_center;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            n_valid += __shfl_xor_sync(0xffffffffu, n_valid, o);
            n_center += __shfl_xor_sync(0xffffffffu, n_center, o);
        }
        const int mode = n_center > 0 ? 2 : (n_valid > 0 ? 1 : 0);         // :137-143
        float best = 0.0f;
        int bi = -1;
        for (int i = lane; i < M; i += 32) {                                // :145-153
            const float *d = det + 5 * i;
            bool in = true;
            if (mode) {
                const float dx = d[2] - d[0];
                const float area = dx * dx;
                const float cx = (d[0] + d[2]) / 2.0f, cy = (d[1] + d[3]) / 2.0f;
                in = cx >= me && cx <= w_me && cy >= me && cy <= h_me && area / hw >= a.min_ratio;
                if (in && mode == 2) {
                    const float dc = cx - x_cen;
                    in = -mcl <= dc && dc <= mcr;
                }
            }
            const float tem = (d[2] - d[0]) + (d[3] - d[1]);
            if (in && tem > best) { best = tem; bi = i; }
        }
        warp_argmax_first(best, bi);
        box = bi;
        if (box >= 0 && a.lmk) {                                            // :158-181
            const float *o = det + 5 * box;
            const float ox0 = o[0], oy0 = o[1], ox1 = o[2], oy1 = o[3];
            int first = 0x7fffffff;
            for (int i = lane; i < M && first == 0x7fffffff; i += 32) {
                const float *d = det + 5 * i;
                if (fabsf(ox0 - d[0]) <= 2.0f && fabsf(oy0 - d[1]) <= 2.0f && fabsf(ox1 - d[2]) <= 2.0f && fabsf(oy1 - d[3]) <= 2.0f) first = i;
            }
#pragma unroll
            for (int o2 = 16; o2 > 0; o2 >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o2));
            kp = first == 0x7fffffff ? -1 : first;
        }
    }
    if (lane == 0) {
        a.sel[2 * b] = box >= 0 ? r0 + box : -1;
        a.sel[2 * b + 1] = kp >= 0 ? r0 + kp : -1;
        if (a.sel_frame_idx) a.sel_frame_idx[b] = b;
    }
    if (a.sel_lmk && lane < 10)
        a.sel_lmk[(size_t)b * 10 + lane] = (kp >= 0 && box >= 0) ? a.lmk[(size_t)(r0 + kp) * 10 + lane] : __int_as_float(0x7fc00000);
}

int select_launch(fd_ctx *ctx, const int *offsets_dev, const float *det_dev, const float *lmk_dev, const FrameDev *frames_dev, int B,
                  const fd_select_params *p, int is_enroll, int *sel_dev, float *sel_lmk_dev, int *sel_frame_idx_dev) {
    SelectArgs a;
    a.offsets = offsets_dev;
    a.det = det_dev;
    a.lmk = lmk_dev;
    a.frames = frames_dev;
    a.B = B;
    a.mcl_ratio = p->margin_center_left_ratio;
    a.mcr_ratio = p->margin_center_right_ratio;
    a.me_ratio = p->margin_edge_ratio;
    a.min_ratio = p->minimum_face_ratio;
    a.enroll = is_enroll ? 1 : 0;
    a.sel = sel_dev;
    a.sel_lmk = sel_lmk_dev;
    a.sel_frame_idx = sel_frame_idx_dev;
    select_kernel<<<(B + 3) / 4, 128, 0, ctx->stream>>>(a);
    FD_LAUNCH_CHECK_NAMED(ctx, "select_kernel");
    return FD_OK;
}

}  // namespace fd
