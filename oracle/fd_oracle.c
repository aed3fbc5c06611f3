/*
 * fd_oracle.c — CPU ORACLE for the rs-face-detection hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is a plain-C restatement of the reference's algorithm for the path
 * preprocess -> RetinaFace head decode -> NMS -> rescale -> 5-point align/warp.
 * It is the parity checker for the CUDA library in rs_face_detection_b200/csrc and the
 * CPU baseline timed by bench.py.  Nothing in the product path links, imports or calls it:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * Build:  gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC (see oracle/Makefile).
 * -ffp-contract=off matters: the reference (Rust, no fast-math) never fuses a*b+c.
 *
 * Parity pin status (see DESIGN.md "Oracle"):
 *   - The reference's own tests hold NO expected values (every #[test] only println!s), and the Rust
 *     crate cannot be compiled here (no cargo/rustc).  The Rust-side functions are therefore restated
 *     line by line and pinned on the reference's test INPUTS with hand-derived answers
 *     (tests/test_oracle_known_answers.py) and, independently, against numpy / Python float32
 *     restatements written from the Rust sources (tests/test_oracle_second_pin.py: nms, argsort,
 *     clip, bbox / landmark pred, the whole post-CNN half of _forward + _postprocess, the letterbox
 *     geometry, the tensor normalisation, FaceSelection, the model preprocessors)  ->  still
 *     "parity unpinned by reference tests" for those: no output of the reference itself exists.
 *   - The three OpenCV calls (resize, estimateAffinePartial2D, warpAffine; opencv crate 0.92.0,
 *     Cargo.lock:1175) are restated from OpenCV's published algorithms and PINNED against this
 *     container's cv2 4.13.0 (tests/test_oracle_vs_cv2.py + committed fixtures in tests/golden/).
 *
 * All citations are relative to /root/reference/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include <limits.h>

#define FDO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------ */
/* a3. base anchors — src/processing/generate_anchors.rs:20-157                                */
/* ------------------------------------------------------------------------------------------ */

/* generate_anchors.rs:20-26 */
static void whctrs(const float a[4], float *w, float *h, float *xc, float *yc) {
    *w = a[2] - a[0] + 1.0f;
    *h = a[3] - a[1] + 1.0f;
    *xc = a[0] + 0.5f * (*w - 1.0f);
    *yc = a[1] + 0.5f * (*h - 1.0f);
}
/* generate_anchors.rs:28-39 */
static void mkanchor(float ws, float hs, float xc, float yc, float out[4]) {
    out[0] = xc - 0.5f * (ws - 1.0f);
    out[1] = yc - 0.5f * (hs - 1.0f);
    out[2] = xc + 0.5f * (ws - 1.0f);
    out[3] = yc + 0.5f * (hs - 1.0f);
}
/* Rust f32::round = half away from zero = C roundf (generate_anchors.rs:145) */

/* generate_anchors.rs:141-148.  NB hs = ws*ratio is NOT rounded (:146). out: n_ratios x 4 */
FDO_API void fdo_ratio_enum(const float anchor[4], const float *ratios, int n_ratios, float *out) {
    float w, h, xc, yc;
    whctrs(anchor, &w, &h, &xc, &yc);
    float size = w * h;
    for (int i = 0; i < n_ratios; ++i) {
        float size_ratio = size / ratios[i];
        float ws = roundf(sqrtf(size_ratio));
        float hs = ws * ratios[i];
        mkanchor(ws, hs, xc, yc, out + 4 * i);
    }
}
/* generate_anchors.rs:151-157. out: n_scales x 4 */
FDO_API void fdo_scale_enum(const float anchor[4], const float *scales, int n_scales, float *out) {
    float w, h, xc, yc;
    whctrs(anchor, &w, &h, &xc, &yc);
    for (int i = 0; i < n_scales; ++i) mkanchor(w * scales[i], h * scales[i], xc, yc, out + 4 * i);
}
/* generate_anchors.rs:41-59 (generate_anchors) and :61-93 (generate_anchors2).
 * out: (n_ratios*n_scales*(dense?2:1)) x 4, ratio-major. Returns the row count. */
FDO_API int fdo_generate_anchors2(int base_size, const float *ratios, int n_ratios, const float *scales,
                                  int n_scales, int stride, int dense_anchor, float *out) {
    float base[4] = {1.0f - 1.0f, 1.0f - 1.0f, (float)base_size - 1.0f, (float)base_size - 1.0f};
    float *ra = (float *)malloc(sizeof(float) * 4 * (size_t)n_ratios);
    fdo_ratio_enum(base, ratios, n_ratios, ra);
    int n = 0;
    for (int r = 0; r < n_ratios; ++r) {
        fdo_scale_enum(ra + 4 * r, scales, n_scales, out + 4 * n);
        n += n_scales;
    }
    free(ra);
    if (dense_anchor) { /* :80-90 */
        for (int i = 0; i < n * 4; ++i) out[n * 4 + i] = out[i] + (float)stride / 2.0f;
        n *= 2;
    }
    return n;
}
FDO_API int fdo_generate_anchors(int base_size, const float *ratios, int n_ratios, const float *scales,
                                 int n_scales, float *out) {
    return fdo_generate_anchors2(base_size, ratios, n_ratios, scales, n_scales, 0, 0, out);
}
/* generate_anchors.rs:95-114: level i uses the single ratio[i], scale[i]. out: n_levels x 4 */
FDO_API void fdo_generate_anchors_fpn(const int *base_size, const float *ratios, const float *scales,
                                      int n_levels, float *out) {
    for (int i = 0; i < n_levels; ++i)
        fdo_generate_anchors(base_size[i], ratios + i, 1, scales + i, 1, out + 4 * i);
}
/* generate_anchors.rs:116-138 with the RetinaFace cfg built at face_detection.rs:55-80:
 * strides sorted descending (32,16,8), base_size 16, ratio {1}, scales {32,16},{8,4},{2,1}.
 * out: 3 x 2 x 4. */
FDO_API void fdo_generate_anchors_fpn2_retinaface(int dense_anchor, float *out) {
    const int strides[3] = {32, 16, 8};
    const float scales[3][2] = {{32.0f, 16.0f}, {8.0f, 4.0f}, {2.0f, 1.0f}};
    const float ratio = 1.0f;
    float tmp[16];
    for (int s = 0; s < 3; ++s) {
        int n = fdo_generate_anchors2(16, &ratio, 1, scales[s], 2, strides[s], dense_anchor, tmp);
        (void)n;
        memcpy(out + 8 * s, tmp, sizeof(float) * 8);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* a4. anchor plane — src/rcnn/anchors.rs:3-21.  out (H,W,A,4) row-major                        */
/* ------------------------------------------------------------------------------------------ */
FDO_API void fdo_anchors_plane(int height, int width, int stride, const float *base, int A, float *out) {
    for (int iw = 0; iw < width; ++iw) {
        float sw = (float)(iw * stride);
        for (int ih = 0; ih < height; ++ih) {
            float sh = (float)(ih * stride);
            for (int k = 0; k < A; ++k) {
                float *o = out + (((size_t)ih * width + iw) * A + k) * 4;
                o[0] = base[k * 4 + 0] + sw;
                o[1] = base[k * 4 + 1] + sh;
                o[2] = base[k * 4 + 2] + sw;
                o[3] = base[k * 4 + 3] + sh;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* a6/a7/a9 + siblings — face_detection.rs:516-570, processing/bbox_transform.rs               */
/* ------------------------------------------------------------------------------------------ */

/* nonlinear_pred (bbox_transform.rs:90-120) == RetinaFaceDetection::bbox_pred (face_detection.rs:516-549)
 * for ncols==4.  ncols may be a multiple of 4 (class-wise deltas, :100-103 stride-4 slices).        */
FDO_API void fdo_nonlinear_pred(const float *boxes, const float *deltas, int n, int ncols, float *out) {
    for (int i = 0; i < n; ++i) {
        const float *b = boxes + 4 * (size_t)i;
        float w = b[2] - b[0] + 1.0f;
        float h = b[3] - b[1] + 1.0f;
        float cx = b[0] + 0.5f * (w - 1.0f);
        float cy = b[1] + 0.5f * (h - 1.0f);
        for (int j = 0; j + 3 < ncols; j += 4) {
            const float *d = deltas + (size_t)i * ncols + j;
            float *o = out + (size_t)i * ncols + j;
            float pcx = d[0] * w + cx;
            float pcy = d[1] * h + cy;
            float pw = expf(d[2]) * w;
            float ph = expf(d[3]) * h;
            o[0] = pcx - 0.5f * (pw - 1.0f);
            o[1] = pcy - 0.5f * (ph - 1.0f);
            o[2] = pcx + 0.5f * (pw - 1.0f);
            o[3] = pcy + 0.5f * (ph - 1.0f);
        }
    }
}
/* face_detection.rs:516-549: first 4 columns regressed, columns >4 copied through (:544-546) */
FDO_API void fdo_bbox_pred(const float *boxes, const float *deltas, int n, int ncols, float *out) {
    for (int i = 0; i < n; ++i) {
        fdo_nonlinear_pred(boxes + 4 * (size_t)i, deltas + (size_t)i * ncols, 1, 4, out + (size_t)i * ncols);
        for (int j = 4; j < ncols; ++j) out[(size_t)i * ncols + j] = deltas[(size_t)i * ncols + j];
    }
}
/* landmark_pred: face_detection.rs:551-570 ((N,5,2)) == bbox_transform.rs:123-160 ((N,10)) — same memory layout */
FDO_API void fdo_landmark_pred(const float *boxes, const float *deltas, int n, float *out) {
    for (int i = 0; i < n; ++i) {
        const float *b = boxes + 4 * (size_t)i;
        float w = b[2] - b[0] + 1.0f;
        float h = b[3] - b[1] + 1.0f;
        float cx = b[0] + 0.5f * (w - 1.0f);
        float cy = b[1] + 0.5f * (h - 1.0f);
        for (int p = 0; p < 5; ++p) {
            out[(size_t)i * 10 + 2 * p + 0] = deltas[(size_t)i * 10 + 2 * p + 0] * w + cx;
            out[(size_t)i * 10 + 2 * p + 1] = deltas[(size_t)i * 10 + 2 * p + 1] * h + cy;
        }
    }
}
/* Rust f32::min/max ignore NaN like C fminf/fmaxf.  bbox_transform.rs:27-45 */
FDO_API void fdo_clip_boxes(float *boxes, int rows, int cols, int im_h, int im_w) {
    float width = (float)im_w - 1.0f, height = (float)im_h - 1.0f;
    for (int i = 0; i < rows; ++i)
        for (int j = 0; j + 3 < cols; j += 4) {
            float *b = boxes + (size_t)i * cols + j;
            b[0] = fmaxf(fminf(b[0], width), 0.0f);
            b[1] = fmaxf(fminf(b[1], height), 0.0f);
            b[2] = fmaxf(fminf(b[2], width), 0.0f);
            b[3] = fmaxf(fminf(b[3], height), 0.0f);
        }
}
/* bbox_transform.rs:47-65 */
FDO_API void fdo_clip_points(float *pts, int rows, int cols, int im_h, int im_w) {
    float width = (float)im_w - 1.0f, height = (float)im_h - 1.0f;
    for (int i = 0; i < rows; ++i)
        for (int j = 0; j + 9 < cols; j += 10) {
            float *p = pts + (size_t)i * cols + j;
            for (int k = 0; k < 10; k += 2) p[k] = fmaxf(fminf(p[k], width), 0.0f);
            for (int k = 1; k < 10; k += 2) p[k] = fmaxf(fminf(p[k], height), 0.0f);
        }
}
/* bbox_transform.rs:162-186 */
FDO_API void fdo_iou_pred(const float *boxes, const float *deltas, int n, int ncols, int num_classes, float *out) {
    memset(out, 0, sizeof(float) * (size_t)n * ncols);
    for (int i = 0; i < n; ++i)
        for (int c = 0; c < num_classes; ++c)
            for (int k = 0; k < 4; ++k)
                out[(size_t)i * ncols + 4 * c + k] = deltas[(size_t)i * ncols + 4 * c + k] + boxes[4 * (size_t)i + k];
}
/* bbox_transform.rs:67-88 (1e-14 is added in f32: a no-op for |w|>=1e-7) */
FDO_API void fdo_nonlinear_transform(const float *ex, const float *gt, int n, float *out) {
    for (int i = 0; i < n; ++i) {
        const float *e = ex + 4 * (size_t)i, *g = gt + 4 * (size_t)i;
        float ew = e[2] - e[0] + 1.0f, eh = e[3] - e[1] + 1.0f;
        float ecx = e[0] + 0.5f * (ew - 1.0f), ecy = e[1] + 0.5f * (eh - 1.0f);
        float gw = g[2] - g[0] + 1.0f, gh = g[3] - g[1] + 1.0f;
        float gcx = g[0] + 0.5f * (gw - 1.0f), gcy = g[1] + 0.5f * (gh - 1.0f);
        out[4 * (size_t)i + 0] = (gcx - ecx) / (ew + 1e-14f);
        out[4 * (size_t)i + 1] = (gcy - ecy) / (eh + 1e-14f);
        out[4 * (size_t)i + 2] = logf(gw / ew);
        out[4 * (size_t)i + 3] = logf(gh / eh);
    }
}
/* rcnn/bbox.rs:4-30 and bbox_transform.rs:2-24 (identical results; `_py` clamps iw/ih with max(0) first) */
FDO_API void fdo_bbox_overlaps(const float *boxes, int n, const float *query, int k, float *out) {
    memset(out, 0, sizeof(float) * (size_t)n * k);
    for (int q = 0; q < k; ++q) {
        const float *qb = query + 4 * (size_t)q;
        float box_area = (qb[2] - qb[0] + 1.0f) * (qb[3] - qb[1] + 1.0f);
        for (int i = 0; i < n; ++i) {
            const float *b = boxes + 4 * (size_t)i;
            float iw = fminf(b[2], qb[2]) - fmaxf(b[0], qb[0]) + 1.0f;
            if (iw > 0.0f) {
                float ih = fminf(b[3], qb[3]) - fmaxf(b[1], qb[1]) + 1.0f;
                if (ih > 0.0f) {
                    float ua = (b[2] - b[0] + 1.0f) * (b[3] - b[1] + 1.0f) + box_area - iw * ih;
                    out[(size_t)i * k + q] = iw * ih / ua;
                }
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* a10. stable descending argsort — utils/utils.rs:87-95, nms.rs:5-6                            */
/* Rust sort_by is a stable merge sort; ties keep ascending index. NaN => -1 (the reference      */
/* panics in argsort_descending via unwrap(); nms.rs treats incomparable as Equal => unspecified)*/
/* ------------------------------------------------------------------------------------------ */
static void merge_sort_desc(const float *s, int32_t *idx, int32_t *tmp, int n) {
    for (int width = 1; width < n; width *= 2) {
        for (int lo = 0; lo < n; lo += 2 * width) {
            int mid = lo + width < n ? lo + width : n;
            int hi = lo + 2 * width < n ? lo + 2 * width : n;
            int i = lo, j = mid, k = lo;
            while (i < mid && j < hi) {
                /* take right only if strictly greater score => stable */
                if (s[idx[j]] > s[idx[i]]) tmp[k++] = idx[j++];
                else tmp[k++] = idx[i++];
            }
            while (i < mid) tmp[k++] = idx[i++];
            while (j < hi) tmp[k++] = idx[j++];
        }
        memcpy(idx, tmp, sizeof(int32_t) * (size_t)n);
    }
}
FDO_API int fdo_argsort_descending(const float *scores, int n, int32_t *order) {
    for (int i = 0; i < n; ++i) {
        if (isnan(scores[i])) return -1;
        order[i] = i;
    }
    if (n > 1) {
        int32_t *tmp = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
        merge_sort_desc(scores, order, tmp, n);
        free(tmp);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* a11. greedy IoU NMS — src/processing/nms.rs:3-65 (THE NMS the pipeline uses)                 */
/* dets (n,5) row-major [x1,y1,x2,y2,score]; keep receives indices into dets in pick order.     */
/* Survivor test is `ovr <= thresh` (:58) so a NaN overlap REMOVES the box.                     */
/* The structure (O(kept * remaining), remaining list rebuilt per pick) is the reference's.     */
/* ------------------------------------------------------------------------------------------ */
FDO_API int fdo_nms(const float *dets, int n, float thresh, int32_t *keep) {
    if (n <= 0) return 0;
    float *sc = (float *)malloc(sizeof(float) * (size_t)n);
    int32_t *order = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    for (int i = 0; i < n; ++i) sc[i] = dets[5 * (size_t)i + 4];
    if (fdo_argsort_descending(sc, n, order) != 0) { free(sc); free(order); return -1; }
    int len = n, nkeep = 0;
    while (len > 0) {
        int i = order[0];
        keep[nkeep++] = i;
        const float *bi = dets + 5 * (size_t)i;
        float area_i = (bi[2] - bi[0] + 1.0f) * (bi[3] - bi[1] + 1.0f);
        int out = 0;
        for (int t = 1; t < len; ++t) {
            int j = order[t];
            const float *bj = dets + 5 * (size_t)j;
            float xx1 = fmaxf(bi[0], bj[0]);
            float yy1 = fmaxf(bi[1], bj[1]);
            float xx2 = fminf(bi[2], bj[2]);
            float yy2 = fminf(bi[3], bj[3]);
            float w = fmaxf(0.0f, xx2 - xx1 + 1.0f);
            float h = fmaxf(0.0f, yy2 - yy1 + 1.0f);
            float inter = w * h;
            float area_j = (bj[2] - bj[0] + 1.0f) * (bj[3] - bj[1] + 1.0f);
            float ovr = inter / (area_i + area_j - inter);
            if (ovr <= thresh) order[out++] = j;
        }
        len = out;
    }
    free(sc); free(order);
    return nkeep;
}
/* Number of IoU pairs the greedy loop above evaluates (the algorithmic work for the NMS roofline). */
FDO_API long long fdo_nms_pairs(const float *dets, int n, float thresh) {
    if (n <= 0) return 0;
    float *sc = (float *)malloc(sizeof(float) * (size_t)n);
    int32_t *order = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    for (int i = 0; i < n; ++i) sc[i] = dets[5 * (size_t)i + 4];
    if (fdo_argsort_descending(sc, n, order) != 0) { free(sc); free(order); return -1; }
    int len = n; long long pairs = 0;
    while (len > 0) {
        int i = order[0];
        const float *bi = dets + 5 * (size_t)i;
        float area_i = (bi[2] - bi[0] + 1.0f) * (bi[3] - bi[1] + 1.0f);
        int out = 0;
        pairs += len - 1;
        for (int t = 1; t < len; ++t) {
            int j = order[t];
            const float *bj = dets + 5 * (size_t)j;
            float w = fmaxf(0.0f, fminf(bi[2], bj[2]) - fmaxf(bi[0], bj[0]) + 1.0f);
            float h = fmaxf(0.0f, fminf(bi[3], bj[3]) - fmaxf(bi[1], bj[1]) + 1.0f);
            float inter = w * h;
            float area_j = (bj[2] - bj[0] + 1.0f) * (bj[3] - bj[1] + 1.0f);
            float ovr = inter / (area_i + area_j - inter);
            if (ovr <= thresh) order[out++] = j;
        }
        len = out;
    }
    free(sc); free(order);
    return pairs;
}
/* Variant: src/rcnn/cpu_nms.rs:10-55 — suppresses on `ovr >= thresh` (:48), including itself-vs-itself
 * (harmless: already kept). The reference uses sort_unstable_by; ties are therefore unspecified there,
 * the oracle uses the stable order. */
FDO_API int fdo_cpu_nms(const float *dets, int n, float thresh, int32_t *keep) {
    if (n <= 0) return 0;
    float *sc = (float *)malloc(sizeof(float) * (size_t)n);
    float *areas = (float *)malloc(sizeof(float) * (size_t)n);
    int32_t *order = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    uint8_t *sup = (uint8_t *)calloc((size_t)n, 1);
    for (int i = 0; i < n; ++i) {
        const float *b = dets + 5 * (size_t)i;
        sc[i] = b[4];
        areas[i] = (b[2] - b[0] + 1.0f) * (b[3] - b[1] + 1.0f);
    }
    if (fdo_argsort_descending(sc, n, order) != 0) { free(sc); free(areas); free(order); free(sup); return -1; }
    int nkeep = 0;
    for (int a = 0; a < n; ++a) {
        int i = order[a];
        if (sup[i]) continue;
        keep[nkeep++] = i;
        const float *bi = dets + 5 * (size_t)i;
        for (int b = 0; b < n; ++b) {
            int j = order[b];
            if (sup[j]) continue;
            const float *bj = dets + 5 * (size_t)j;
            float w = fmaxf(fminf(bi[2], bj[2]) - fmaxf(bi[0], bj[0]) + 1.0f, 0.0f);
            float h = fmaxf(fminf(bi[3], bj[3]) - fmaxf(bi[1], bj[1]) + 1.0f, 0.0f);
            float inter = w * h;
            float ovr = inter / (areas[i] + areas[j] - inter);
            if (ovr >= thresh) sup[j] = 1;
        }
    }
    free(sc); free(areas); free(order); free(sup);
    return nkeep;
}
/* Contract of the reference's C symbol `_nms` (src/gpu_nms.hpp:6-8, nms_kernel.cu:91-144): boxes are
 * ALREADY sorted by score descending; keep = indices into that sorted array; suppress on ovr > thresh. */
FDO_API int fdo_nms_sorted(const float *boxes, int n, int boxes_dim, float thresh, int32_t *keep) {
    uint8_t *sup = (uint8_t *)calloc((size_t)(n > 0 ? n : 1), 1);
    int nkeep = 0;
    for (int i = 0; i < n; ++i) {
        if (sup[i]) continue;
        keep[nkeep++] = i;
        const float *bi = boxes + (size_t)boxes_dim * i;
        float area_i = (bi[2] - bi[0] + 1.0f) * (bi[3] - bi[1] + 1.0f);
        for (int j = i + 1; j < n; ++j) {
            if (sup[j]) continue;
            const float *bj = boxes + (size_t)boxes_dim * j;
            float w = fmaxf(0.0f, fminf(bi[2], bj[2]) - fmaxf(bi[0], bj[0]) + 1.0f);
            float h = fmaxf(0.0f, fminf(bi[3], bj[3]) - fmaxf(bi[1], bj[1]) + 1.0f);
            float inter = w * h;
            float area_j = (bj[2] - bj[0] + 1.0f) * (bj[3] - bj[1] + 1.0f);
            float ovr = inter / (area_i + area_j - inter);
            if (!(ovr <= thresh)) sup[j] = 1;
        }
    }
    free(sup);
    return nkeep;
}

/* ------------------------------------------------------------------------------------------ */
/* §8c-R  cv::resize INTER_LINEAR 8UC3 (called at face_detection.rs:156, face_alignment.rs:98)   */
/* OpenCV imgproc/resize.cpp: 11-bit coefficient tables, HResizeLinear (int32 rows) +            */
/* VResizeLinear<uchar,int,short> ((b*(S>>4))>>16 ... +2)>>2.  Pinned against cv2 4.13.0.       */
/* ------------------------------------------------------------------------------------------ */
static inline int cv_round_f(float v) { return (int)lrintf(v); }   /* round-half-even (default FE mode) */
static inline short sat_short(int v) { return (short)(v < -32768 ? -32768 : (v > 32767 ? 32767 : v)); }

FDO_API void fdo_resize_linear_u8c3(const uint8_t *src, int sh, int sw, int spitch,
                                    uint8_t *dst, int dh, int dw, int dpitch) {
    if (sh == dh && sw == dw) {
        for (int y = 0; y < sh; ++y) memcpy(dst + (size_t)y * dpitch, src + (size_t)y * spitch, (size_t)sw * 3);
        return;
    }
    double scale_x = 1.0 / ((double)dw / sw), scale_y = 1.0 / ((double)dh / sh);
    int *xofs = (int *)malloc(sizeof(int) * (size_t)dw * 2);
    short *xa = (short *)malloc(sizeof(short) * (size_t)dw * 2);
    for (int dx = 0; dx < dw; ++dx) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = (int)floorf(fx);
        fx -= (float)sx;
        if (sx < 0) { fx = 0.f; sx = 0; }
        if (sx >= sw - 1) { fx = 0.f; sx = sw - 1; }
        xofs[2 * dx] = sx;
        xofs[2 * dx + 1] = sx + 1 < sw ? sx + 1 : sw - 1;
        xa[2 * dx] = sat_short(cv_round_f((1.f - fx) * 2048.f));
        xa[2 * dx + 1] = sat_short(cv_round_f(fx * 2048.f));
    }
    int *rows = (int *)malloc(sizeof(int) * (size_t)dw * 3 * 2);
    int *r0 = rows, *r1 = rows + (size_t)dw * 3;
    for (int dy = 0; dy < dh; ++dy) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = (int)floorf(fy);
        fy -= (float)sy;
        int b0 = sat_short(cv_round_f((1.f - fy) * 2048.f));
        int b1 = sat_short(cv_round_f(fy * 2048.f));
        int y0 = sy < 0 ? 0 : (sy > sh - 1 ? sh - 1 : sy);
        int y1 = sy + 1 < 0 ? 0 : (sy + 1 > sh - 1 ? sh - 1 : sy + 1);
        const uint8_t *s0 = src + (size_t)y0 * spitch, *s1 = src + (size_t)y1 * spitch;
        for (int dx = 0; dx < dw; ++dx) {
            int x0 = xofs[2 * dx] * 3, x1 = xofs[2 * dx + 1] * 3, a0 = xa[2 * dx], a1 = xa[2 * dx + 1];
            for (int c = 0; c < 3; ++c) {
                r0[dx * 3 + c] = s0[x0 + c] * a0 + s0[x1 + c] * a1;
                r1[dx * 3 + c] = s1[x0 + c] * a0 + s1[x1 + c] * a1;
            }
        }
        uint8_t *d = dst + (size_t)dy * dpitch;
        for (int i = 0; i < dw * 3; ++i)
            d[i] = (uint8_t)((((b0 * (r0[i] >> 4)) >> 16) + ((b1 * (r1[i] >> 4)) >> 16) + 2) >> 2);
    }
    free(xofs); free(xa); free(rows);
}

/* ------------------------------------------------------------------------------------------ */
/* a1. letterbox geometry — face_detection.rs:140-153 (f32 maths, `as i32` truncation)          */
/* ------------------------------------------------------------------------------------------ */
FDO_API void fdo_letterbox_geometry(int img_h, int img_w, int size_w, int size_h, int *new_w, int *new_h, float *det_scale) {
    float im_ratio = (float)img_h / (float)img_w;
    float model_ratio = (float)size_h / (float)size_w;
    if (im_ratio > model_ratio) {
        *new_h = size_h;
        *new_w = (int)((float)(*new_h) / im_ratio);
    } else {
        *new_w = size_w;
        *new_h = (int)((float)(*new_w) * im_ratio);
    }
    *det_scale = (float)(*new_h) / (float)img_h;
}
/* a1: resize + zero canvas + top-left ROI copy (face_detection.rs:156-188). det_img: size_h x size_w x 3 */
FDO_API float fdo_preprocess_letterbox(const uint8_t *img, int h, int w, int pitch, int size_w, int size_h, uint8_t *det_img) {
    int nw, nh; float det_scale;
    fdo_letterbox_geometry(h, w, size_w, size_h, &nw, &nh, &det_scale);
    memset(det_img, 0, (size_t)size_h * size_w * 3);
    if (nw > 0 && nh > 0) fdo_resize_linear_u8c3(img, h, w, pitch, det_img, nh, nw, size_w * 3);
    return det_scale;
}
/* a2: u8 HWC BGR -> f32 NCHW RGB (face_detection.rs:220-230): t[i,y,x] = (px[2-i]/scale - means[2-i]) / stds[2-i] */
FDO_API void fdo_to_tensor(const uint8_t *det_img, int rows, int cols, float pixel_scale, const float means[3],
                           const float stds[3], float *out) {
    for (int i = 0; i < 3; ++i)
        for (int y = 0; y < rows; ++y)
            for (int x = 0; x < cols; ++x) {
                uint8_t p = det_img[((size_t)y * cols + x) * 3 + (2 - i)];
                out[((size_t)i * rows + y) * cols + x] = ((float)p / pixel_scale - means[2 - i]) / stds[2 - i];
            }
}

/* ------------------------------------------------------------------------------------------ */
/* a5..a13  RetinaFaceDetection::_forward (post-CNN half) + _postprocess                        */
/* face_detection.rs:319-470, 473-493                                                           */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int image_w, image_h;         /* padded detector input: clip range (face_detection.rs:373) */
    int n_strides;
    int strides[8];
    int num_anchors;              /* A */
    float base_anchors[8][4][4];  /* [stride][a][4] */
    float bbox_stds[4];
    float landmark_std;
    float conf_thr, iou_thr;
} fdo_det_cfg;

/* heads: for stride s: heads[3s] = scores (2A,H,W), heads[3s+1] = bbox (4A,H,W), heads[3s+2] = lmk (10A,H,W)
 * (one image).  det_out (cap,5), lmk_out (cap,10).  Returns M (>=0) or -1 on NaN score.
 * If n_candidates != NULL it receives K (pre-NMS candidates).  Structure deliberately mirrors the
 * reference: anchors regenerated per call (:329), every anchor decoded before selection (:372-379). */
FDO_API int fdo_detect_post(const fdo_det_cfg *cfg, const float *const *heads, const int *fh, const int *fw,
                            float det_scale, float *det_out, float *lmk_out, int cap, int *n_candidates) {
    int A = cfg->num_anchors;
    size_t total = 0;
    for (int s = 0; s < cfg->n_strides; ++s) total += (size_t)fh[s] * fw[s] * A;
    float *prop = (float *)malloc(sizeof(float) * total * 4);
    float *score = (float *)malloc(sizeof(float) * total);
    float *lmk = (float *)malloc(sizeof(float) * total * 10);
    size_t K = 0;
    for (int s = 0; s < cfg->n_strides; ++s) {
        int H = fh[s], W = fw[s], stride = cfg->strides[s];
        size_t n = (size_t)H * W * A, hw = (size_t)H * W;
        const float *sc = heads[3 * s], *bb = heads[3 * s + 1], *lm = heads[3 * s + 2];
        float *anch = (float *)malloc(sizeof(float) * n * 4);
        float *deltas = (float *)malloc(sizeof(float) * n * 4);
        float *ldel = (float *)malloc(sizeof(float) * n * 10);
        float *boxes = (float *)malloc(sizeof(float) * n * 4);
        float *lpred = (float *)malloc(sizeof(float) * n * 10);
        float *sflat = (float *)malloc(sizeof(float) * n);
        fdo_anchors_plane(H, W, stride, &cfg->base_anchors[s][0][0], A, anch);           /* :329 */
        for (int h = 0; h < H; ++h)
            for (int w = 0; w < W; ++w)
                for (int a = 0; a < A; ++a) {
                    size_t r = ((size_t)h * W + w) * A + a, px = (size_t)h * W + w;
                    sflat[r] = sc[(size_t)(A + a) * hw + px];                             /* :322,336-349 */
                    for (int k = 0; k < 4; ++k)
                        deltas[r * 4 + k] = bb[(size_t)(4 * a + k) * hw + px] * cfg->bbox_stds[k];   /* :350-371 */
                    for (int k = 0; k < 10; ++k)
                        ldel[r * 10 + k] = lm[(size_t)(10 * a + k) * hw + px] * cfg->landmark_std;   /* :382-398 */
                }
        fdo_bbox_pred(anch, deltas, (int)n, 4, boxes);                                    /* :372 */
        fdo_clip_boxes(boxes, (int)n, 4, cfg->image_h, cfg->image_w);                     /* :373 */
        fdo_landmark_pred(anch, ldel, (int)n, lpred);                                     /* :399 */
        for (size_t r = 0; r < n; ++r)
            if (sflat[r] >= cfg->conf_thr) {                                              /* :375 */
                memcpy(prop + K * 4, boxes + r * 4, sizeof(float) * 4);
                score[K] = sflat[r];
                memcpy(lmk + K * 10, lpred + r * 10, sizeof(float) * 10);
                ++K;
            }
        free(anch); free(deltas); free(ldel); free(boxes); free(lpred); free(sflat);
    }
    if (n_candidates) *n_candidates = (int)K;
    int M = 0;
    if (K > 0) {
        int32_t *order = (int32_t *)malloc(sizeof(int32_t) * K);
        if (fdo_argsort_descending(score, (int)K, order) != 0) {                          /* :423 */
            free(order); free(prop); free(score); free(lmk); return -1;
        }
        float *pre = (float *)malloc(sizeof(float) * K * 5);
        for (size_t i = 0; i < K; ++i) {                                                  /* :424-430 */
            memcpy(pre + i * 5, prop + (size_t)order[i] * 4, sizeof(float) * 4);
            pre[i * 5 + 4] = score[order[i]];
        }
        int32_t *keep = (int32_t *)malloc(sizeof(int32_t) * K);
        int nk = fdo_nms(pre, (int)K, cfg->iou_thr, keep);                                /* :431 */
        for (int i = 0; i < nk && M < cap; ++i, ++M) {                                    /* :432-464 */
            const float *row = pre + (size_t)keep[i] * 5;
            for (int k = 0; k < 4; ++k) det_out[(size_t)M * 5 + k] = row[k] / det_scale;  /* :477-481 */
            det_out[(size_t)M * 5 + 4] = row[4];
            const float *l = lmk + (size_t)order[keep[i]] * 10;
            for (int k = 0; k < 10; ++k) lmk_out[(size_t)M * 10 + k] = l[k] / det_scale;  /* :483 */
        }
        free(order); free(pre); free(keep);
    }
    free(prop); free(score); free(lmk);
    return M;
}

/* Decode only (pre-sort candidate list in concat order 32|16|8), for stage-level parity checks.
 * cand_box (cap,4), cand_score (cap), cand_lmk (cap,10), cand_index (cap) = global anchor index. */
FDO_API int fdo_decode_candidates(const fdo_det_cfg *cfg, const float *const *heads, const int *fh, const int *fw,
                                  float *cand_box, float *cand_score, float *cand_lmk, int32_t *cand_index, int cap) {
    int A = cfg->num_anchors, K = 0; size_t base = 0;
    for (int s = 0; s < cfg->n_strides; ++s) {
        int H = fh[s], W = fw[s], stride = cfg->strides[s];
        size_t hw = (size_t)H * W;
        const float *sc = heads[3 * s], *bb = heads[3 * s + 1], *lm = heads[3 * s + 2];
        for (int h = 0; h < H; ++h)
            for (int w = 0; w < W; ++w)
                for (int a = 0; a < A; ++a) {
                    size_t r = ((size_t)h * W + w) * A + a, px = (size_t)h * W + w;
                    float scv = sc[(size_t)(A + a) * hw + px];
                    if (!(scv >= cfg->conf_thr)) continue;
                    if (K >= cap) return -2;
                    float anc[4], d[4], ld[10], bx[4];
                    const float *ba = &cfg->base_anchors[s][a][0];
                    float sw = (float)(w * stride), sh = (float)(h * stride);
                    anc[0] = ba[0] + sw; anc[1] = ba[1] + sh; anc[2] = ba[2] + sw; anc[3] = ba[3] + sh;
                    for (int k = 0; k < 4; ++k) d[k] = bb[(size_t)(4 * a + k) * hw + px] * cfg->bbox_stds[k];
                    for (int k = 0; k < 10; ++k) ld[k] = lm[(size_t)(10 * a + k) * hw + px] * cfg->landmark_std;
                    fdo_bbox_pred(anc, d, 1, 4, bx);
                    fdo_clip_boxes(bx, 1, 4, cfg->image_h, cfg->image_w);
                    memcpy(cand_box + (size_t)K * 4, bx, sizeof(bx));
                    cand_score[K] = scv;
                    fdo_landmark_pred(anc, ld, 1, cand_lmk + (size_t)K * 10);
                    cand_index[K] = (int32_t)(base + r);
                    ++K;
                }
        base += hw * A;
    }
    return K;
}

/* ------------------------------------------------------------------------------------------ */
/* §8c-E  cv::estimateAffinePartial2D(from,to,LMEDS,3.0,2000,0.99,10) — face_alignment.rs:50-59  */
/* OpenCV calib3d/ptsetreg.cpp: LMeDSPointSetRegistrator + AffinePartial2DEstimatorCallback,     */
/* RNG(-1) reseeded per call, then refinement = least squares over the inliers (LM on a linear   */
/* model converges there; cv2 agrees to ~6e-8).  Pinned against cv2 4.13.0 (tests/golden).       */
/* M (2x3 double, row-major).  inliers[n] optional.  Returns 1 ok, 0 = empty matrix (failure).   */
/* ------------------------------------------------------------------------------------------ */
typedef struct { uint64_t state; } cv_rng;
static inline unsigned cv_rng_next(cv_rng *r) {
    r->state = (uint64_t)(unsigned)r->state * 4164903690U + (unsigned)(r->state >> 32);
    return (unsigned)r->state;
}
static inline int cv_rng_uniform(cv_rng *r, int a, int b) { return a == b ? a : (int)(cv_rng_next(r) % (unsigned)(b - a) + a); }

static int lmeds_niters(double p, double ep, int modelPoints, int maxIters) {
    /* RANSACUpdateNumIters */
    p = p < 0. ? 0. : (p > 1. ? 1. : p);
    ep = ep < 0. ? 0. : (ep > 1. ? 1. : ep);
    double num = 1. - p; if (num < DBL_MIN) num = DBL_MIN;
    double denom = 1. - pow(1. - ep, modelPoints);
    if (denom < DBL_MIN) return 0;
    num = log(num); denom = log(denom);
    return denom >= 0 || -num >= maxIters * (-denom) ? maxIters : (int)lrint(num / denom);
}
static void fit2(const float *f, const float *t, int i0, int i1, double M[6]) {
    /* AffinePartial2DEstimatorCallback::runKernel: exact similarity through 2 correspondences */
    double x1 = f[2 * i0], y1 = f[2 * i0 + 1], x2 = f[2 * i1], y2 = f[2 * i1 + 1];
    double X1 = t[2 * i0], Y1 = t[2 * i0 + 1], X2 = t[2 * i1], Y2 = t[2 * i1 + 1];
    double d = 1. / ((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2));
    double S0 = d * ((X1 - X2) * (x1 - x2) + (Y1 - Y2) * (y1 - y2));
    double S1 = d * ((Y1 - Y2) * (x1 - x2) - (X1 - X2) * (y1 - y2));
    double S2 = d * ((Y1 - Y2) * (x1 * y2 - x2 * y1) - (X1 * y2 - X2 * y1) * (y1 - y2) - (X1 * x2 - X2 * x1) * (x1 - x2));
    double S3 = d * (-(X1 - X2) * (x1 * y2 - x2 * y1) - (Y1 * x2 - Y2 * x1) * (x1 - x2) - (Y1 * y2 - Y2 * y1) * (y1 - y2));
    M[0] = S0; M[1] = -S1; M[2] = S2; M[3] = S1; M[4] = S0; M[5] = S3;
}
static void affine_err(const float *f, const float *t, int n, const double M[6], float *err) {
    /* Affine2DEstimatorCallback::computeError: float32 result of double arithmetic */
    for (int i = 0; i < n; ++i) {
        double a = M[0] * f[2 * i] + M[1] * f[2 * i + 1] + M[2] - t[2 * i];
        double b = M[3] * f[2 * i] + M[4] * f[2 * i + 1] + M[5] - t[2 * i + 1];
        err[i] = (float)(a * a + b * b);
    }
}
static int cmp_float(const void *a, const void *b) { float x = *(const float *)a, y = *(const float *)b; return (x > y) - (x < y); }

/* Least-squares similarity over the points flagged in mask (fixed operation order: the CUDA kernel
 * in csrc/fd_align.cu performs the identical sequence so M is bit-identical on device). */
static void ls_similarity(const float *f, const float *t, int n, const uint8_t *mask, double M[6]) {
    double sx = 0, sy = 0, sX = 0, sY = 0; int c = 0;
    for (int i = 0; i < n; ++i) if (mask[i]) { sx += f[2 * i]; sy += f[2 * i + 1]; sX += t[2 * i]; sY += t[2 * i + 1]; ++c; }
    double mx = sx / c, my = sy / c, mX = sX / c, mY = sY / c;
    double num_a = 0, num_b = 0, den = 0;
    for (int i = 0; i < n; ++i) if (mask[i]) {
        double dx = f[2 * i] - mx, dy = f[2 * i + 1] - my, dX = t[2 * i] - mX, dY = t[2 * i + 1] - mY;
        num_a += dx * dX + dy * dY;
        num_b += dx * dY - dy * dX;
        den += dx * dx + dy * dy;
    }
    double a = num_a / den, b = num_b / den;
    M[0] = a; M[1] = -b; M[2] = mX - (a * mx - b * my);
    M[3] = b; M[4] = a;  M[5] = mY - (b * mx + a * my);
}

FDO_API int fdo_estimate_affine_partial_2d_lmeds(const float *from, const float *to, int n, double M[6], uint8_t *inliers) {
    const int modelPoints = 2;
    const double confidence = 0.99; const int maxIters = 2000;
    uint8_t maskbuf[64]; float errbuf[64], sorted[64];
    if (n < modelPoints || n > 64) { if (inliers) memset(inliers, 0, (size_t)(n > 0 ? n : 0)); return 0; }
    uint8_t *mask = inliers ? inliers : maskbuf;
    double best[6] = {0, 0, 0, 0, 0, 0};
    if (n == modelPoints) {
        fit2(from, to, 0, 1, best);
        for (int i = 0; i < 6; ++i) if (!isfinite(best[i])) return 0;
        memset(mask, 1, (size_t)n);
        memcpy(M, best, sizeof(best));
        /* refinement over 2 points reproduces the exact fit */
        return 1;
    }
    cv_rng rng; rng.state = 0xFFFFFFFFFFFFFFFFULL;
    int niters = lmeds_niters(confidence, 0.45, modelPoints, maxIters);
    if (niters < 3) niters = 3;
    double minMedian = DBL_MAX;
    for (int iter = 0; iter < niters; ++iter) {
        /* RANSACPointSetRegistrator::getSubset: draw modelPoints distinct indices (redraw while
         * the index repeats); checkSubset never rejects a 2-point sample. */
        int idx[2];
        idx[0] = cv_rng_uniform(&rng, 0, n);
        do { idx[1] = cv_rng_uniform(&rng, 0, n); } while (idx[1] == idx[0]);
        double model[6];
        fit2(from, to, idx[0], idx[1], model);
        int finite = 1; for (int k = 0; k < 6; ++k) if (!isfinite(model[k])) finite = 0;
        if (!finite) continue;   /* cv::checkRange on the model: runKernel returns 0 models */
        affine_err(from, to, n, model, errbuf);
        memcpy(sorted, errbuf, sizeof(float) * (size_t)n);
        qsort(sorted, (size_t)n, sizeof(float), cmp_float);
        double median = n % 2 != 0 ? (double)sorted[n / 2] : (double)(sorted[n / 2 - 1] + sorted[n / 2]) * 0.5;
        if (median < minMedian) { minMedian = median; memcpy(best, model, sizeof(best)); }
    }
    if (!(minMedian < DBL_MAX)) return 0;
    double sigma = 2.5 * 1.4826 * (1 + 5. / (n - modelPoints)) * sqrt(minMedian);
    if (sigma < 0.001) sigma = 0.001;
    affine_err(from, to, n, best, errbuf);
    float t = (float)(sigma * sigma);
    int good = 0;
    for (int i = 0; i < n; ++i) { mask[i] = errbuf[i] <= t; good += mask[i]; }
    if (good < modelPoints) return 0;
    ls_similarity(from, to, n, mask, M);
    return 1;
}

/* ------------------------------------------------------------------------------------------ */
/* §8c-W  cv::warpAffine INTER_LINEAR, BORDER_CONSTANT(0), 8UC3 — face_alignment.rs:119-126       */
/* OpenCV imgproc/imgwarp.cpp: M inverted in double, 10-bit fixed-point coordinates (AB_BITS),   */
/* 5-bit sub-pixel (INTER_BITS), 15-bit weights (INTER_REMAP_COEF_BITS).  Pinned against cv2.    */
/* ------------------------------------------------------------------------------------------ */
static inline int sat_int_d(double v) {
    double r = nearbyint(v);
    if (r >= 2147483647.0) return INT_MAX;
    if (r <= -2147483648.0) return INT_MIN;
    if (r != r) return INT_MIN;
    return (int)r;
}
FDO_API void fdo_invert_affine(const double M[6], double iM[6]) {
    double D = M[0] * M[4] - M[1] * M[3];
    D = D != 0 ? 1. / D : 0;
    double A11 = M[4] * D, A22 = M[0] * D;
    iM[0] = A11; iM[1] = M[1] * (-D); iM[3] = M[3] * (-D); iM[4] = A22;
    double b1 = -iM[0] * M[2] - iM[1] * M[5];
    double b2 = -iM[3] * M[2] - iM[4] * M[5];
    iM[2] = b1; iM[5] = b2;
}
FDO_API void fdo_warp_affine_u8c3(const uint8_t *src, int sh, int sw, int spitch, const double M[6],
                                  uint8_t *dst, int dh, int dw, int dpitch) {
    double iM[6];
    fdo_invert_affine(M, iM);
    for (int y = 0; y < dh; ++y) {
        int X0 = sat_int_d((iM[1] * y + iM[2]) * 1024) + 16;
        int Y0 = sat_int_d((iM[4] * y + iM[5]) * 1024) + 16;
        uint8_t *d = dst + (size_t)y * dpitch;
        for (int x = 0; x < dw; ++x) {
            int adelta = sat_int_d(iM[0] * x * 1024), bdelta = sat_int_d(iM[3] * x * 1024);
            int X = (int)((unsigned)X0 + (unsigned)adelta) >> 5, Y = (int)((unsigned)Y0 + (unsigned)bdelta) >> 5;
            int sx = sat_short(X >> 5), sy = sat_short(Y >> 5);
            int ax = X & 31, ay = Y & 31;
            /* BilinearTab_i[ay*32+ax]: float weights *32768, exact (multiples of 32) */
            int w00 = (32 - ay) * (32 - ax) * 32, w01 = (32 - ay) * ax * 32, w10 = ay * (32 - ax) * 32, w11 = ay * ax * 32;
            int in_x0 = sx >= 0 && sx < sw, in_x1 = sx + 1 >= 0 && sx + 1 < sw;
            int in_y0 = sy >= 0 && sy < sh, in_y1 = sy + 1 >= 0 && sy + 1 < sh;
            for (int c = 0; c < 3; ++c) {
                int v00 = in_y0 && in_x0 ? src[(size_t)sy * spitch + sx * 3 + c] : 0;
                int v01 = in_y0 && in_x1 ? src[(size_t)sy * spitch + (sx + 1) * 3 + c] : 0;
                int v10 = in_y1 && in_x0 ? src[(size_t)(sy + 1) * spitch + sx * 3 + c] : 0;
                int v11 = in_y1 && in_x1 ? src[(size_t)(sy + 1) * spitch + (sx + 1) * 3 + c] : 0;
                d[x * 3 + c] = (uint8_t)((v00 * w00 + v01 * w01 + v10 * w10 + v11 * w11 + (1 << 14)) >> 15);
            }
        }
    }
}

/* Rust `f32::max` (returns the other operand when one is NaN) and `as i32` (truncates, saturates, NaN -> 0). */
static inline float rs_f32_max(float a, float b) { return a != a ? b : (b != b ? a : (a > b ? a : b)); }
static inline int rs_f32_as_i32(float v) {
    if (v != v) return 0;
    if (v >= 2147483648.0f) return 2147483647;
    if (v <= -2147483648.0f) return (-2147483647 - 1);
    return (int)v;
}

/* a14 fallback. FaceAlignment::call when the transformation matrix is empty (face_alignment.rs:64-116).
 * bbox = Option<Array1<f32>> (NULL = None, :66-72).  Returns 1 and the resized crop, or 0 where the reference returns Err:
 * Mat::roi (:92-95) rejects a rectangle that is not inside the image, cv::resize (:98) an empty one. */
FDO_API int fdo_align_fallback(const uint8_t *img, int h, int w, int pitch, const float *bbox, int crop_w, int crop_h, uint8_t *crop) {
    float det[4];
    if (!bbox) {                                     /* :67-72 */
        det[0] = (float)w * 0.0625f;
        det[1] = (float)h * 0.0625f;
        det[2] = (float)w - det[0];
        det[3] = (float)h - det[1];
    } else {
        memcpy(det, bbox, sizeof(det));              /* :74 */
    }
    const float margin = 44.0f;                      /* :77 */
    float bb[4];
    bb[0] = rs_f32_max(det[0] - margin / 2.0f, 0.0f);            /* :79 */
    bb[1] = rs_f32_max(det[1] - margin / 2.0f, 0.0f);            /* :80 */
    bb[2] = rs_f32_max(det[2] + margin / 2.0f, (float)w);        /* :81 — `max`, as written */
    bb[3] = rs_f32_max(det[1] + margin / 2.0f, (float)h);        /* :82 — det[1] and `max`, as written */
    const int x0 = rs_f32_as_i32(bb[0]), y0 = rs_f32_as_i32(bb[1]), x1 = rs_f32_as_i32(bb[2]), y1 = rs_f32_as_i32(bb[3]);
    const int width = x1 - x0, height = y1 - y0;     /* :88-89 */
    /* Mat::roi: 0 <= x, 0 <= width, x + width <= cols (same for y); cv::resize: !src.empty() */
    if (x0 < 0 || width < 0 || (long long)x0 + width > w || y0 < 0 || height < 0 || (long long)y0 + height > h) return 0;
    if (width == 0 || height == 0) return 0;
    fdo_resize_linear_u8c3(img + (size_t)y0 * pitch + (size_t)x0 * 3, height, width, pitch, crop, crop_h, crop_w, crop_w * 3);
    return 1;
}

/* a14. FaceAlignment::call (face_alignment.rs:27-141): similarity warp (:50-59, :119-126), or the bbox-crop fallback
 * (:64-116) when the estimate is empty.  Returns the mode: 1 = warp (M_out written), 2 = fallback crop, 0 = the reference
 * returns Err (fallback ROI outside the image).  bbox NULL = None.  (landmarks = None never reaches here: the reference
 * returns Err from estimateAffinePartial2D's assertion on the empty Mat.) */
FDO_API int fdo_align_face(const uint8_t *img, int h, int w, int pitch, const float *bbox, const float lmk[10], const float tmpl[10],
                           int crop_w, int crop_h, uint8_t *crop, double M_out[6]) {
    double M[6];
    if (!fdo_estimate_affine_partial_2d_lmeds(lmk, tmpl, 5, M, NULL))
        return fdo_align_fallback(img, h, w, pitch, bbox, crop_w, crop_h, crop) ? 2 : 0;
    if (M_out) memcpy(M_out, M, sizeof(M));
    fdo_warp_affine_u8c3(img, h, w, pitch, M, crop, crop_h, crop_w, crop_w * 3);
    return 1;
}

/* ---- FaceSelection::call (pipeline/module/face_selection.rs:72-189), SURVEY 8(f) N3 --------------------------------------
 * face_boxes (M,5) rows [x1,y1,x2,y2,score]; has_kps: key_points is Some.  params = {margin_center_left_ratio,
 * margin_center_right_ratio, margin_edge_ratio, minimum_face_ratio} (face_pipeline/config.rs:107-116).
 * Writes the row index of the selected box and of the row whose key points are returned (-1 = None). */
FDO_API void fdo_face_selection(int img_h, int img_w, const float *face_boxes, int M, int has_kps, int is_enroll,
                                const float params[4], int *box_index, int *kp_index) {
    *box_index = -1;
    *kp_index = -1;
    if (is_enroll) { /* get_biggest_area_face (:28-53): nothing is selected without key points */
        if (!has_kps) return;
        float biggest = 0.0f;
        for (int i = 0; i < M; ++i) {
            const float *b = face_boxes + 5 * i;
            volatile float area = (b[2] - b[0]) * (b[3] - b[1]);
            if (area > biggest) {
                biggest = area;
                *box_index = i;
                *kp_index = i;
            }
        }
        return; /* is_face_area_big_enough (:55-70) does not change the result (:84-101) */
    }
    const float w = (float)img_w, h = (float)img_h;
    volatile float mcl = params[0] * w, mcr = params[1] * w; /* :103-104 */
    volatile float me0 = params[2] * w;
    const float me = fminf(50.0f, me0);                       /* :105-106 */
    volatile float x_cen = w / 2.0f;                           /* :108 */
    volatile float hw = h * w;
    volatile float w_me = w - me, h_me = h - me;
    int n_valid = 0, n_center = 0;
    unsigned char *valid = (unsigned char *)calloc((size_t)(M > 0 ? M : 1), 1), *center = (unsigned char *)calloc((size_t)(M > 0 ? M : 1), 1);
    for (int i = 0; i < M; ++i) { /* :110-127 */
        const float *b = face_boxes + 5 * i;
        volatile float dx = b[2] - b[0];
        volatile float area = dx * dx; /* (x_max - x_min) * (x_max - x_min): the reference squares the width (:115) */
        volatile float sx = b[0] + b[2], sy = b[1] + b[3];
        volatile float cx = sx / 2.0f, cy = sy / 2.0f;
        volatile float ratio = area / hw;
        if (cx >= me && cx <= w_me && cy >= me && cy <= h_me && ratio >= params[3]) {
            valid[i] = 1;
            ++n_valid;
            volatile float d = cx - x_cen; /* :129-135 */
            if (-mcl <= d && d <= mcr) {
                center[i] = 1;
                ++n_center;
            }
        }
    }
    const int mode = n_center > 0 ? 2 : (n_valid > 0 ? 1 : 0); /* :137-143: centre boxes, else valid boxes, else every box */
    float max_size = 0.0f;
    for (int i = 0; i < M; ++i) { /* :145-153 */
        if ((mode == 2 && !center[i]) || (mode == 1 && !valid[i])) continue;
        const float *b = face_boxes + 5 * i;
        volatile float ww = b[2] - b[0], hh = b[3] - b[1];
        volatile float tem = ww + hh;
        if (tem > max_size) {
            max_size = tem;
            *box_index = i;
        }
    }
    free(valid);
    free(center);
    if (*box_index < 0) return; /* :154-156 */
    if (has_kps) {            /* :158-181: key points of the FIRST row within 2 px of the selected box */
        const float *o = face_boxes + 5 * (*box_index);
        for (int i = 0; i < M; ++i) {
            const float *b = face_boxes + 5 * i;
            if (fabsf(o[0] - b[0]) <= 2.0f && fabsf(o[1] - b[1]) <= 2.0f && fabsf(o[2] - b[2]) <= 2.0f && fabsf(o[3] - b[3]) <= 2.0f) {
                *kp_index = i;
                break;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* N1: post-align model preprocessors.  FaceExtraction::_preprocess                  */
/* (face_extraction.rs:38-77: mean 127.5, mul 0.0078125), FaceQuality::call (face_quality.rs:43-101: */
/* mean {123.675,116.28,103.53}, mul {0.01712475,0.017507,0.01742919}), FaceQualityAssessment::call  */
/* (face_quality_assessment.rs:48-88: mean 127.5, mul 0.00784313725):                               */
/* cv::resize INTER_LINEAR -> cvtColor BGR2RGB -> (p as f32 - mean[i]) * mul[i] -> NCHW.            */
/* ------------------------------------------------------------------------------------------ */
FDO_API void fdo_model_preprocess(const uint8_t *img, int h, int w, int pitch, int out_h, int out_w,
                                  const float mean_rgb[3], const float mul_rgb[3], float *out) {
    uint8_t *rs = (uint8_t *)malloc((size_t)out_h * out_w * 3);
    fdo_resize_linear_u8c3(img, h, w, pitch, rs, out_h, out_w, out_w * 3);
    for (int i = 0; i < 3; ++i)
        for (int y = 0; y < out_h; ++y)
            for (int x = 0; x < out_w; ++x) {
                uint8_t p = rs[((size_t)y * out_w + x) * 3 + (2 - i)];   /* RGB channel i = BGR channel 2-i */
                out[((size_t)i * out_h + y) * out_w + x] = ((float)p - mean_rgb[i]) * mul_rgb[i];
            }
    free(rs);
}

/* ------------------------------------------------------------------------------------------ */
/* Whole-frame CPU path (bench.py cpu_baseline / --impl reference):                             */
/* preprocess -> tensor -> decode -> sort -> NMS -> rescale -> align every detection.           */
/* ------------------------------------------------------------------------------------------ */
FDO_API int fdo_pipeline_frame(const fdo_det_cfg *cfg, const uint8_t *img, int h, int w, int pitch,
                               const float *const *heads, const int *fh, const int *fw,
                               const float pixel_means[3], const float pixel_stds[3], float pixel_scale,
                               const float tmpl[10], int crop_w, int crop_h,
                               float *tensor_out, float *det_out, float *lmk_out, int cap, uint8_t *crops_out) {
    uint8_t *det_img = (uint8_t *)malloc((size_t)cfg->image_h * cfg->image_w * 3);
    float det_scale = fdo_preprocess_letterbox(img, h, w, pitch, cfg->image_w, cfg->image_h, det_img);
    fdo_to_tensor(det_img, cfg->image_h, cfg->image_w, pixel_scale, pixel_means, pixel_stds, tensor_out);
    free(det_img);
    int M = fdo_detect_post(cfg, heads, fh, fw, det_scale, det_out, lmk_out, cap, NULL);
    if (M < 0) return M;
    for (int i = 0; i < M; ++i) {
        uint8_t *crop = crops_out + (size_t)i * crop_w * crop_h * 3;
        /* FacePipeline::extract passes the face's own box as the fallback bbox (face_pipeline/pipeline.rs:216) */
        if (!fdo_align_face(img, h, w, pitch, det_out + (size_t)i * 5, lmk_out + (size_t)i * 10, tmpl, crop_w, crop_h, crop, NULL))
            memset(crop, 0, (size_t)crop_w * crop_h * 3);
    }
    return M;
}
