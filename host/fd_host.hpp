// fd_host.hpp — C++ host-side mirror of the reference's operator interface over the C ABI (include/fd_b200.h).
// The reference is compiled Rust and no Rust toolchain exists in this image, so the host layer above the C ABI is
// given in C++ with the reference's names, argument meaning and error behaviour (errors -> exceptions, as the Rust
// returns Err).  Header-only; link with -lfd_b200.
#pragma once
#include <array>
#include <cstdint>
#include <functional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include "../include/fd_b200.h"

namespace fd {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error("fd_b200 error " + std::to_string(c) + ": " + m), code(c) {}
};
inline void check(int rc) {
    if (rc != FD_OK) throw Error(rc, fd_last_error());
}

// row-major 2-D / 3-D float arrays standing in for ndarray::Array2<f32> / Array3<f32>
struct Array2 {
    size_t rows = 0, cols = 0;
    std::vector<float> v;
    Array2() = default;
    Array2(size_t r, size_t c) : rows(r), cols(c), v(r * c) {}
    float *data() { return v.data(); }
    const float *data() const { return v.data(); }
};

class Context {
public:
    explicit Context(int device = 0, const fd_config *cfg = nullptr) { check(fd_ctx_create(device, cfg, &ctx_)); }
    ~Context() { fd_ctx_destroy(ctx_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    fd_ctx *get() const { return ctx_; }
private:
    fd_ctx *ctx_ = nullptr;
};

namespace processing {
// processing::nms::nms (src/processing/nms.rs:3)
inline std::vector<size_t> nms(Context &c, const Array2 &dets, float thresh) {
    std::vector<int32_t> keep(dets.rows ? dets.rows : 1);
    int n = 0;
    check(fd_nms(c.get(), dets.data(), (int)dets.rows, thresh, keep.data(), &n));
    return std::vector<size_t>(keep.begin(), keep.begin() + n);
}
// processing::bbox_transform (src/processing/bbox_transform.rs)
inline void clip_boxes(Context &c, Array2 &boxes, std::pair<size_t, size_t> im_shape) {
    check(fd_clip_boxes(c.get(), boxes.data(), (int)boxes.rows, (int)boxes.cols, (int)im_shape.first, (int)im_shape.second));
}
inline void clip_points(Context &c, Array2 &pts, std::pair<size_t, size_t> im_shape) {
    check(fd_clip_points(c.get(), pts.data(), (int)pts.rows, (int)pts.cols, (int)im_shape.first, (int)im_shape.second));
}
inline Array2 nonlinear_pred(Context &c, const Array2 &boxes, const Array2 &deltas) {
    Array2 out(boxes.rows ? deltas.rows : 0, deltas.cols);
    check(fd_nonlinear_pred(c.get(), boxes.data(), deltas.data(), (int)boxes.rows, (int)deltas.cols, out.data()));
    return out;
}
inline Array2 landmark_pred(Context &c, const Array2 &boxes, const Array2 &deltas) {
    Array2 out(boxes.rows ? deltas.rows : 0, deltas.cols);
    check(fd_landmark_pred(c.get(), boxes.data(), deltas.data(), (int)boxes.rows, out.data()));
    return out;
}
inline Array2 bbox_overlaps_py(Context &c, const Array2 &boxes, const Array2 &query) {
    Array2 out(boxes.rows, query.rows);
    check(fd_bbox_overlaps(c.get(), boxes.data(), (int)boxes.rows, query.data(), (int)query.rows, out.data()));
    return out;
}
// processing::generate_anchors (src/processing/generate_anchors.rs)
inline Array2 generate_anchors2(int base_size, const std::vector<float> &ratios, const std::vector<float> &scales, int stride, bool dense) {
    Array2 out(ratios.size() * scales.size() * 2, 4);
    int n = 0;
    check(fd_generate_anchors2(base_size, ratios.data(), (int)ratios.size(), scales.data(), (int)scales.size(), stride, dense, out.data(), &n));
    out.rows = (size_t)n;
    out.v.resize((size_t)n * 4);
    return out;
}
}  // namespace processing

namespace rcnn {
// rcnn::anchors::anchors (src/rcnn/anchors.rs:3) -> (H,W,A,4) flattened
inline std::vector<float> anchors(Context &c, size_t height, size_t width, size_t stride, const Array2 &base) {
    std::vector<float> out(height * width * base.rows * 4);
    check(fd_anchors_plane(c.get(), (int)height, (int)width, (int)stride, base.data(), (int)base.rows, out.data()));
    return out;
}
inline Array2 bbox_overlaps(Context &c, const Array2 &boxes, const Array2 &query) { return processing::bbox_overlaps_py(c, boxes, query); }
}  // namespace rcnn

// A BGR u8 image standing in for opencv::core::Mat (CV_8UC3)
struct Mat {
    const uint8_t *data;
    int rows, cols, step;
};

// pipeline::module::face_detection::RetinaFaceDetection (src/pipeline/module/face_detection.rs:19-571).  The CNN stays
// behind the serving boundary: `infer` receives the (1,3,H,W) tensor and returns the 9 head tensors.
class RetinaFaceDetection {
public:
    using Infer = std::function<std::vector<std::vector<float>>(const std::vector<float> &)>;
    RetinaFaceDetection(Context &c, Infer infer, float confidence_threshold = 0.7f, float iou_threshold = 0.45f)
        : c_(c), infer_(std::move(infer)), conf_(confidence_threshold), iou_(iou_threshold) {
        check(fd_ctx_get_config(c.get(), &cfg_));
        check(fd_ctx_total_anchors(c.get(), &cap_));
    }
    // call(&Mat) -> (det (M,5), landmarks (M,5,2))   (face_detection.rs:496)
    std::pair<Array2, Array2> call(const Mat &image) {
        std::vector<float> tensor((size_t)3 * cfg_.image_h * cfg_.image_w);
        float det_scale = 0.f;
        check(fd_preprocess(c_.get(), image.data, image.rows, image.cols, image.step, tensor.data(), &det_scale));
        auto heads = infer_(tensor);
        std::vector<const float *> hp;
        for (auto &h : heads) hp.push_back(h.data());
        Array2 det((size_t)cap_, 5), lmk((size_t)cap_, 10);
        int n = 0;
        check(fd_detect(c_.get(), hp.data(), (int)hp.size(), det_scale, conf_, iou_, det.data(), lmk.data(), cap_, &n));
        det.rows = lmk.rows = (size_t)n;
        det.v.resize((size_t)n * 5);
        lmk.v.resize((size_t)n * 10);
        return {det, lmk};
    }
private:
    Context &c_;
    Infer infer_;
    float conf_, iou_;
    fd_config cfg_{};
    int32_t cap_ = 0;
};

// pipeline::module::face_alignment::FaceAlignment (src/pipeline/module/face_alignment.rs:14-141)
class FaceAlignment {
public:
    explicit FaceAlignment(Context &c) : c_(c) { check(fd_ctx_get_config(c.get(), &cfg_)); }
    // call(&Mat, bbox (4, or nullptr = None), landmarks (5x2, or nullptr = None)) -> crop (crop_h x crop_w x 3 u8): the
    // similarity warp, or the bbox-crop fallback (:64-116) when the estimate is empty.  Throws where the reference returns Err.
    std::vector<uint8_t> call(const Mat &img, const float *bbox, const float *landmarks_5x2) {
        std::vector<uint8_t> crop((size_t)cfg_.crop_h * cfg_.crop_w * 3);
        check(fd_align(c_.get(), img.data, img.rows, img.cols, img.step, bbox, landmarks_5x2, crop.data(), nullptr, nullptr));
        return crop;
    }
private:
    Context &c_;
    fd_config cfg_{};
};

// utils::utils::byte_data_to_opencv (src/utils/utils.rs:8-52) for baseline JPEG: encoded bytes -> BGR pixels (rows * cols * 3),
// bit-identical to cv::imdecode.  Throws for streams the decoder does not cover (the reference hands those to OpenCV).
inline std::vector<uint8_t> byte_data_to_opencv(Context &c, const uint8_t *im_bytes, size_t n, int *rows, int *cols) {
    int h = 0, w = 0, ss = 0;
    check(fd_jpeg_info(im_bytes, n, &h, &w, &ss));
    std::vector<uint8_t> img((size_t)h * w * 3);
    check(fd_imdecode(c.get(), im_bytes, n, img.data(), w * 3));
    *rows = h;
    *cols = w;
    return img;
}

// pipeline::module::face_selection::FaceSelection (src/pipeline/module/face_selection.rs:5-189)
class FaceSelection {
public:
    FaceSelection(Context &c, float margin_center_left_ratio = 0.3f, float margin_center_right_ratio = 0.3f, float margin_edge_ratio = 0.1f,
                  float minimum_face_ratio = 0.0075f)
        : c_(c), p_{margin_center_left_ratio, margin_center_right_ratio, margin_edge_ratio, minimum_face_ratio} {}
    // call(&Mat, face_boxes (M,5), key_points (M,5,2) or nullptr, is_enroll) -> {row of the box, row of the key points}; -1 = None
    std::pair<int, int> call(const Mat &img, const float *face_boxes, const float *key_points, int M, bool is_enroll = false) {
        int bi = -1, ki = -1;
        check(fd_face_selection(c_.get(), img.rows, img.cols, face_boxes, key_points, M, is_enroll ? 1 : 0, &p_, &bi, &ki));
        return {bi, ki};
    }
private:
    Context &c_;
    fd_select_params p_;
};

}  // namespace fd
