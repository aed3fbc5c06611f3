"""pipeline::module::face_detection::RetinaFaceDetection (src/pipeline/module/face_detection.rs).

The CNN forward pass stays behind the serving boundary (Triton gRPC in the reference, face_detection.rs:279): the
caller supplies `infer`, a callable tensor -> 9 head tensors.  Everything around it runs on the GPU through the C ABI.
"""
from .. import Context


class RetinaFaceDetection:
    def __init__(self, infer, image_size=(640, 640), confidence_threshold=0.7, iou_threshold=0.45, ctx=None, device=0):
        from ..ffi import default_config
        cfg = default_config()
        cfg.image_w, cfg.image_h = image_size
        cfg.conf_thr, cfg.iou_thr = confidence_threshold, iou_threshold
        self.ctx = ctx or Context(device, cfg)
        self.infer = infer
        self.confidence_threshold, self.iou_threshold = confidence_threshold, iou_threshold

    def _preprocess(self, img):                       # face_detection.rs:131-230 (letterbox + tensor)
        return self.ctx.preprocess(img)

    def _forward(self, tensor, det_scale):            # :232-471 with :473-493 folded in (rescale on the device)
        heads = self.infer(tensor)
        return self.ctx.detect(heads, det_scale, self.confidence_threshold, self.iou_threshold)

    def call(self, image):                            # :496-513 -> (det (M,5), landmarks (M,5,2))
        tensor, det_scale = self._preprocess(image)
        return self._forward(tensor, det_scale)
