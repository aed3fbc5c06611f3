"""Mirror of the reference's `pipeline::module` detector / aligner (src/pipeline/module/)."""
from .face_detection import RetinaFaceDetection  # noqa: F401
from .face_alignment import FaceAlignment  # noqa: F401
from .face_selection import FaceSelection  # noqa: F401
