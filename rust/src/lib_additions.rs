//! The `mod` lines a maintainer adds to the crate root (src/main.rs / lib.rs of the reference) for the files in this tree.
//! The paths mirror the reference's module tree, so each file REPLACES the body of the file of the same name.
pub mod ffi;                 // generated from include/fd_b200.h (scripts/gen_rust_ffi.py)
pub mod ctx;
pub mod processing {
    pub mod nms;             // src/processing/nms.rs
    pub mod bbox_transform;  // src/processing/bbox_transform.rs
    pub mod generate_anchors;
}
pub mod rcnn {
    pub mod anchors;         // src/rcnn/anchors.rs
    pub mod bbox;
    pub mod cpu_nms;
    pub mod gpu_nms;         // the binding the reference left commented out
}
pub mod utils {
    pub mod utils;           // byte_data_to_opencv (src/utils/utils.rs:8-52); the reference's body stays as utils_opencv
}
pub mod pipeline {
    pub mod module {
        pub mod face_detection;
        pub mod face_alignment;
        pub mod face_selection;
    }
}
