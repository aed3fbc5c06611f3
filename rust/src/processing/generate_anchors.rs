//! Drop-in bodies for src/processing/generate_anchors.rs:41-138 (init-time tables: host arithmetic inside libfd_b200, no
//! GPU involved).  `Config` / `AnchorConfig` (generate_anchors.rs:7-18) and the private helpers stay as in the reference.
use ndarray::{Array1, Array2};
use std::collections::HashMap;
use crate::ffi;

#[derive(Debug, Clone)]
pub struct Config {
    pub rpn_anchor_cfg: HashMap<String, AnchorConfig>,
}

#[derive(Debug, Clone)]
pub struct AnchorConfig {
    pub base_size: i32,
    pub ratios: Vec<f32>,
    pub scales: Vec<f32>,
    pub allowed_border: i32,
}

/// generate_anchors.rs:41-59
pub fn generate_anchors(base_size: usize, ratios: Array1<f32>, scales: Array1<f32>) -> Array2<f32> {
    generate_anchors2(base_size, ratios, scales, 0, false)
}

/// generate_anchors.rs:61-93
pub fn generate_anchors2(base_size: usize, ratios: Array1<f32>, scales: Array1<f32>, stride: usize, dense_anchor: bool) -> Array2<f32> {
    let (r, s) = (ratios.to_vec(), scales.to_vec());
    let cap = r.len() * s.len() * if dense_anchor { 2 } else { 1 };
    let mut out = vec![0f32; (cap * 4).max(4)];
    let mut n = 0;
    ffi::check(unsafe {
        ffi::fd_generate_anchors2(base_size as i32, r.as_ptr(), r.len() as i32, s.as_ptr(), s.len() as i32, stride as i32,
                                  dense_anchor as i32, out.as_mut_ptr(), &mut n)
    })
    .expect("fd_generate_anchors2");
    out.truncate(n as usize * 4);
    Array2::from_shape_vec((n as usize, 4), out).unwrap()
}

/// generate_anchors.rs:95-114 — level i uses ratios[i], scales[i].
pub fn generate_anchors_fpn(base_size: Vec<i32>, ratios: Vec<f32>, scales: Vec<f32>) -> Vec<Array2<f32>> {
    let n = base_size.len();
    let mut out = vec![0f32; (n * 4).max(4)];
    ffi::check(unsafe { ffi::fd_generate_anchors_fpn(base_size.as_ptr(), ratios.as_ptr(), scales.as_ptr(), n as i32, out.as_mut_ptr()) })
        .expect("fd_generate_anchors_fpn");
    (0..n).map(|i| Array2::from_shape_vec((1, 4), out[4 * i..4 * i + 4].to_vec()).unwrap()).collect()
}

/// generate_anchors.rs:116-138 — strides processed in descending order; `cfg = None` is `unimplemented!` in the reference.
pub fn generate_anchors_fpn2(dense_anchor: bool, cfg: Option<&Config>) -> Vec<Array2<f32>> {
    let config = cfg.unwrap_or_else(|| unimplemented!("Config loading not implemented"));
    let mut cfgs: Vec<ffi::fd_anchor_cfg> = Vec::new();
    for (k, v) in config.rpn_anchor_cfg.iter() {
        let mut c = ffi::fd_anchor_cfg {
            stride: k.parse::<i32>().unwrap(), base_size: v.base_size, n_ratios: v.ratios.len() as i32, n_scales: v.scales.len() as i32,
            ratios: [0.0; 8], scales: [0.0; 8], allowed_border: v.allowed_border,
        };
        c.ratios[..v.ratios.len()].copy_from_slice(&v.ratios);
        c.scales[..v.scales.len()].copy_from_slice(&v.scales);
        cfgs.push(c);
    }
    let per = if dense_anchor { 2 } else { 1 };
    let cap: usize = cfgs.iter().map(|c| (c.n_ratios * c.n_scales) as usize * per).sum();
    let mut out = vec![0f32; (cap * 4).max(4)];
    let mut rows = vec![0i32; cfgs.len()];
    let mut strides = vec![0i32; cfgs.len()];
    ffi::check(unsafe {
        ffi::fd_generate_anchors_fpn2(dense_anchor as i32, cfgs.as_ptr(), cfgs.len() as i32, out.as_mut_ptr(), rows.as_mut_ptr(), strides.as_mut_ptr())
    })
    .expect("fd_generate_anchors_fpn2");
    let mut res = Vec::new();
    let mut row = 0usize;
    for &r in rows.iter() {
        let r = r as usize;
        res.push(Array2::from_shape_vec((r, 4), out[4 * row..4 * (row + r)].to_vec()).unwrap());
        row += r;
    }
    res
}
