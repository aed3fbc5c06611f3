//! src/ctx.rs — the per-thread, per-device `fd_ctx` behind the reference-signature wrappers.
//! One fd_ctx per GPU per host thread (include/fd_b200.h: a ctx is not internally locked).  The handle is owned by a
//! type that implements Drop, so a thread that exits releases its stream and workspaces.
use std::cell::RefCell;
use std::collections::HashMap;
use crate::ffi;

pub struct Ctx(*mut ffi::fd_ctx);

impl Ctx {
    pub fn new(device_id: i32, cfg: Option<&ffi::fd_config>) -> anyhow::Result<Ctx> {
        let mut c = std::ptr::null_mut();
        let p = cfg.map_or(std::ptr::null(), |c| c as *const ffi::fd_config);
        ffi::check(unsafe { ffi::fd_ctx_create(device_id, p, &mut c) })?;
        Ok(Ctx(c))
    }
    pub fn raw(&self) -> *mut ffi::fd_ctx { self.0 }
}

impl Drop for Ctx {
    fn drop(&mut self) {
        unsafe { ffi::fd_ctx_destroy(self.0) }
    }
}

thread_local! {
    static CTXS: RefCell<HashMap<i32, Ctx>> = RefCell::new(HashMap::new());
    static DEVICE: RefCell<i32> = RefCell::new(0);
}

/// Device the free functions (`processing::*`, `rcnn::*`) of this thread run on (default 0).
pub fn set_device(device_id: i32) { DEVICE.with(|d| *d.borrow_mut() = device_id); }

/// Runs `f` with this thread's default-config context on its current device, creating it on first use.
pub fn with_ctx<R>(f: impl FnOnce(*mut ffi::fd_ctx) -> R) -> R {
    let dev = DEVICE.with(|d| *d.borrow());
    CTXS.with(|m| {
        let mut m = m.borrow_mut();
        if !m.contains_key(&dev) {
            // there is no CPU fallback: without a GPU this is where the crate fails, loudly
            m.insert(dev, Ctx::new(dev, None).expect("fd_ctx_create (libfd_b200 needs a CUDA device)"));
        }
        f(m[&dev].raw())
    })
}
