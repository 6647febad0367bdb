//! Raw bindings to include/rtrace.h (the subset the host needs).
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct RtScene {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct RtCamera {
    pub eye: [f32; 3],
    pub right: [f32; 3],
    pub up: [f32; 3],
    pub forward: [f32; 3],
}

#[repr(C)]
#[derive(Default, Clone, Copy)]
pub struct RtStats {
    pub primary_rays: u64,
    pub shadow_rays: u64,
    pub kernel_ms: f64,
    pub total_ms: f64,
    pub kernel_launches: u32,
    pub gpus: u32,
}

pub type RtFrameCallback = extern "C" fn(user: *mut c_void, frame: u32, rgba: *const u8, len: usize);

extern "C" {
    pub fn rt_last_error() -> *const c_char;
    pub fn rt_set_device(device: c_int) -> c_int;
    pub fn rt_scene_create(level: u32, origin: *const f32, radius: f32, light: *const f32, eye: *const f32,
                           out: *mut *mut RtScene) -> c_int;
    pub fn rt_scene_destroy(s: *mut RtScene);
    pub fn rt_scene_counts(s: *const RtScene, groups: *mut u64, items: *mut u64) -> c_int;
    pub fn rt_render_region(s: *const RtScene, width: u16, height: u16, spp: u16, l: u16, b: u16, r: u16, t: u16,
                            rgba_out: *mut u8, rgba_len: usize) -> c_int;
    pub fn rt_render_preview(s: *const RtScene, camera: *const RtCamera, width: u32, height: u32, step: u32,
                             rgba_out: *mut u8, rgba_len: usize, stream: *mut c_void) -> c_int;
    pub fn rt_render_frame_multi(scenes: *const *mut RtScene, ngpu: c_int, camera: *const RtCamera, width: u32,
                                 height: u32, spp: u32, rgba_out: *mut u8, rgba_len: usize, stats: *mut RtStats) -> c_int;
    pub fn rt_render_sweep(s: *const RtScene, cameras: *const RtCamera, n_frames: u32, width: u32, height: u32,
                           spp: u32, cb: RtFrameCallback, user: *mut c_void, stats: *mut RtStats) -> c_int;
}

pub fn last_error() -> String {
    unsafe { std::ffi::CStr::from_ptr(rt_last_error()) }.to_string_lossy().into_owned()
}
