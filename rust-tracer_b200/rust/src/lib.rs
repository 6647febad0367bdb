//! `sphere_tracer` with the reference crate's public names (Scene, Renderer, RenderOptions, the
//! RGBABufferWriter seam and the PPM sink), rendering through librtrace_b200 on B200 GPUs.
//! UNBUILT in this repository (no Rust toolchain); the tested twin is ../host/render.hpp.
pub mod ffi;

use std::io::{self, Seek, Write};
use std::sync::Arc;
use std::time::{Duration, Instant};

/// Stand-in for `threadpool::ThreadPool` in `Renderer::render`'s signature (render.rs:260-263): the
/// reference sizes its CPU pool with it; here the work runs on the GPUs of the scene and the value is
/// accepted and ignored, so that callers written against the reference keep compiling.
pub struct ThreadPool {
    pub threads: usize,
}

impl ThreadPool {
    pub fn new(threads: usize) -> ThreadPool {
        assert!(threads >= 1);
        ThreadPool { threads }
    }
}

#[derive(Clone, Copy)]
pub struct RenderOptions {
    pub width: u16,
    pub height: u16,
    pub samples_per_pixel: u16,
}

/// `[l, r) x [b, t)`, `b` being the upper image row.
#[derive(Clone, Copy, PartialEq)]
pub struct ImageRegion {
    pub l: u16,
    pub t: u16,
    pub r: u16,
    pub b: u16,
}

impl ImageRegion {
    pub fn width(&self) -> u16 { self.r - self.l }
    pub fn height(&self) -> u16 { self.t - self.b }
    pub fn area(&self) -> usize { self.width() as usize * self.height() as usize }
    pub fn contains(&self, o: &ImageRegion) -> bool { o.l >= self.l && o.b >= self.b && o.t <= self.t && o.r <= self.r }
}

pub struct RGBABuffer {
    pub buf: Vec<u8>,
    pub reg: ImageRegion,
}

impl RGBABuffer {
    pub fn new(reg: &ImageRegion) -> RGBABuffer { RGBABuffer { buf: vec![0u8; reg.area() * 4], reg: *reg } }
}

/// The output seam: total resolution first, then finished regions anywhere inside it.
pub trait RGBABufferWriter {
    fn begin(&mut self, x: u16, y: u16);
    fn write_rgba_buffer(&mut self, buffer: &RGBABuffer);
}

/// One replica of the flattened scene per GPU.
pub struct Scene {
    replicas: Vec<*mut ffi::RtScene>,
}

impl Scene {
    /// `Scene::default()` generalised by pyramid level and GPU count.
    pub fn with_level(level: u32, gpus: usize) -> Scene {
        let (origin, light, eye) = ([0.0f32, -1.0, 0.0], [-1.0f32, -3.0, 2.0], [0.0f32, 0.0, -4.0]);
        let mut replicas = Vec::new();
        for g in 0..gpus.max(1) {
            let mut s = std::ptr::null_mut();
            unsafe {
                assert!(ffi::rt_set_device(g as i32) == 0, "{}", ffi::last_error());
                let rc = ffi::rt_scene_create(level, origin.as_ptr(), 1.0, light.as_ptr(), eye.as_ptr(), &mut s);
                assert!(rc == 0, "Scene::default: {}", ffi::last_error());
            }
            replicas.push(s);
        }
        unsafe { ffi::rt_set_device(0) };
        Scene { replicas }
    }
}

impl Default for Scene {
    fn default() -> Scene { Scene::with_level(8, 1) }
}

impl Drop for Scene {
    fn drop(&mut self) {
        for s in self.replicas.drain(..) {
            unsafe { ffi::rt_scene_destroy(s) };
        }
    }
}

pub struct Renderer;

impl Renderer {
    /// One region of the frame on GPU 0 (same semantics as the reference's `render_region`).
    pub fn render_region(o: &RenderOptions, scene: &Scene, buf: &mut RGBABuffer) {
        let r = buf.reg;
        let rc = unsafe {
            ffi::rt_render_region(scene.replicas[0], o.width, o.height, o.samples_per_pixel, r.l, r.b, r.r, r.t,
                                  buf.buf.as_mut_ptr(), buf.buf.len())
        };
        assert!(rc == 0, "render_region: {}", ffi::last_error());
    }

    /// Undersampled preview (README "season 2"): one traced pixel per `step` x `step` block, 1 sample per pixel.
    pub fn render_preview(o: &RenderOptions, scene: &Scene, writer: &mut dyn RGBABufferWriter, step: u32) {
        writer.begin(o.width, o.height);
        let mut frame = RGBABuffer::new(&ImageRegion { l: 0, r: o.width, b: 0, t: o.height });
        let rc = unsafe {
            ffi::rt_render_preview(scene.replicas[0], std::ptr::null(), o.width as u32, o.height as u32, step,
                                   frame.buf.as_mut_ptr(), frame.buf.len(), std::ptr::null_mut())
        };
        assert!(rc == 0, "render_preview: {}", ffi::last_error());
        writer.write_rgba_buffer(&frame);
    }

    /// The reference's signature (render.rs:260-263).  The whole frame on every GPU of the scene: interleaved
    /// blocks of 16 rows, each GPU copying its blocks straight into the host frame.  Unlike the reference,
    /// sizes need not be multiples of 64.
    pub fn render(o: &RenderOptions, scene: Arc<Scene>, writer: &mut dyn RGBABufferWriter, _pool: &ThreadPool) {
        writer.begin(o.width, o.height);
        let mut frame = RGBABuffer::new(&ImageRegion { l: 0, r: o.width, b: 0, t: o.height });
        let rc = unsafe {
            ffi::rt_render_frame_multi(scene.replicas.as_ptr(), scene.replicas.len() as i32, std::ptr::null(),
                                       o.width as u32, o.height as u32, o.samples_per_pixel as u32,
                                       frame.buf.as_mut_ptr(), frame.buf.len(), std::ptr::null_mut())
        };
        assert!(rc == 0, "render: {}", ffi::last_error());
        writer.write_rgba_buffer(&frame);
    }

    /// A sweep of frames (extension): frames are drawn from one queue by the scene's GPUs
    /// (rt_render_sweep_multi) and `sink(frame, bytes)` is called in frame order on this thread;
    /// `rgb` delivers RGB8 (the body of the P6 file).
    pub fn render_sweep<F: FnMut(u32, &[u8])>(o: &RenderOptions, scene: &Scene, cameras: &[ffi::RtCamera], rgb: bool,
                                              mut sink: F) {
        extern "C" fn trampoline<F: FnMut(u32, &[u8])>(user: *mut std::os::raw::c_void, frame: u32, data: *const u8, len: usize) {
            let sink = unsafe { &mut *(user as *mut F) };
            sink(frame, unsafe { std::slice::from_raw_parts(data, len) });
        }
        let rc = unsafe {
            ffi::rt_render_sweep_multi(scene.replicas.as_ptr(), scene.replicas.len() as i32, cameras.as_ptr(),
                                       cameras.len() as u32, o.width as u32, o.height as u32,
                                       o.samples_per_pixel as u32, rgb as i32, trampoline::<F>,
                                       &mut sink as *mut F as *mut std::os::raw::c_void, std::ptr::null_mut())
        };
        assert!(rc == 0, "render_sweep: {}", ffi::last_error());
    }
}

pub enum FileOrAnyWriter {
    AnyWriter(io::Stdout),
    FileWriter(io::BufWriter<std::fs::File>),
}

/// Binary PPM (P6, or P5 grey) of the RGBA frame with alpha dropped; a file sink rewrites the whole image
/// when a buffer arrives and none was written yet or the last write is a second old (render.rs:426-432);
/// flushed once more on drop (render.rs:331-335).
pub struct PPMStdoutRGBABufferWriter<'a> {
    out: &'a mut FileOrAnyWriter,
    dims: Option<(u16, u16)>,
    image: Option<RGBABuffer>,
    rgb: bool,
    dirty: bool,
    last_written_at: Option<Instant>,
}

impl<'a> PPMStdoutRGBABufferWriter<'a> {
    pub fn new(write_rgb: bool, out: &'a mut FileOrAnyWriter) -> Self {
        PPMStdoutRGBABufferWriter { out, dims: None, image: None, rgb: write_rgb, dirty: false, last_written_at: None }
    }

    fn flush_image(&mut self) {
        if !self.dirty { return; }
        let (w, h) = self.dims.expect("begin() called");
        let image = self.image.as_ref().unwrap();
        let mut body = Vec::with_capacity(image.buf.len() / 4 * 3);
        for px in image.buf.chunks(4) {
            if self.rgb { body.extend_from_slice(&px[..3]); }
            else { body.push(((px[0] as f32 + px[1] as f32 + px[2] as f32) / 3.0) as u8); }
        }
        let header = format!("{}\n{} {}\n255\n", if self.rgb { "P6" } else { "P5" }, w, h);
        match *self.out {
            FileOrAnyWriter::FileWriter(ref mut f) => {
                f.get_mut().set_len(0).unwrap();
                f.seek(io::SeekFrom::Start(0)).unwrap();
                f.write_all(header.as_bytes()).unwrap();
                f.write_all(&body).unwrap();
                f.flush().ok();
            }
            FileOrAnyWriter::AnyWriter(ref mut o) => {
                o.write_all(header.as_bytes()).unwrap();
                o.write_all(&body).unwrap();
                o.flush().ok();
            }
        }
        self.dirty = false;
    }
}

impl<'a> RGBABufferWriter for PPMStdoutRGBABufferWriter<'a> {
    fn begin(&mut self, x: u16, y: u16) {
        self.dims = Some((x, y));
        self.image = Some(RGBABuffer::new(&ImageRegion { l: 0, r: x, b: 0, t: y }));
    }

    fn write_rgba_buffer(&mut self, buffer: &RGBABuffer) {
        let image = self.image.as_mut().unwrap();
        assert!(image.reg.contains(&buffer.reg));
        let w = buffer.reg.width() as usize * 4;
        for y in buffer.reg.b..buffer.reg.t {
            let dst = ((y - image.reg.b) as usize * image.reg.width() as usize + (buffer.reg.l - image.reg.l) as usize) * 4;
            let src = (y - buffer.reg.b) as usize * w;
            image.buf[dst..dst + w].copy_from_slice(&buffer.buf[src..src + w]);
        }
        self.dirty = true;
        // "Flush full image right away" (render.rs:426-432): a file sink, at most once per second
        if let FileOrAnyWriter::FileWriter(_) = *self.out {
            if self.last_written_at.map_or(true, |t| t + Duration::from_secs(1) <= Instant::now()) {
                self.last_written_at = Some(Instant::now());
                self.flush_image();
            }
        }
    }
}

impl<'a> Drop for PPMStdoutRGBABufferWriter<'a> {
    fn drop(&mut self) { self.flush_image(); }
}
