//! Facade with ochre's public names (`ochre::{Vec2, Mat2x2, Transform, PathCmd, Rasterizer,
//! TileBuilder, TILE_SIZE}`, reference src/lib.rs:47-53) over the C ABI of `include/ochre_b200.h`.
//!
//! NOTE: this crate is written but was never compiled: the build image has no Rust toolchain.
//! All logic lives behind the C ABI; this file only marshals.
use std::os::raw::{c_char, c_int};

pub const TILE_SIZE: usize = 8;

#[repr(C)]
#[derive(Copy, Clone, Debug, PartialEq)]
pub struct Vec2 { pub x: f32, pub y: f32 }
impl Vec2 { pub fn new(x: f32, y: f32) -> Vec2 { Vec2 { x, y } } }

#[repr(C)]
#[derive(Copy, Clone, Debug, PartialEq)]
pub struct Mat2x2(pub [f32; 4]);
impl Mat2x2 {
    pub fn new(a: f32, b: f32, c: f32, d: f32) -> Mat2x2 { Mat2x2([a, b, c, d]) }
    pub fn id() -> Mat2x2 { Mat2x2([1.0, 0.0, 0.0, 1.0]) }
    pub fn scale(s: f32) -> Mat2x2 { Mat2x2([s, 0.0, 0.0, s]) }
    pub fn rotate(a: f32) -> Mat2x2 { Mat2x2([a.cos(), a.sin(), -a.sin(), a.cos()]) }
    fn mul(self, r: Mat2x2) -> Mat2x2 {
        Mat2x2([self.0[0] * r.0[0] + self.0[1] * r.0[2], self.0[0] * r.0[1] + self.0[1] * r.0[3],
                self.0[2] * r.0[0] + self.0[3] * r.0[2], self.0[2] * r.0[1] + self.0[3] * r.0[3]])
    }
    fn apply(self, v: Vec2) -> Vec2 { Vec2::new(self.0[0] * v.x + self.0[1] * v.y, self.0[2] * v.x + self.0[3] * v.y) }
}

#[repr(C)]
#[derive(Copy, Clone, Debug, PartialEq)]
pub struct Transform { pub matrix: Mat2x2, pub offset: Vec2 }
impl Transform {
    pub fn new(matrix: Mat2x2, offset: Vec2) -> Transform { Transform { matrix, offset } }
    pub fn id() -> Transform { Transform::new(Mat2x2::id(), Vec2::new(0.0, 0.0)) }
    pub fn translate(x: f32, y: f32) -> Transform { Transform::new(Mat2x2::id(), Vec2::new(x, y)) }
    pub fn scale(s: f32) -> Transform { Transform::new(Mat2x2::scale(s), Vec2::new(0.0, 0.0)) }
    pub fn rotate(a: f32) -> Transform { Transform::new(Mat2x2::rotate(a), Vec2::new(0.0, 0.0)) }
    pub fn then(self, t: Transform) -> Transform {
        let mo = t.matrix.apply(self.offset);
        Transform::new(t.matrix.mul(self.matrix), Vec2::new(mo.x + t.offset.x, mo.y + t.offset.y))
    }
    pub fn apply(self, v: Vec2) -> Vec2 { let r = self.matrix.apply(v); Vec2::new(r.x + self.offset.x, r.y + self.offset.y) }
    fn to_c(&self) -> OchreTransform { OchreTransform { m: self.matrix.0, ox: self.offset.x, oy: self.offset.y } }
}

#[derive(Copy, Clone)]
pub enum PathCmd { Move(Vec2), Line(Vec2), Quadratic(Vec2, Vec2), Cubic(Vec2, Vec2, Vec2), Conic(Vec2, Vec2, f32), Close }

#[repr(C)]
#[derive(Copy, Clone)]
struct OchreCmd { tag: u32, v: [f32; 6] }
#[repr(C)]
#[derive(Copy, Clone)]
struct OchreTransform { m: [f32; 4], ox: f32, oy: f32 }
#[repr(C)]
#[derive(Copy, Clone)]
struct OchreSpan { x: i16, y: i16, w: u16, pad: u16 }
#[repr(C)]
#[derive(Copy, Clone)]
struct OchrePathRange { tile_start: u32, n_tiles: u32, span_start: u32, n_spans: u32 }
#[repr(C)]
struct OchreResult {
    n_paths: u32, n_tiles: u32, n_spans: u32, reserved: u32,
    tile_off: *const u32, tile_xy: *const i16, alpha: *const u8, span_off: *const u32, spans: *const OchreSpan,
    n_cmds: u64, n_lines: u64, n_records: u64, n_chunks: u64, kernel_launches: u64, device_ms: f32, stage_ms: [f32; 8],
    ranges: *const OchrePathRange,
}
#[repr(C)]
struct Ctx { _p: [u8; 0] }

extern "C" {
    fn ochre_b200_create(device: c_int, out: *mut *mut Ctx) -> c_int;
    fn ochre_b200_destroy(ctx: *mut Ctx) -> c_int;
    fn ochre_b200_rasterize(ctx: *mut Ctx, cmds: *const OchreCmd, cmd_off: *const u32, xf: *const OchreTransform,
                            n_paths: u32, flags: u32, cmd_off_host: *const u32, out: *mut OchreResult) -> c_int;
    fn ochre_b200_rasterize_paints(ctx: *mut Ctx, cmds: *const OchreCmd, cmd_off: *const u32, xf: *const OchreTransform,
                                   stroke_width: *const f32, n_paths: u32, flags: u32, cmd_off_host: *const u32,
                                   out: *mut OchreResult) -> c_int;
    fn ochre_b200_last_error(ctx: *const Ctx) -> *const c_char;
    fn ochre_b200_stroke_path(path: *const OchreCmd, n: usize, width: f32, out: *mut *mut OchreCmd, n_out: *mut usize) -> c_int;
    fn ochre_b200_free(p: *mut std::ffi::c_void);
}

impl PathCmd {
    pub fn transform(&self, t: Transform) -> PathCmd {
        match *self {
            PathCmd::Move(p) => PathCmd::Move(t.apply(p)),
            PathCmd::Line(p) => PathCmd::Line(t.apply(p)),
            PathCmd::Quadratic(c, p) => PathCmd::Quadratic(t.apply(c), t.apply(p)),
            PathCmd::Cubic(a, b, p) => PathCmd::Cubic(t.apply(a), t.apply(b), t.apply(p)),
            PathCmd::Conic(c, p, w) => PathCmd::Conic(t.apply(c), t.apply(p), w),
            PathCmd::Close => PathCmd::Close,
        }
    }
    fn to_c(&self) -> OchreCmd {
        match *self {
            PathCmd::Move(p) => OchreCmd { tag: 0, v: [p.x, p.y, 0.0, 0.0, 0.0, 0.0] },
            PathCmd::Line(p) => OchreCmd { tag: 1, v: [p.x, p.y, 0.0, 0.0, 0.0, 0.0] },
            PathCmd::Quadratic(c, p) => OchreCmd { tag: 2, v: [c.x, c.y, p.x, p.y, 0.0, 0.0] },
            PathCmd::Cubic(a, b, p) => OchreCmd { tag: 3, v: [a.x, a.y, b.x, b.y, p.x, p.y] },
            PathCmd::Conic(c, p, w) => OchreCmd { tag: 4, v: [c.x, c.y, p.x, p.y, w, 0.0] },
            PathCmd::Close => OchreCmd { tag: 5, v: [0.0; 6] },
        }
    }
}

pub trait TileBuilder {
    fn tile(&mut self, x: i16, y: i16, data: [u8; TILE_SIZE * TILE_SIZE]);
    fn span(&mut self, x: i16, y: i16, width: u16);
}

/// One CUDA device (`ochre_b200_ctx`).  Not `Sync`; one per thread.
pub struct Device { ctx: *mut Ctx }
impl Device {
    pub fn new(device: i32) -> Result<Device, i32> {
        let mut ctx = std::ptr::null_mut();
        let rc = unsafe { ochre_b200_create(device, &mut ctx) };
        if rc != 0 { Err(rc) } else { Ok(Device { ctx }) }
    }
}
impl Drop for Device { fn drop(&mut self) { unsafe { ochre_b200_destroy(self.ctx); } } }

thread_local! { static DEFAULT: Device = Device::new(0).expect("ochre-b200: no CUDA device (there is no CPU fallback)"); }

/// Same public surface as `ochre::Rasterizer` (reference src/rasterizer.rs:40-180).
pub struct Rasterizer { cmds: Vec<OchreCmd> }
impl Rasterizer {
    pub fn new() -> Rasterizer { Rasterizer { cmds: Vec::new() } }
    pub fn move_to(&mut self, p: Vec2) { self.cmds.push(PathCmd::Move(p).to_c()); }
    pub fn line_to(&mut self, p: Vec2) { self.cmds.push(PathCmd::Line(p).to_c()); }
    pub fn command(&mut self, c: PathCmd) { self.cmds.push(c.to_c()); }
    pub fn fill(&mut self, path: &[PathCmd], transform: Transform) {
        for c in path { self.cmds.push(c.transform(transform).to_c()); }
    }
    pub fn stroke(&mut self, path: &[PathCmd], width: f32, transform: Transform) {
        let src: Vec<OchreCmd> = path.iter().map(|c| c.to_c()).collect();
        let (mut out, mut n) = (std::ptr::null_mut(), 0usize);
        let rc = unsafe { ochre_b200_stroke_path(src.as_ptr(), src.len(), width, &mut out, &mut n) };
        if rc != 0 { panic!("stroke: {}", rc); }
        let poly = unsafe { std::slice::from_raw_parts(out, n) };
        for c in poly {
            let p = Vec2::new(c.v[0], c.v[1]);
            self.cmds.push(match c.tag { 0 => PathCmd::Move(transform.apply(p)), 1 => PathCmd::Line(transform.apply(p)), _ => PathCmd::Close }.to_c());
        }
        unsafe { ochre_b200_free(out as *mut _) };
    }
    pub fn finish<B: TileBuilder>(self, builder: &mut B) {
        finish_batch(vec![self], std::slice::from_mut(builder));
    }
}

/// One paint of a document, as `examples/svg.rs:140-154` hands them to a fresh `Rasterizer` each:
/// `fill(path, transform)` when `stroke_width` is `None`, else `stroke(path, width, transform)`.
pub struct Paint<'a> { pub path: &'a [PathCmd], pub transform: Transform, pub stroke_width: Option<f32> }

/// All paints of a document in one GPU submission; strokes are flattened and offset on the device
/// (`ochre_b200_rasterize_paints`).  `builders[i]` receives paint i's calls in the reference's order.
pub fn finish_paints<B: TileBuilder>(paints: &[Paint], builders: &mut [B]) {
    let mut cmds = Vec::new();
    let mut off = vec![0u32];
    let mut xf = Vec::with_capacity(paints.len());
    let mut width = Vec::with_capacity(paints.len());
    for p in paints {
        cmds.extend(p.path.iter().map(|c| c.to_c()));
        off.push(cmds.len() as u32);
        xf.push(p.transform.to_c());
        width.push(p.stroke_width.unwrap_or(0.0));
    }
    submit(&cmds, &off, &xf, Some(&width), builders);
}

/// `finish` for many rasterisers in one GPU submission (the throughput entry point).
pub fn finish_batch<B: TileBuilder>(rasterizers: Vec<Rasterizer>, builders: &mut [B]) {
    let mut cmds = Vec::new();
    let mut off = vec![0u32];
    for r in &rasterizers { cmds.extend_from_slice(&r.cmds); off.push(cmds.len() as u32); }
    let xf = vec![OchreTransform { m: [1.0, 0.0, 0.0, 1.0], ox: 0.0, oy: 0.0 }; rasterizers.len()];
    submit(&cmds, &off, &xf, None, builders);
}

fn submit<B: TileBuilder>(cmds: &[OchreCmd], off: &[u32], xf: &[OchreTransform], width: Option<&[f32]>, builders: &mut [B]) {
    let n_paths = off.len() - 1;
    DEFAULT.with(|dev| unsafe {
        let mut res: OchreResult = std::mem::zeroed();
        let rc = match width {
            Some(w) => ochre_b200_rasterize_paints(dev.ctx, cmds.as_ptr(), off.as_ptr(), xf.as_ptr(), w.as_ptr(), n_paths as u32, 0, std::ptr::null(), &mut res),
            None => ochre_b200_rasterize(dev.ctx, cmds.as_ptr(), off.as_ptr(), xf.as_ptr(), n_paths as u32, 0, std::ptr::null(), &mut res),
        };
        if rc != 0 {
            panic!("ochre_b200_rasterize: {} ({})", std::ffi::CStr::from_ptr(ochre_b200_last_error(dev.ctx)).to_string_lossy(), rc);
        }
        let tile_off = std::slice::from_raw_parts(res.tile_off, n_paths + 1);
        let span_off = std::slice::from_raw_parts(res.span_off, n_paths + 1);
        let xy = std::slice::from_raw_parts(res.tile_xy, 2 * res.n_tiles as usize);
        let alpha = std::slice::from_raw_parts(res.alpha, 64 * res.n_tiles as usize);
        let spans = std::slice::from_raw_parts(res.spans, res.n_spans as usize);
        for (p, b) in builders.iter_mut().enumerate() {
            let (mut s, s1) = (span_off[p] as usize, span_off[p + 1] as usize);
            for t in tile_off[p] as usize..tile_off[p + 1] as usize {
                let (x, y) = (xy[2 * t], xy[2 * t + 1]);
                let mut data = [0u8; 64];
                data.copy_from_slice(&alpha[64 * t..64 * t + 64]);
                b.tile(x, y, data);
                if s < s1 && spans[s].y == y && spans[s].x == x + TILE_SIZE as i16 {
                    b.span(spans[s].x, spans[s].y, spans[s].w);
                    s += 1;
                }
            }
        }
    });
}
