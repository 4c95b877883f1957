//! `B200Renderer`: the reference crate's traits over `libswfr_b200.so` (goes to `rs/src/b200/mod.rs`).
//!
//! | reference interface | file:line | here |
//! |---|---|---|
//! | `HeadlessGfxRenderer::new(&instance, w, h) -> Result<_, &'static str>` | rs/src/headless_renderer.rs:60-64 | `B200Renderer::new(w, h)` |
//! | `define_shape(&DefineShape) -> usize` | rs/src/headless_renderer.rs:229-231 | `define_shape` |
//! | `ClientAssetStore::{register_shape, register_morph_shape}` | rs/src/asset.rs:9-12 | `impl ClientAssetStore` |
//! | `SwfRenderer::render(&mut self, stage: Stage)` | rs/src/swf_renderer.rs:3-5 | `impl SwfRenderer` |
//! | `Renderer::set_stage(DisplayItem::Shape(id, Matrix))` | rs/src/renderer.rs:81-87 | `impl Renderer` |
//! | `get_image() -> Result<Image, &'static str>` | rs/src/headless_renderer.rs:233-244, 725-868 | `get_image` |
//! | `Drop` (wait idle, free) | rs/src/headless_renderer.rs:871-904 | `impl Drop` |
//!
//! A handle owns one CUDA device and is `!Sync` like the reference's `&mut self` API; different handles may be
//! driven from different threads (one per GPU: frames shard over handles without any collective).

pub mod convert;
pub mod ffi;

use self::convert::{to_c_define_morph_shape, to_c_define_shape};
use self::ffi::*;
use crate::asset::{ClientAssetStore, MorphShapeId, ShapeId};
use crate::renderer::{DisplayItem, Image, ImageMetadata, Renderer};
use crate::stage::{DisplayPrimitive, Matrix2D, Stage};
use crate::swf_renderer::SwfRenderer;
use std::ffi::CStr;
use std::os::raw::c_int;
use swf_tree::tags::{DefineMorphShape, DefineShape};

pub struct B200Renderer {
  handle: *mut swfr_renderer,
  width: usize,
  height: usize,
  rendered: bool,
}

/// `Matrix2D([f32; 6])` (rs/src/stage.rs:12-26) has exactly the layout of `swfr_display_primitive.matrix`:
/// `[scale_x, scale_y, rotate_skew0, rotate_skew1, translate_x, translate_y]`.
fn primitive(p: &DisplayPrimitive) -> swfr_display_primitive {
  match p {
    DisplayPrimitive::Shape(s) => swfr_display_primitive {
      kind: SWFR_PRIM_SHAPE,
      id: s.id.0 as u32,
      matrix: s.matrix.0,
      ratio: 0,
      flags: 0,
      ratio_f: 0.0,
      color_transform: swfr_color_transform::default(),
    },
    DisplayPrimitive::MorphShape(m) => swfr_display_primitive {
      kind: SWFR_PRIM_MORPH_SHAPE,
      id: m.id.0 as u32,
      matrix: m.matrix.0,
      ratio: m.ratio.0, // MorphRatio(u16): 0 = start, 65535 = end (rs/src/stage.rs:28-34)
      flags: 0,
      ratio_f: 0.0,
      color_transform: swfr_color_transform::default(),
    },
  }
}

/// `swf_tree::Matrix` (Sfixed16P16 epsilons + twips) -> `Matrix2D`, as `rs/src/lib.rs:122-126` builds it for the
/// headless test.
fn matrix2d(m: &swf_tree::Matrix) -> Matrix2D {
  Matrix2D([
    m.scale_x.epsilons as f32 / 65536.0,
    m.scale_y.epsilons as f32 / 65536.0,
    m.rotate_skew0.epsilons as f32 / 65536.0,
    m.rotate_skew1.epsilons as f32 / 65536.0,
    m.translate_x as f32,
    m.translate_y as f32,
  ])
}

impl B200Renderer {
  /// One renderer per GPU; `device` = CUDA ordinal.
  pub fn on_device(device: i32, width: usize, height: usize) -> Result<Self, &'static str> {
    if unsafe { swfr_abi_version() } != SWFR_ABI_VERSION {
      return Err("libswfr_b200.so: ABI version mismatch");
    }
    let mut handle = std::ptr::null_mut();
    match unsafe { swfr_create(device as c_int, width as u32, height as u32, &mut handle) } {
      SWFR_OK => Ok(B200Renderer { handle, width, height, rendered: false }),
      // headless_renderer.rs:60-75: "Failed to find a compatible GPU adapter" (there is no CPU fallback)
      SWFR_ERR_CUDA => Err("Failed to find a compatible GPU adapter"),
      _ => Err("Failed to create the renderer"),
    }
  }

  /// `HeadlessGfxRenderer::new(&instance, width, height)`
  pub fn new(width: usize, height: usize) -> Result<Self, &'static str> {
    Self::on_device(0, width, height)
  }

  fn last_error(&self) -> String {
    unsafe { CStr::from_ptr(swfr_last_error(self.handle)) }.to_string_lossy().into_owned()
  }

  /// The reference panics / throws at these sites ("Invalid fill ID", unknown shape id, ...): same here, with the
  /// library's message.
  fn check(&self, rc: c_int, what: &str) {
    if rc != SWFR_OK {
      panic!("{}: {} ({})", what, self.last_error(), rc);
    }
  }

  /// `HeadlessGfxRenderer::define_shape`
  pub fn define_shape(&mut self, tag: &DefineShape) -> usize {
    self.register_shape(tag).0
  }

  /// `Renderer.addBitmap(tag)` of the TypeScript interface (ts/src/lib/renderer.ts:4-8); the Rust crate has no
  /// bitmap entry point yet.  `data` = the DefineBitmap payload with media type image/x-swf-bmp.
  pub fn add_bitmap_x_swf_bmp(&mut self, id: u16, data: &[u8]) -> Result<(), String> {
    match unsafe { swfr_register_bitmap_xswfbmp(self.handle, id, data.as_ptr(), data.len()) } {
      SWFR_OK => Ok(()),
      _ => Err(self.last_error()),
    }
  }

  /// Straight RGBA8 rows, `stride` bytes apart.
  pub fn add_bitmap_rgba(&mut self, id: u16, width: u32, height: u32, rgba: &[u8], stride: usize) -> Result<(), String> {
    assert!(stride >= width as usize * 4 && rgba.len() >= (height as usize - 1) * stride + width as usize * 4);
    match unsafe { swfr_register_bitmap(self.handle, id, width, height, rgba.as_ptr(), stride) } {
      SWFR_OK => Ok(()),
      _ => Err(self.last_error()),
    }
  }

  /// N stages in one set of launches (frame batches, a morph-ratio sweep): frames 0..n-1.
  pub fn render_batch(&mut self, stages: &[Stage]) {
    let prims: Vec<Vec<swfr_display_primitive>> =
      stages.iter().map(|s| s.display_root.iter().map(primitive).collect()).collect();
    let c: Vec<swfr_stage> = stages
      .iter()
      .zip(prims.iter())
      .map(|(s, p)| swfr_stage {
        background_color: swfr_rgba8 {
          r: s.background_color.r,
          g: s.background_color.g,
          b: s.background_color.b,
          a: s.background_color.a,
        },
        n_primitives: p.len() as u32,
        display_root: p.as_ptr(),
      })
      .collect();
    let rc = unsafe { swfr_render_batch(self.handle, c.as_ptr(), c.len() as u32) };
    self.check(rc, "render");
    self.rendered = true;
  }

  /// `get_image` + `download_image`: straight-alpha RGBA8, tight rows, frame `frame` of the last render.
  pub fn get_frame(&mut self, frame: u32) -> Result<Image, &'static str> {
    if !self.rendered {
      return Err("Failed to render: self.stage is None"); // headless_renderer.rs:233-236
    }
    let stride = self.width * 4;
    let mut data = vec![0u8; stride * self.height];
    match unsafe { swfr_read_image(self.handle, frame, data.as_mut_ptr(), stride, 0) } {
      SWFR_OK => Ok(Image { meta: ImageMetadata { width: self.width, height: self.height, stride }, data }),
      _ => Err("Failed to render"),
    }
  }

  /// `HeadlessGfxRenderer::get_image`
  pub fn get_image(&mut self) -> Result<Image, &'static str> {
    self.get_frame(0)
  }

  /// The windowed renderer clears to `stage.background_color` (rs/src/gfx_renderer.rs:292-301); the headless one
  /// and the TypeScript renderer start transparent (the default here).
  pub fn clear_to_background(&mut self, on: bool) {
    let rc = unsafe { swfr_set_option(self.handle, SWFR_OPT_CLEAR_TO_BACKGROUND, on as u64) };
    self.check(rc, "set_option");
  }
}

impl ClientAssetStore for B200Renderer {
  fn register_shape(&mut self, tag: &DefineShape) -> ShapeId {
    let c = to_c_define_shape(tag); // borrowed for the call only, like `&tag`
    let mut id = 0u32;
    let rc = unsafe { swfr_register_shape(self.handle, c.as_ptr(), &mut id) };
    self.check(rc, "register_shape");
    ShapeId(id as usize)
  }

  fn register_morph_shape(&mut self, tag: &DefineMorphShape) -> MorphShapeId {
    let c = to_c_define_morph_shape(tag);
    let mut id = 0u32;
    let rc = unsafe { swfr_register_morph_shape(self.handle, c.as_ptr(), &mut id) };
    self.check(rc, "register_morph_shape");
    MorphShapeId(id as usize)
  }
}

impl SwfRenderer for B200Renderer {
  /// Takes the stage by value like the trait does; asynchronous on the renderer's stream (`get_image` waits).
  fn render(&mut self, stage: Stage) -> () {
    self.render_batch(std::slice::from_ref(&stage));
  }
}

impl Renderer for B200Renderer {
  /// The older single-item API used by the headless test (rs/src/lib.rs:118-130).
  fn set_stage(&mut self, item: DisplayItem) -> () {
    let DisplayItem::Shape(id, matrix) = item;
    let prim = swfr_display_primitive {
      kind: SWFR_PRIM_SHAPE,
      id: id as u32,
      matrix: matrix2d(&matrix).0,
      ratio: 0,
      flags: 0,
      ratio_f: 0.0,
      color_transform: swfr_color_transform::default(),
    };
    let stage = swfr_stage { background_color: swfr_rgba8::default(), n_primitives: 1, display_root: &prim };
    let rc = unsafe { swfr_render(self.handle, &stage) };
    self.check(rc, "set_stage");
    self.rendered = true;
  }
}

impl Drop for B200Renderer {
  fn drop(&mut self) {
    unsafe { swfr_destroy(self.handle) } // waits for the device, then frees (headless_renderer.rs:871-904)
  }
}

// The handle is used from one thread at a time (every method takes &mut self); moving it to another thread is fine.
unsafe impl Send for B200Renderer {}
