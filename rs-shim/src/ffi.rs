//! `#[repr(C)]` mirrors of `include/swfr.h` (ABI version 2).  Field order and types follow the header line by line;
//! `tests/test_host_abi.py` checks the same layouts from Python (ctypes) against the compiled library.
#![allow(non_camel_case_types, dead_code)]

use std::os::raw::{c_char, c_int, c_void};

pub const SWFR_ABI_VERSION: u32 = 3;

pub const SWFR_OK: c_int = 0;
pub const SWFR_ERR_INVALID_HANDLE: c_int = -1;
pub const SWFR_ERR_INVALID_ID: c_int = -2;
pub const SWFR_ERR_INVALID_FILL_ID: c_int = -3;
pub const SWFR_ERR_UNSUPPORTED_STYLE: c_int = -4;
pub const SWFR_ERR_OOM: c_int = -5;
pub const SWFR_ERR_CUDA: c_int = -6;
pub const SWFR_ERR_INVALID_ARGUMENT: c_int = -7;
pub const SWFR_ERR_MALFORMED: c_int = -8;

/// swf_tree::StraightSRgba8
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct swfr_rgba8 {
  pub r: u8,
  pub g: u8,
  pub b: u8,
  pub a: u8,
}

/// swf_tree::Matrix: Sfixed16P16 epsilons + twips translation
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct swfr_swf_matrix {
  pub scale_x: i32,
  pub scale_y: i32,
  pub rotate_skew0: i32,
  pub rotate_skew1: i32,
  pub translate_x: i32,
  pub translate_y: i32,
}

pub const SWFR_FILL_SOLID: u32 = 0;
pub const SWFR_FILL_LINEAR_GRADIENT: u32 = 1;
pub const SWFR_FILL_RADIAL_GRADIENT: u32 = 2;
pub const SWFR_FILL_FOCAL_GRADIENT: u32 = 3;
pub const SWFR_FILL_BITMAP: u32 = 4;

pub const SWFR_SPREAD_PAD: u8 = 0;
pub const SWFR_SPREAD_REFLECT: u8 = 1;
pub const SWFR_SPREAD_REPEAT: u8 = 2;
pub const SWFR_COLOR_SRGB: u8 = 0;
pub const SWFR_COLOR_LINEAR_RGB: u8 = 1;

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct swfr_color_stop {
  pub ratio: u8,
  pub color: swfr_rgba8,
  pub morph_color: swfr_rgba8,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct swfr_gradient {
  pub spread: u8,
  pub color_space: u8,
  pub n_colors: u16,
  pub colors: *const swfr_color_stop,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct swfr_fill_style {
  pub type_: u32,
  pub color: swfr_rgba8,
  pub morph_color: swfr_rgba8,
  pub matrix: swfr_swf_matrix,
  pub gradient: swfr_gradient,
  pub focal_point: i16,
  pub bitmap_id: u16,
  pub repeating: u8,
  pub smoothed: u8,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct swfr_line_style {
  pub width: u16,
  pub morph_width: u16,
  pub fill: swfr_fill_style,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct swfr_styles {
  pub n_fill: u32,
  pub fill: *const swfr_fill_style,
  pub n_line: u32,
  pub line: *const swfr_line_style,
}

pub const SWFR_RECORD_EDGE: u32 = 0;
pub const SWFR_RECORD_STYLE_CHANGE: u32 = 1;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct swfr_shape_record {
  pub type_: u32,
  pub delta_x: i32,
  pub delta_y: i32,
  pub control_delta_x: i32,
  pub control_delta_y: i32,
  pub morph_delta_x: i32,
  pub morph_delta_y: i32,
  pub morph_control_delta_x: i32,
  pub morph_control_delta_y: i32,
  pub has_control_delta: u8,
  pub has_morph_control_delta: u8,
  pub has_move_to: u8,
  pub has_morph_move_to: u8,
  pub has_left_fill: u8,
  pub has_right_fill: u8,
  pub has_line_style: u8,
  pub has_new_styles: u8,
  pub move_to_x: i32,
  pub move_to_y: i32,
  pub morph_move_to_x: i32,
  pub morph_move_to_y: i32,
  pub left_fill: u32,
  pub right_fill: u32,
  pub line_style: u32,
  pub new_styles: *const swfr_styles,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct swfr_define_shape {
  pub id: u16,
  pub bounds: [i32; 4],
  pub morph_bounds: [i32; 4],
  pub initial_styles: swfr_styles,
  pub n_records: u32,
  pub records: *const swfr_shape_record,
}

pub const SWFR_PRIM_SHAPE: u32 = 0;
pub const SWFR_PRIM_MORPH_SHAPE: u32 = 1;
pub const SWFR_PRIM_RATIO_F32: u16 = 1;

/// swf-tree ColorTransformWithAlpha in raw units (mult: Sfixed8P8 epsilons, add: integers); not an input of the
/// reference's Stage - left zeroed (flag clear) by this shim
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct swfr_color_transform {
  pub red_mult: i16,
  pub green_mult: i16,
  pub blue_mult: i16,
  pub alpha_mult: i16,
  pub red_add: i16,
  pub green_add: i16,
  pub blue_add: i16,
  pub alpha_add: i16,
}

/// DisplayPrimitive::{Shape(StoredShape), MorphShape(StoredMorphShape)} - rs/src/stage.rs:36-59
#[repr(C)]
#[derive(Clone, Copy)]
pub struct swfr_display_primitive {
  pub kind: u32,
  pub id: u32,
  pub matrix: [f32; 6],
  pub ratio: u16,
  pub flags: u16,
  pub ratio_f: f32,
  pub color_transform: swfr_color_transform,
}

/// Stage - rs/src/stage.rs:4-9
#[repr(C)]
#[derive(Clone, Copy)]
pub struct swfr_stage {
  pub background_color: swfr_rgba8,
  pub n_primitives: u32,
  pub display_root: *const swfr_display_primitive,
}

pub const SWFR_OPT_RETAIN_COMPILED: u32 = 1;
pub const SWFR_OPT_FRAMES_PER_PASS: u32 = 2;
pub const SWFR_OPT_PROFILE: u32 = 3;
pub const SWFR_OPT_HOST_THREADS: u32 = 4;
pub const SWFR_OPT_CLEAR_TO_BACKGROUND: u32 = 5;
pub const SWFR_OPT_OCCLUSION_CHUNKS: u32 = 7;

pub type swfr_renderer = c_void;
pub type swfr_batch = c_void;

extern "C" {
  pub fn swfr_abi_version() -> u32;
  pub fn swfr_status_string(status: c_int) -> *const c_char;
  pub fn swfr_last_error(r: *const swfr_renderer) -> *const c_char;
  pub fn swfr_create(device: c_int, width: u32, height: u32, out: *mut *mut swfr_renderer) -> c_int;
  pub fn swfr_destroy(r: *mut swfr_renderer);
  pub fn swfr_set_option(r: *mut swfr_renderer, key: u32, value: u64) -> c_int;
  pub fn swfr_register_shape(r: *mut swfr_renderer, tag: *const swfr_define_shape, out_id: *mut u32) -> c_int;
  pub fn swfr_register_morph_shape(r: *mut swfr_renderer, tag: *const swfr_define_shape, out_id: *mut u32) -> c_int;
  pub fn swfr_register_bitmap(r: *mut swfr_renderer, id: u16, w: u32, h: u32, rgba: *const u8, stride: usize) -> c_int;
  pub fn swfr_register_bitmap_xswfbmp(r: *mut swfr_renderer, id: u16, data: *const u8, len: usize) -> c_int;
  pub fn swfr_render(r: *mut swfr_renderer, stage: *const swfr_stage) -> c_int;
  pub fn swfr_render_batch(r: *mut swfr_renderer, stages: *const swfr_stage, n: u32) -> c_int;
  pub fn swfr_batch_create(r: *mut swfr_renderer, stages: *const swfr_stage, n: u32, out: *mut *mut swfr_batch) -> c_int;
  pub fn swfr_batch_render(r: *mut swfr_renderer, b: *mut swfr_batch) -> c_int;
  pub fn swfr_batch_destroy(r: *mut swfr_renderer, b: *mut swfr_batch);
  pub fn swfr_sync(r: *mut swfr_renderer) -> c_int;
  pub fn swfr_read_image(r: *mut swfr_renderer, frame: u32, dst: *mut u8, stride: usize, premultiplied: c_int) -> c_int;
  pub fn swfr_read_frames_async(r: *mut swfr_renderer, first: u32, count: u32, dst: *mut u8) -> c_int;
  pub fn swfr_write_pam(rgba: *const u8, width: u32, height: u32, stride: usize, out: *mut u8, cap: u64, n: *mut u64) -> c_int;
}
